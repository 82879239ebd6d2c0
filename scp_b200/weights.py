"""Parameter inventory (state_dict names/shapes) of SCP-EHEM and SCP-OctAttention, and a seeded
synthetic initialiser.

The names and shapes are the reference's (models/ehem.py:11-70, models/dgcnn.py:75-119,
models/swin_transformer.py:406-434,583-595,335-340, models/oct_attention.py:10-46,
models/attention_model.py:25-37,98-106) so a real checkpoint's ``state_dict`` loads unchanged.
No checkpoints exist offline, so tests and benchmarks use ``synth_state_dict``: every tensor is
a deterministic function of (seed, name), independent of module construction order, which lets
the reference model (in the build container) and this package share bit-identical weights.
"""
import math
import zlib

import numpy as np
import torch

EHEM_SELF_DEPTHS = (4, 4, 4, 4, 2)     # ehem.py:17-23
EHEM_CROSS_DEPTHS = (2, 2, 1, 1)       # ehem.py:25-31
SWIN_DIM = 256
SWIN_HEADS = 4
SWIN_WINDOW = 512


def _mlp(prefix, dims):
    out = []
    for i in range(len(dims) - 1):
        out.append((f"{prefix}.{2 * i}.weight", (dims[i + 1], dims[i]), "linear_w"))
        out.append((f"{prefix}.{2 * i}.bias", (dims[i + 1],), "linear_b"))
    return out


def _swin(prefix, depths):
    out = []
    d = SWIN_DIM
    for i, depth in enumerate(depths):
        for j in range(depth):
            b = f"{prefix}.layers.{i}.blocks.{j}"
            out += [(f"{b}.layernorm_before.weight", (d,), "ln_w"),
                    (f"{b}.layernorm_before.bias", (d,), "ln_b"),
                    (f"{b}.attention.self.relative_position_bias_table", (2 * SWIN_WINDOW - 1, SWIN_HEADS), "relpos"),
                    (f"{b}.attention.self.relative_position_index", (SWIN_WINDOW, SWIN_WINDOW), "relpos_index")]
            for nm in ("query", "key", "value"):
                out += [(f"{b}.attention.self.{nm}.weight", (d, d), "linear_w"),
                        (f"{b}.attention.self.{nm}.bias", (d,), "linear_b")]
            out += [(f"{b}.attention.output.dense.weight", (d, d), "linear_w"),
                    (f"{b}.attention.output.dense.bias", (d,), "linear_b"),
                    (f"{b}.layernorm_after.weight", (d,), "ln_w"),
                    (f"{b}.layernorm_after.bias", (d,), "ln_b"),
                    (f"{b}.intermediate.dense.weight", (4 * d, d), "linear_w"),
                    (f"{b}.intermediate.dense.bias", (4 * d,), "linear_b"),
                    (f"{b}.output.dense.weight", (d, 4 * d), "linear_w"),
                    (f"{b}.output.dense.bias", (d,), "linear_b")]
        if i < len(depths) - 1:
            p = f"{prefix}.layers.{i}.downsample"
            out += [(f"{p}.reduction.weight", (d, 2 * d), "linear_w"),
                    (f"{p}.norm.weight", (2 * d,), "ln_w"),
                    (f"{p}.norm.bias", (2 * d,), "ln_b")]
    return out


def ehem_spec(max_level=19):
    g = "geo_feat_generator"
    out = []
    for name, cin, cout in (("conv1", 6, 64), ("conv2", 288, 128), ("conv3", 384, 256)):
        out += [(f"{g}.{name}.0.weight", (cout, cin, 1, 1), "conv_w"),
                (f"{g}.{name}.1.weight", (cout,), "bn_w"),
                (f"{g}.{name}.1.bias", (cout,), "bn_b"),
                (f"{g}.{name}.1.running_mean", (cout,), "bn_mean"),
                (f"{g}.{name}.1.running_var", (cout,), "bn_var"),
                (f"{g}.{name}.1.num_batches_tracked", (), "counter")]
    out += [(f"{g}.occ_enc.weight", (256, 16), "embed"),
            (f"{g}.level_enc.weight", (max_level, 4), "embed"),
            (f"{g}.octant_enc.weight", (9, 4), "embed")]
    out += _mlp(f"{g}.mlp2", (80, 80, 64, 64))
    out += _mlp(f"{g}.mlp3", (64, 128, 128, 128))
    out += _mlp(f"{g}.edge_mlp1", (448, 256, 256, 256))
    out += _mlp(f"{g}.edge_mlp2", (512, 256, 256, 128))
    out += _swin("swin_self_transformer", EHEM_SELF_DEPTHS)
    out += _swin("swin_cross_transformer", EHEM_CROSS_DEPTHS)
    out += _mlp("ancient_mlp", (1280, 1024, 512, 256))
    out += _mlp("prob_pred_mlp1", (256, 256, 256, 255))
    out += _mlp("pre_occ_mlp", (16, 16, 16, 16))
    out += _mlp("pre_attn_mlp", (256, 256, 240, 240))
    out += _mlp("prob_pred_mlp2", (1280, 768, 512, 255))
    return out


def octattn_spec(context_size=1024, embed=600, hidden=300, layers=3, token_num=255,
                 max_octree_level=12, pos_embed=True):
    out = [("mask", (context_size, context_size), "causal_mask")]
    for i in range(layers):
        p = f"transformer_encoder.layers.{i}"
        for nm in ("mlp_key", "mlp_query", "mlp_value"):
            out += [(f"{p}.attn.{nm}.weight", (embed, embed), "linear_w"),
                    (f"{p}.attn.{nm}.bias", (embed,), "linear_b")]
        out += [(f"{p}.linear1.weight", (hidden, embed), "linear_w"), (f"{p}.linear1.bias", (hidden,), "linear_b"),
                (f"{p}.linear2.weight", (embed, hidden), "linear_w"), (f"{p}.linear2.bias", (embed,), "linear_b"),
                (f"{p}.norm1.weight", (embed,), "ln_w"), (f"{p}.norm1.bias", (embed,), "ln_b"),
                (f"{p}.norm2.weight", (embed,), "ln_w"), (f"{p}.norm2.bias", (embed,), "ln_b")]
    if pos_embed:                                   # attention_model.py:142-144: the module exists only with cfg.model.pos_embed
        out += [("transformer_encoder.position_enc.pe", (context_size, embed), "sin_pe")]
    out += [("occ_enc.weight", (token_num + 1, 128), "embed"),
            ("level_enc.weight", (max_octree_level + 1, 6), "embed"),
            ("octant_enc.weight", (9, 4), "embed"),
            ("abs_pos_enc.weight", (12, 3), "linear_w"), ("abs_pos_enc.bias", (12,), "linear_b"),
            ("decoder0.weight", (embed, embed), "linear_w"), ("decoder0.bias", (embed,), "linear_b"),
            ("decoder1.weight", (token_num, embed), "linear_w"), ("decoder1.bias", (token_num,), "linear_b")]
    return out


def _rng(seed, name):
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0xFFFFFFFF)


def synth_state_dict(spec, seed=0, sharpen=True):
    """Seeded weights.  ``sharpen=False`` mimics torch's default init statistics (near-uniform
    PMFs); ``sharpen=True`` is the "random-init+" of SURVEY.md section 8c: non-trivial
    rel-pos tables / BatchNorm statistics / LayerNorm affine and a scaled final classifier so
    that the PMFs are peaked and parity tolerances mean something."""
    sd = {}
    is_oct = any(name == "decoder1.weight" for name, _, _ in spec)
    # EHEM: default-init MLP chains shrink the token-dependent signal, so "random-init+" also applies a
    # global weight gain and shrinks biases; OctAttention only needs the scaled classifier.
    gain, bias_scale, last = (1.0, 1.0, 30.0) if is_oct else (2.0, 0.2, 2.0)
    if not sharpen:
        gain, bias_scale, last = 1.0, 1.0, 1.0
    fan_in_of = {}
    for name, shape, kind in spec:
        if kind in ("linear_w", "conv_w"):
            fan_in_of[name.rsplit(".", 1)[0]] = int(np.prod(shape[1:]))
    for name, shape, kind in spec:
        r = _rng(seed, name)
        if kind in ("linear_w", "conv_w"):
            bound = 1.0 / math.sqrt(int(np.prod(shape[1:])))
            w = r.uniform(-bound, bound, shape) * gain
            if name.startswith(("prob_pred_mlp1.4", "prob_pred_mlp2.4", "decoder1")):
                w = w * last
            t = torch.from_numpy(w.astype(np.float32))
        elif kind == "linear_b":
            bound = 1.0 / math.sqrt(fan_in_of[name.rsplit(".", 1)[0]])
            t = torch.from_numpy((r.uniform(-bound, bound, shape) * bias_scale).astype(np.float32))
        elif kind == "embed":
            t = torch.from_numpy(r.normal(0, 1, shape).astype(np.float32))
        elif kind in ("ln_w", "bn_w"):
            v = 1.0 + (r.normal(0, 0.2, shape) if sharpen else 0.0)
            if kind == "bn_w" and sharpen:
                v = v * np.where(r.random_sample(shape) < 0.15, -1.0, 1.0)   # exercise negative BN scales
            t = torch.from_numpy(np.broadcast_to(v, shape).astype(np.float32).copy())
        elif kind in ("ln_b", "bn_b"):
            v = r.normal(0, 0.1, shape) if sharpen else np.zeros(shape)
            t = torch.from_numpy(v.astype(np.float32))
        elif kind == "bn_mean":
            v = r.normal(0, 0.2, shape) if sharpen else np.zeros(shape)
            t = torch.from_numpy(v.astype(np.float32))
        elif kind == "bn_var":
            v = r.uniform(0.5, 1.5, shape) if sharpen else np.ones(shape)
            t = torch.from_numpy(v.astype(np.float32))
        elif kind == "counter":
            t = torch.zeros((), dtype=torch.int64)
        elif kind == "relpos":
            v = r.normal(0, 0.5, shape) if sharpen else np.zeros(shape)
            t = torch.from_numpy(v.astype(np.float32))
        elif kind == "relpos_index":
            i = torch.arange(shape[0])
            t = (i[:, None] - i[None, :] + shape[0] - 1).to(torch.int64)    # swin_transformer.py:425-429
        elif kind == "causal_mask":
            t = torch.full(shape, float("-inf")).triu(1)                     # oct_attention.py:38-46
        elif kind == "sin_pe":
            pos = torch.arange(0, shape[0], dtype=torch.float).unsqueeze(1)  # attention_model.py:12-16
            div = torch.exp(torch.arange(0, shape[1], 2).float() * (-math.log(10000.0) / shape[1]))
            t = torch.zeros(shape)
            t[:, 0::2] = torch.sin(pos * div)
            t[:, 1::2] = torch.cos(pos * div)
        else:
            raise ValueError(kind)
        sd[name] = t
    return sd
