"""SCP-EHEM entropy model on the scp_b200 CUDA operators.

Drop-in for the reference's ``models.ehem.EHEM`` (models/ehem.py:10-136): same constructor
(``EHEM(cfg)``), same ``forward(data, pos, enc=True) -> (logits1, logits2)`` contract, same
``state_dict`` names/shapes (scp_b200/weights.py).  The arithmetic is NOT torch: every layer is a
call into libscp_b200.so.  Besides the reference interface there is ``forward_ragged`` which runs
any number of context windows of any length as ONE ragged batch (windows are independent at
encode time, SURVEY.md section 5 "long-context").

Layer map (reference file:line -> operator):
  dgcnn.py:122-129 embeddings            -> scp_ehem_embed
  dgcnn.py:10-28 knn                      -> scp_knn
  dgcnn.py:48-71,79-87,133-144 edge conv  -> scp_linear ([Wa; Wb-Wa]) + scp_edge_gather_max
  dgcnn.py:92-119 MLPs, ehem.py:34-68     -> scp_linear (fused bias + LeakyReLU)
  swin_transformer.py:583-706 SwinLayer   -> scp_layernorm, scp_linear (QKV), scp_swin_attention,
                                             scp_linear (+residual), scp_linear (GELU), scp_linear (+residual)
  swin_transformer.py:322-367 merging     -> scp_pair_concat, scp_layernorm, scp_linear
  ehem.py:72-86 concat_states             -> scp_copy_cols / scp_upsample_cols
"""
import os

import torch
from torch import nn

from .. import weights as W
from ..ops import CudaOps, V


class _Node(nn.Module):
    """Anonymous container so that parameters registered under dotted names reproduce the
    reference's state_dict keys."""


def _register(root, name, tensor, is_buffer):
    parts = name.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


_BUFFER_KINDS = ("bn_mean", "bn_var", "counter", "relpos_index", "causal_mask", "sin_pe")


class EHEM(nn.Module):
    def __init__(self, cfg, ops=None, seed=0, sharpen=True):
        super().__init__()
        self.cfg = cfg
        self.spec = W.ehem_spec(cfg.model.max_level)
        for (name, shape, kind), t in zip(self.spec, W.synth_state_dict(self.spec, seed, sharpen).values()):
            _register(self, name, t, kind in _BUFFER_KINDS)
        self._ops = ops
        self._prep = None
        self.k = 20                                           # dgcnn.py:75

    # -- plumbing ---------------------------------------------------------------------------
    @property
    def ops(self):
        if self._ops is None:
            self._ops = CudaOps()
        return self._ops

    def load_state_dict(self, *a, **k):
        self._prep = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._prep = None
        return super()._apply(fn, *a, **k)

    def _prepare(self):
        """Derived weights: folded BatchNorm, [Wa; Wb-Wa] edge-conv weights, fused QKV."""
        if self._prep is not None:
            return self._prep
        if self._ops is not None and hasattr(self._ops, "lib"):
            self._ops.lib.scp_gemm_cache_clear()              # weights may have changed in place
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        P = {"sd": sd}
        g = "geo_feat_generator"
        for name in ("conv1", "conv2", "conv3"):
            w = sd[f"{g}.{name}.0.weight"][:, :, 0, 0]
            d = w.shape[1] // 2
            wa, wb = w[:, :d], w[:, d:]
            P[f"{name}.w"] = torch.cat([wa, wb - wa], 0).contiguous()          # uv = x @ [Wa; Wb-Wa]^T
            s = sd[f"{g}.{name}.1.weight"] / torch.sqrt(sd[f"{g}.{name}.1.running_var"] + 1e-5)
            P[f"{name}.s"] = s.contiguous()
            P[f"{name}.t"] = (sd[f"{g}.{name}.1.bias"] - sd[f"{g}.{name}.1.running_mean"] * s).contiguous()
        for enc, depths in (("swin_self_transformer", W.EHEM_SELF_DEPTHS), ("swin_cross_transformer", W.EHEM_CROSS_DEPTHS)):
            for i, depth in enumerate(depths):
                for j in range(depth):
                    a = f"{enc}.layers.{i}.blocks.{j}.attention.self"
                    P[f"{a}.qkv.w"] = torch.cat([sd[f"{a}.query.weight"], sd[f"{a}.key.weight"], sd[f"{a}.value.weight"]], 0).contiguous()
                    P[f"{a}.qkv.b"] = torch.cat([sd[f"{a}.query.bias"], sd[f"{a}.key.bias"], sd[f"{a}.value.bias"]], 0).contiguous()
                    P[f"{a}.kv.w"] = P[f"{a}.qkv.w"][256:]
                    P[f"{a}.kv.b"] = P[f"{a}.qkv.b"][256:]
        self._prep = P
        return P

    # -- building blocks --------------------------------------------------------------------
    def _mlp(self, prefix, x, out, acts=("leaky", "leaky", "none"), **kw):
        """nn.Sequential(Linear, LeakyReLU, Linear, LeakyReLU, Linear); ``out`` is the view to write."""
        ops, sd = self.ops, self._prep["sd"]
        rows = kw.get("rows", x[0].shape[0])
        eng = kw.pop("engine", None)
        cur, first = x, True
        for li, act in zip((0, 2, 4), acts):
            w, b = sd[f"{prefix}.{li}.weight"], sd[f"{prefix}.{li}.bias"]
            dst = out if li == 4 else V(ops.empty(rows, w.shape[0], x[0]))
            if first:
                ops.linear(cur, w, b, dst, act=act, engine=eng, **kw)
                first = False
            else:
                ops.linear(cur, w, b, dst, act=act, engine=eng)
            cur = dst
        return out

    def _swin_layer(self, pre, h, seqs, shift, query=None, dst=None):
        ops, P = self.ops, self._prep
        sd = P["sd"]
        T = h.shape[0]
        a = f"{pre}.attention.self"
        ln = ops.empty(T, 256, h)
        ops.layernorm(V(h), sd[f"{pre}.layernorm_before.weight"], sd[f"{pre}.layernorm_before.bias"], V(ln))
        att = ops.empty(T, 256, h)
        qb, kb, vb = sd[f"{a}.query.bias"], sd[f"{a}.key.bias"], sd[f"{a}.value.bias"]
        rel = sd[f"{a}.relative_position_bias_table"]
        if query is None:
            qkv = ops.empty(T, 768, h)
            ops.linear(V(ln), P[f"{a}.qkv.w"], P[f"{a}.qkv.b"], V(qkv))
            ops.swin_attention(V(qkv, 0, 256), V(qkv, 256, 256), V(qkv, 512, 256), qb, kb, vb, rel, W.SWIN_HEADS, seqs,
                               shift, V(att))
        else:                                                   # cross: K,V from the hidden stream, Q from `query`
            lq = ops.empty(T, 256, h)
            ops.layernorm(V(query), sd[f"{pre}.layernorm_before.weight"], sd[f"{pre}.layernorm_before.bias"], V(lq))
            q = ops.empty(T, 256, h)
            ops.linear(V(lq), sd[f"{a}.query.weight"], qb, V(q))
            kv = ops.empty(T, 512, h)
            ops.linear(V(ln), P[f"{a}.kv.w"], P[f"{a}.kv.b"], V(kv))
            ops.swin_attention(V(q), V(kv, 0, 256), V(kv, 256, 256), qb, kb, vb, rel, W.SWIN_HEADS, seqs, shift, V(att))
        h2 = ops.empty(T, 256, h)
        ops.linear(V(att), sd[f"{pre}.attention.output.dense.weight"], sd[f"{pre}.attention.output.dense.bias"], V(h2),
                   res=V(h))
        ops.layernorm(V(h2), sd[f"{pre}.layernorm_after.weight"], sd[f"{pre}.layernorm_after.bias"], V(ln))
        mid = ops.empty(T, 1024, h)
        ops.linear(V(ln), sd[f"{pre}.intermediate.dense.weight"], sd[f"{pre}.intermediate.dense.bias"], V(mid), act="gelu")
        h3 = ops.empty(T, 256, h) if dst is None else dst            # dst: a [T,256] column slice of the concatenated states
        ops.linear(V(mid), sd[f"{pre}.output.dense.weight"], sd[f"{pre}.output.dense.bias"], V(h3), res=V(h2))
        return h3

    def _merge(self, pre, h, seqs):
        ops, sd = self.ops, self._prep["sd"]
        dst = seqs.half()
        pc = ops.empty(dst.total, 512, h)
        ops.pair_concat(V(h), seqs, dst, V(pc))
        ln = ops.empty(dst.total, 512, h)
        ops.layernorm(V(pc), sd[f"{pre}.norm.weight"], sd[f"{pre}.norm.bias"], V(ln))
        out = ops.empty(dst.total, 256, h)
        ops.linear(V(ln), sd[f"{pre}.reduction.weight"], None, V(out))
        return out

    def _swin_encoder(self, enc, depths, h, seqs, out, query=None):
        """SwinEncoder.forward + EHEM.concat_states: writes [stage0 | up(stage1) | up^2(stage2) ...] into the
        first 256*len(depths) columns of ``out`` (rows = finest tokens)."""
        ops = self.ops
        fine = seqs
        for i, depth in enumerate(depths):
            for j in range(depth):
                # the last block of stage 0 writes its output straight into the first 256 columns of `out` (no copy)
                h = self._swin_layer(f"{enc}.layers.{i}.blocks.{j}", h, seqs, 0 if j % 2 == 0 else W.SWIN_WINDOW // 2, query,
                                     dst=out[:, 0:256] if (i == 0 and j == depth - 1) else None)
            if i > 0:
                ops.upsample_cols(V(h), seqs, fine, i, V(out, 256 * i, 256))
            if i < len(depths) - 1:
                pre = f"{enc}.layers.{i}.downsample"
                h = self._merge(pre, h, seqs)
                if query is not None:
                    query = self._merge(pre, query, seqs)
                seqs = seqs.half()

    # -- forward ----------------------------------------------------------------------------
    @torch.no_grad()
    def forward_ragged(self, ctx, pos, offsets):
        """ctx uint8 [T,4,3] (level, octant, occ; self occupancy is ignored), pos float32 [T,3]; ``offsets``
        cut the token stream into context windows, every window of EVEN length (the caller appends the pad token
        of ehem.py:92-99 to odd windows).  Returns (logits1 [T/2,255] for even tokens, logits2 [T/2,255] for odd)."""
        feat_a, logits1 = self.phase1(ctx, pos, offsets)
        return logits1, self.phase2(ctx, feat_a, offsets)

    @torch.no_grad()
    def phase1(self, ctx, pos, offsets):
        """Everything that depends on the ancestors only (ehem.py:100-113; the first call of EHEM.decode, :138-160):
        returns (feat_a [T,256], logits1 [T/2,255]).  The self occupancy column of ``ctx`` is not read."""
        ops = self.ops
        P = self._prepare()
        sd = P["sd"]
        g = "geo_feat_generator"
        T = ctx.shape[0]
        assert all((b - a) % 2 == 0 for a, b in zip(offsets[:-1], offsets[1:])), "windows must have even length"
        ctx = ctx.reshape(T, 12).contiguous()
        pos = pos.contiguous()
        seqs = ops.seqs(offsets)
        k = self.k
        # DGCNN feature generator --------------------------------------------------------
        P123 = ops.empty(T, 448, pos)
        F2 = ops.empty(T, 144, pos)
        F3 = ops.empty(T, 192, pos)
        EC2 = ops.empty(T, 512, pos)
        FEAT = ops.empty(T, 256, pos)
        ops.ehem_embed(ctx, sd[f"{g}.occ_enc.weight"], sd[f"{g}.level_enc.weight"], sd[f"{g}.octant_enc.weight"], V(F2, 64, 80))
        idx = ops.knn(V(pos), seqs, k)
        uv = ops.empty(T, 128, pos)
        ops.linear(V(pos), P["conv1.w"], None, V(uv), engine="simt")
        ops.edge_gather_max(V(uv), 64, idx, P["conv1.s"], P["conv1.t"], V(P123, 0, 64), y2=V(F2, 0, 64))
        idx = ops.knn(V(F2), seqs, k)
        uv = ops.empty(T, 256, pos)
        # conv2 / mlp2 feed kNN #3 (neighbour sets): error-compensated 3xFP16 products keep them at fp32 accuracy (parity,
        # bpp and round-trip tests unchanged, +1.9 % frames/s over the fp32 SIMT GEMM = "simt"); encoder and decoder read
        # the same setting
        geo = os.environ.get("SCP_GEO_ENGINE", "auto")
        ops.linear(V(F2), P["conv2.w"], None, V(uv), engine=geo)
        ops.edge_gather_max(V(uv), 128, idx, P["conv2.s"], P["conv2.t"], V(P123, 64, 128), y2=V(F3, 0, 128))
        self._mlp(f"{g}.mlp2", V(F2, 64, 80), V(F3, 128, 64), engine=geo)
        idx = ops.knn(V(F3), seqs, k)
        uv = ops.empty(T, 512, pos)
        ops.linear(V(F3), P["conv3.w"], None, V(uv))
        ops.edge_gather_max(V(uv), 256, idx, P["conv3.s"], P["conv3.t"], V(P123, 192, 256), y2=V(EC2, 0, 256))
        self._mlp(f"{g}.mlp3", V(F3, 128, 64), V(FEAT, 0, 128))
        self._mlp(f"{g}.edge_mlp1", V(P123), V(EC2, 256, 256))
        self._mlp(f"{g}.edge_mlp2", V(EC2), V(FEAT, 128, 128))
        del uv, idx, P123, F2, F3, EC2
        # self Swin encoder + ancient MLP ------------------------------------------------
        SELF = ops.empty(T, 1280, pos)
        self._swin_encoder("swin_self_transformer", W.EHEM_SELF_DEPTHS, FEAT, seqs, SELF)
        feat_a = ops.empty(T, 256, pos)
        self._mlp("ancient_mlp", V(SELF), V(feat_a))
        del SELF, FEAT
        H = T // 2
        logits1 = ops.empty(H, 255, pos)
        self._mlp("prob_pred_mlp1", V(feat_a), V(logits1), row_step=2, row_off=0, rows=H)
        return feat_a, logits1

    @torch.no_grad()
    def phase2(self, ctx, feat_a, offsets):
        """Group 2 (ehem.py:117-125; the second call of EHEM.decode, :162-180): logits of the odd tokens given the
        occupancy of the even tokens (read from the self column of ``ctx`` at even positions) -> logits2 [T/2,255]."""
        ops = self.ops
        P = self._prepare()
        sd = P["sd"]
        g = "geo_feat_generator"
        T = ctx.shape[0]
        ctx = ctx.reshape(T, 12).contiguous()
        pos = feat_a
        H = T // 2
        half = ops.seqs([o // 2 for o in offsets])
        # group 2: cross Swin with the even tokens' true occupancy ---------------------------
        PRE = ops.empty(H, 256, pos)
        occ16 = ops.empty(H, 16, pos)
        ops.ehem_embed_occ(ctx, sd[f"{g}.occ_enc.weight"], V(occ16))
        self._mlp("pre_occ_mlp", V(occ16), V(PRE, 0, 16))
        self._mlp("pre_attn_mlp", V(feat_a), V(PRE, 16, 240), row_step=2, row_off=0, rows=H)
        CROSS = ops.empty(H, 1280, pos)
        ops.copy_cols(V(feat_a), V(CROSS, 1024, 256), row_step=2, row_off=1, rows=H)
        fa2 = CROSS[:, 1024:1280]                                 # the odd tokens' ancestor features: query stream of the cross encoder
        self._swin_encoder("swin_cross_transformer", W.EHEM_CROSS_DEPTHS, PRE, half, CROSS, query=fa2)
        logits2 = ops.empty(H, 255, pos)
        self._mlp("prob_pred_mlp2", V(CROSS), V(logits2))
        return logits2

    @torch.no_grad()
    def forward(self, data, pos, enc=True):
        """Reference interface (models/ehem.py:88-136): data int64 [B,csz,4,3] (level, octant, occ),
        pos float32 [B,3,csz] -> (logits1 [B,ceil(csz/2),255], logits2 [B,floor(csz/2),255])."""
        B, csz = data.shape[0], data.shape[1]
        ctx = data.to(torch.uint8)
        p = pos.transpose(1, 2).to(torch.float32)
        padded = csz % 2 == 1
        if padded:                                               # ehem.py:92-99
            pad = torch.zeros_like(ctx[:, :1])
            pad[:, :, :, 2] = 255
            ctx = torch.cat((ctx, pad), 1)
            p = torch.cat((p, torch.zeros_like(p[:, :1])), 1)
        c2 = ctx.shape[1]
        l1, l2 = self.forward_ragged(ctx.reshape(B * c2, 4, 3).contiguous(), p.reshape(B * c2, 3).contiguous(),
                                     [i * c2 for i in range(B + 1)])
        l1 = l1.reshape(B, c2 // 2, 255)
        l2 = l2.reshape(B, c2 // 2, 255)
        if padded:
            l2 = l2[:, :-1]
        return l1, l2

    @torch.no_grad()
    def decode(self, data, pos, pre_occ=None):
        """Reference interface (models/ehem.py:138-180): first call (``pre_occ`` None) -> logits of the even tokens
        [B,ceil(csz/2),255] and caches the ancestor features; second call with the decoded occupancy symbols of the even
        tokens (int [B,ceil(csz/2)], 0..254) -> logits of the odd tokens [B,floor(csz/2),255]."""
        B, csz = data.shape[0], data.shape[1]
        if pre_occ is None:
            ctx = data.to(torch.uint8)
            p = pos.transpose(1, 2).to(torch.float32)
            if csz % 2 == 1:                                         # ehem.py:142-149
                pad = torch.zeros_like(ctx[:, :1])
                pad[:, :, :, 2] = 255
                ctx = torch.cat((ctx, pad), 1)
                p = torch.cat((p, torch.zeros_like(p[:, :1])), 1)
            c2 = ctx.shape[1]
            ctx = ctx.reshape(B * c2, 4, 3).contiguous()
            offsets = [i * c2 for i in range(B + 1)]
            feat_a, l1 = self.phase1(ctx, p.reshape(B * c2, 3).contiguous(), offsets)
            self._dec = (ctx, feat_a, offsets, c2)
            return l1.reshape(B, c2 // 2, 255)
        ctx, feat_a, offsets, c2 = self._dec
        ctx[0::2, 3, 2] = pre_occ.reshape(-1).to(torch.uint8)
        l2 = self.phase2(ctx, feat_a, offsets).reshape(B, c2 // 2, 255)
        return l2[:, :-1] if csz % 2 == 1 else l2
