"""SCP-OctAttention on the scp_b200 CUDA operators.

Drop-in for the reference's ``models.oct_attention.OctAttention`` (models/oct_attention.py:9-83) with
``models.attention_model.TransformerModule`` (attention_model.py:98-155): same constructor, same
``forward(data, pos) -> logits [B,csz,255]`` and state_dict names (scp_b200/weights.py).

  oct_attention.py:52-79,85-99 + attention_model.py:20-22  -> scp_octattn_embed (both streams, sqrt(600), PE)
  attention_model.py:58-95 two-stream causal attention       -> scp_linear (K,Q,V of both streams) + scp_octattn_attention
  attention_model.py:112-125 residual + LayerNorm + FFN      -> scp_layernorm (fused residual), scp_linear (ReLU)
  oct_attention.py:82 decoder                                -> scp_linear (ReLU), scp_linear
"""
import torch
from torch import nn

from .. import weights as W
from ..ops import CudaOps, V
from .ehem import _register, _BUFFER_KINDS


class OctAttention(nn.Module):
    def __init__(self, cfg, ops=None, seed=0, sharpen=True):
        super().__init__()
        self.cfg = cfg
        m = cfg.model
        self.embed = 4 * (m.occ_embed_dim + m.level_embed_dim + m.octant_embed_dim + m.abs_pos_embed_dim)
        assert self.embed == 600 and m.head_num * 150 == 600, "kernels are specialised for the reference's 4 x 150 heads"
        self.pos_embed = bool(getattr(m, "pos_embed", True))
        self.spec = W.octattn_spec(m.context_size, self.embed, m.hidden_dimension, m.layer_num, m.token_num,
                                   m.max_octree_level, self.pos_embed)
        for (name, shape, kind), t in zip(self.spec, W.synth_state_dict(self.spec, seed, sharpen).values()):
            _register(self, name, t, kind in _BUFFER_KINDS)
        self._ops = ops
        self._sd = None

    def load_state_dict(self, *a, **k):
        self._sd = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._sd = None
        return super()._apply(fn, *a, **k)

    def _prepare(self):
        """Stable per-weight tensor objects: the operator layer keys its split-weight cache on them (scp_b200/ops.py)."""
        if self._sd is None:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            for i in range(self.cfg.model.layer_num):
                a = f"transformer_encoder.layers.{i}.attn"
                # one GEMM per stream: [key; value] of the known stream, [key; query; value] of the unknown stream
                sd[f"{a}.kv.weight"] = torch.cat([sd[f"{a}.mlp_key.weight"], sd[f"{a}.mlp_value.weight"]], 0).contiguous()
                sd[f"{a}.kv.bias"] = torch.cat([sd[f"{a}.mlp_key.bias"], sd[f"{a}.mlp_value.bias"]], 0).contiguous()
                sd[f"{a}.kqv.weight"] = torch.cat([sd[f"{a}.mlp_key.weight"], sd[f"{a}.mlp_query.weight"], sd[f"{a}.mlp_value.weight"]], 0).contiguous()
                sd[f"{a}.kqv.bias"] = torch.cat([sd[f"{a}.mlp_key.bias"], sd[f"{a}.mlp_query.bias"], sd[f"{a}.mlp_value.bias"]], 0).contiguous()
            self._sd = sd
        return self._sd

    @property
    def ops(self):
        if self._ops is None:
            self._ops = CudaOps()
        return self._ops

    @torch.no_grad()
    def forward_ragged(self, ctx, ctx_pos, offsets, pos_scale):
        """ctx uint8 [T,4,3] (level, octant, occ), ctx_pos int32 [T,4,3] ancestor cell origins,
        pos_scale = 1/2^max_level (encode_dataset.py:48).  Returns logits [T,255]."""
        ops = self.ops
        sd = self._prepare()
        m = self.cfg.model
        T = ctx.shape[0]
        seqs = ops.seqs(offsets)
        assert max(seqs.lengths) <= m.context_size
        ctx = ctx.reshape(T, 12).contiguous()
        ctx_pos = ctx_pos.reshape(T, 12).contiguous()
        like = torch.empty(0, dtype=torch.float32, device=ctx.device)
        E, EU = ops.empty(T, 600, like), ops.empty(T, 600, like)
        level_base = 10 if self.cfg.train.type == "obj" else 12            # oct_attention.py:57-60
        ops.octattn_embed(ctx, ctx_pos, float(pos_scale), level_base, m.max_octree_level, seqs, sd, E, EU)
        for i in range(m.layer_num):
            p = f"transformer_encoder.layers.{i}"
            # K, V of the known stream; K, Q, V of the unknown stream (attention_model.py:65-70), one buffer so
            # that all five operands share a row stride
            ALL = ops.empty(T, 3000, like)          # columns: K | V of the known stream, K | Q | V of the unknown stream
            ops.linear(V(E), sd[f"{p}.attn.kv.weight"], sd[f"{p}.attn.kv.bias"], V(ALL, 0, 1200))
            ops.linear(V(EU), sd[f"{p}.attn.kqv.weight"], sd[f"{p}.attn.kqv.bias"], V(ALL, 1200, 1800))
            A, AU = ops.empty(T, 600, like), ops.empty(T, 600, like)
            ops.octattn_attention(V(ALL, 1800, 600), V(ALL, 0, 600), V(ALL, 1200, 600), V(ALL, 600, 600),
                                  V(ALL, 2400, 600), m.head_num, 150, seqs, V(A), V(AU))
            for src, att in ((E, A), (EU, AU)):
                x1 = ops.empty(T, 600, like)
                ops.layernorm(V(att), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], V(x1), res=V(src))
                h = ops.empty(T, m.hidden_dimension, like)
                ops.linear(V(x1), sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"], V(h), act="relu")
                f = ops.empty(T, 600, like)
                ops.linear(V(h), sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"], V(f))
                ops.layernorm(V(f), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], V(src), res=V(x1))
        d0 = ops.empty(T, 600, like)
        ops.linear(V(EU), sd["decoder0.weight"], sd["decoder0.bias"], V(d0), act="relu")
        logits = ops.empty(T, m.token_num, like)
        ops.linear(V(d0), sd["decoder1.weight"], sd["decoder1.bias"], V(logits))
        return logits

    @torch.no_grad()
    def forward(self, data, pos=None):
        """Reference interface (oct_attention.py:48-83): data int64 [B,csz,4,3] (occ, level, octant),
        pos float32 [B,csz,4,3] (already divided by 2^max_level) -> logits [B,csz,255]."""
        B, csz = data.shape[:2]
        ctx = torch.stack((data[..., 1], data[..., 2], data[..., 0]), -1).to(torch.uint8)     # -> (level, octant, occ)
        scale = 1.0 / float(1 << 21)
        ipos = torch.round(pos.double() * (1 << 21)).to(torch.int32)      # exact: pos = int / 2^max_level
        out = self.forward_ragged(ctx.reshape(B * csz, 4, 3), ipos.reshape(B * csz, 4, 3),
                                  [i * csz for i in range(B + 1)], scale)
        return out.reshape(B, csz, -1)
