"""TEST INFRASTRUCTURE: a plain-torch (CPU, fp32) emulation of the operator contracts of
include/scp_b200.h (A8-A12).  It lets the host-side orchestration in scp_b200/models be checked against
the reference's golden logits without a GPU, and documents each operator's semantics.  The product never
imports this file."""
import math

import torch
import torch.nn.functional as F


class EmuSeqs:
    def __init__(self, offsets):
        self.offsets = [int(o) for o in offsets]
        self._half = None

    @property
    def total(self):
        return self.offsets[-1] - self.offsets[0]

    @property
    def lengths(self):
        return [b - a for a, b in zip(self.offsets[:-1], self.offsets[1:])]

    def half(self):
        if self._half is None:
            o = [0]
            for n in self.lengths:
                o.append(o[-1] + (n + 1) // 2)
            self._half = EmuSeqs(o)
        return self._half


def _v(view):
    t, col, n = view
    return t[:, col:col + n]


class EmuOps:
    name = "emu"

    def seqs(self, offsets):
        return EmuSeqs(offsets)

    def empty(self, rows, cols, like):
        return torch.full((rows, cols), float("nan"), dtype=torch.float32)

    def linear(self, x, w, b, y, act="none", res=None, row_step=1, row_off=0, rows=None, engine=None):
        xv = _v(x)[row_off::row_step]
        rows = _v(y).shape[0] if rows is None else rows
        o = F.linear(xv[:rows], w, b)
        o = {"none": lambda t: t, "leaky": lambda t: F.leaky_relu(t, 0.01), "gelu": F.gelu, "relu": F.relu}[act](o)
        if res is not None:
            o = o + _v(res)[:rows]
        _v(y)[:rows] = o

    def layernorm(self, x, g, b, y, res=None, eps=1e-5):
        xv = _v(x) if res is None else _v(x) + _v(res)
        _v(y)[:] = F.layer_norm(xv, (xv.shape[1],), g, b, eps)

    def ehem_embed(self, ctx, occ_enc, level_enc, octant_enc, y):
        c = ctx.reshape(-1, 4, 3).long()
        o = torch.cat([occ_enc[c[:, :3, 2]].reshape(len(c), -1), level_enc[c[:, :, 0]].reshape(len(c), -1),
                       octant_enc[c[:, :, 1]].reshape(len(c), -1)], 1)
        _v(y)[:] = o

    def ehem_embed_occ(self, ctx, occ_enc, y):
        c = ctx.reshape(-1, 4, 3).long()
        _v(y)[:] = occ_enc[c[0::2, 3, 2]]

    def knn(self, x, seqs, k):
        xv = _v(x)
        idx = torch.empty((xv.shape[0], k), dtype=torch.int32)
        for a, b in zip(seqs.offsets[:-1], seqs.offsets[1:]):
            s = xv[a:b]
            xx = (s * s).sum(1)
            pd = 2 * (s @ s.T) - xx[None, :] - xx[:, None]
            kk = min(k, b - a)
            top = pd.topk(kk, dim=1)[1] + a
            if kk < k:
                top = torch.cat([top, torch.arange(a, b)[:, None].expand(-1, k - kk)], 1)
            idx[a:b] = top.int()
        return idx

    def edge_gather_max(self, uv, C, idx, s, t, y, y2=None):
        u = _v(uv)
        nb = u[:, :C][idx.long()]                      # [n,k,C]
        sel = torch.where(s >= 0, nb.max(1)[0], nb.min(1)[0])
        _v(y)[:] = F.leaky_relu(s * (sel + u[:, C:2 * C]) + t, 0.2)
        if y2 is not None:
            _v(y2)[:] = _v(y)

    def swin_attention(self, q, k, v, qb, kb, vb, relpos, heads, seqs, shift, y):
        qv, kv, vv, out = _v(q), _v(k), _v(v), _v(y)
        ws = 512
        i = torch.arange(ws)
        bias = relpos[(i[:, None] - i[None, :]) + ws - 1].permute(2, 0, 1)       # [h,ws,ws]
        for a, b in zip(seqs.offsets[:-1], seqs.offsets[1:]):
            S = b - a
            if S == 0:
                continue
            Sp = (S + ws - 1) // ws * ws

            def prep(t, bias_row):
                p = torch.cat([t[a:b], bias_row[None].expand(Sp - S, -1)], 0)
                if shift:
                    p = torch.roll(p, -shift, 0)
                return p.reshape(Sp // ws, ws, heads, 64).permute(0, 2, 1, 3)
            Q, K, Vv = prep(qv, qb), prep(kv, kb), prep(vv, vb)
            sc = Q @ K.transpose(-1, -2) / 8.0 + bias[None]
            if shift:
                reg = (i >= ws // 2).float()
                m = (reg[:, None] != reg[None, :]).float() * -100.0
                sc[-1] = sc[-1] + m[None]
            o = (torch.softmax(sc, -1) @ Vv).permute(0, 2, 1, 3).reshape(Sp, heads * 64)
            if shift:
                o = torch.roll(o, shift, 0)
            out[a:b] = o[:S]

    def pair_concat(self, x, src, dst, y):
        xv, out = _v(x), _v(y)
        for (a, b), (c, d) in zip(zip(src.offsets[:-1], src.offsets[1:]), zip(dst.offsets[:-1], dst.offsets[1:])):
            s = xv[a:b]
            if (b - a) % 2:
                s = torch.cat([s, torch.zeros_like(s[:1])], 0)
            out[c:d] = torch.cat([s[0::2], s[1::2]], 1)

    def upsample_cols(self, x, src, dst, shift, y):
        xv, out = _v(x), _v(y)
        for (a, b), (c, d) in zip(zip(src.offsets[:-1], src.offsets[1:]), zip(dst.offsets[:-1], dst.offsets[1:])):
            j = torch.arange(d - c) >> shift
            out[c:d] = xv[a:b][j]

    def copy_cols(self, x, y, row_step=1, row_off=0, rows=None):
        out = _v(y)
        rows = out.shape[0] if rows is None else rows
        out[:rows] = _v(x)[row_off::row_step][:rows]

    def octattn_embed(self, ctx, ctx_pos, pos_scale, level_base, max_level, seqs, p, e, eu):
        c = ctx.reshape(-1, 4, 3).long()
        lvl = c[:, :, 0] - torch.clamp(c[:, 3:, 0] - level_base, min=0)
        lvl = lvl.clamp(0, max_level)
        occ = p["occ_enc.weight"][c[:, :, 2]]
        occ_u = occ.clone()
        occ_u[:, 3] = p["occ_enc.weight"][255]
        pos = ctx_pos.reshape(-1, 4, 3).float() * pos_scale
        rest = torch.cat([p["level_enc.weight"][lvl], p["octant_enc.weight"][c[:, :, 1]],
                          F.linear(pos, p["abs_pos_enc.weight"], p["abs_pos_enc.bias"])], 2)
        pe = p.get("transformer_encoder.position_enc.pe")          # absent with cfg.model.pos_embed False
        for a, b in zip(seqs.offsets[:-1], seqs.offsets[1:]):
            for o, dst in ((occ, e), (occ_u, eu)):
                dst[a:b] = torch.cat([o[a:b], rest[a:b]], 2).reshape(b - a, 600) * math.sqrt(600) + (pe[:b - a] if pe is not None else 0.0)

    def octattn_attention(self, qu, k, ku, v, vu, heads, hd, seqs, out, out_u):
        QU, K, KU, Vv, VU = (_v(t) for t in (qu, k, ku, v, vu))
        O, OU = _v(out), _v(out_u)
        for a, b in zip(seqs.offsets[:-1], seqs.offsets[1:]):
            S = b - a
            sp = lambda t: t[a:b].reshape(S, heads, hd).permute(1, 0, 2)
            q, kk, kku, vv, vvu = sp(QU), sp(K), sp(KU), sp(Vv), sp(VU)
            mask = torch.full((S, S), float("-inf")).triu(1)
            sc = q @ kk.transpose(1, 2) / math.sqrt(hd)
            att = torch.softmax(sc + mask, -1)
            O[a:b] = (att @ vv).permute(1, 0, 2).reshape(S, heads * hd)
            eye = torch.eye(S)
            scu = (1 - eye) * sc + torch.diag_embed((q * kku).sum(2) / math.sqrt(hd))
            attu = torch.softmax(scu + mask, -1)
            ou = ((1 - eye) * attu) @ vv + torch.diagonal(attu, dim1=1, dim2=2)[..., None] * vvu
            OU[a:b] = ou.permute(1, 0, 2).reshape(S, heads * hd)
