"""Entropy-model operators: thin torch-tensor wrappers over the C ABI (include/scp_b200.h, A8-A12).

``CudaOps`` is the only backend the product uses; it raises if no sm_100 device / library is present.
Tensors are views into caller-owned buffers: every op takes (tensor, column offset, columns) style
arguments through ``View`` so that concatenations are written in place instead of copied."""
import ctypes as C
import os
import weakref

import torch

from . import _lib

ACT = {"none": 0, "leaky": 1, "gelu": 2, "relu": 3}
ENGINE = {"auto": 0, "simt": 1, "tf32": 2, "tf32x3": 3, "f16x3": 4}


class Seqs:
    """Ragged batch description (host offsets + device tables).  ``half()`` gives the patch-merged
    (ceil(S/2)) sequences of swin_transformer.py:756."""

    def __init__(self, offsets, ops):
        self.offsets = [int(o) for o in offsets]
        self.ops = ops
        self.handle = ops._seqs_create(self.offsets)
        self._half = None

    @property
    def total(self):
        return self.offsets[-1] - self.offsets[0]

    @property
    def lengths(self):
        return [b - a for a, b in zip(self.offsets[:-1], self.offsets[1:])]

    def half(self):
        if self._half is None:
            offs = [0]
            for n in self.lengths:
                offs.append(offs[-1] + (n + 1) // 2)
            self._half = Seqs(offs, self.ops)
        return self._half

    def __del__(self):
        try:
            self.ops._seqs_destroy(self.handle)
        except Exception:
            pass


def V(t, col=0, ncol=None):
    """(tensor, column offset, columns) view of a 2-D row-major tensor."""
    return (t, col, t.shape[1] - col if ncol is None else ncol)


class _Rec:
    def __init__(self, ops, tag, flops, nbytes):
        self.ops, self.tag, self.flops, self.nbytes = ops, tag, flops, nbytes

    def __enter__(self):
        if self.ops.prof is not None:
            pool = self.ops.event_pool
            self.e0 = pool.pop() if pool else torch.cuda.Event(enable_timing=True)
            self.e1 = pool.pop() if pool else torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.ops.prof is not None:
            self.e1.record()
            self.ops.prof.append((self.tag, self.flops, self.nbytes, self.e0, self.e1))


# The library caches the hi/lo splits of every weight matrix under its device pointer and cannot see contents.  This table
# ties each cached pointer to the tensor OBJECT and in-place version it was split from; ``CudaOps.linear`` drops the cache
# entry when either changed (a new tensor at a recycled address, ``load_state_dict`` copying into the same storage).
_WEIGHT_SEEN = {}


def _weight_is_current(lib, w):
    key = w.data_ptr()
    ent = _WEIGHT_SEEN.get(key)
    if ent is not None and ent[0]() is w and ent[1] == w._version:
        return
    lib.scp_gemm_cache_drop(C.c_void_p(key))
    if len(_WEIGHT_SEEN) > 4096:                                   # forget dead tensors
        for k in [k for k, (r, _) in _WEIGHT_SEEN.items() if r() is None]:
            del _WEIGHT_SEEN[k]
    _WEIGHT_SEEN[key] = (weakref.ref(w), w._version)


class CudaOps:
    name = "cuda"

    def __init__(self, engine=None):
        self.lib = _lib.require_device()
        self.engine = ENGINE[engine or os.environ.get("SCP_GEMM", "auto")]
        # 3xFP16 engine: activations are converted unscaled (cvt.rn.satfinite), so |x| must stay below 65504 (weights carry
        # their own power-of-two scale).  SCP_CHECK_RANGE=1 verifies every Linear input (one reduction + sync per call)
        self.check_range = os.environ.get("SCP_CHECK_RANGE", "0") == "1"
        self.prof = None          # bench.py: list of (kernel tag, flops, bytes, start event, end event)
        self.event_pool = []      # pre-created timing events (creating them inside the timed region is slow)

    def reserve_events(self, n):
        self.event_pool = [torch.cuda.Event(enable_timing=True) for _ in range(n)]

    def _rec(self, tag, flops, nbytes):
        """Context manager that brackets one launch with CUDA events on the current stream when profiling is on."""
        return _Rec(self, tag, flops, nbytes)

    # -- plumbing ---------------------------------------------------------------------------
    def _seqs_create(self, offsets):
        arr = (C.c_int64 * len(offsets))(*offsets)
        h = self.lib.scp_seqs_create_async(arr, len(offsets) - 1, _lib.stream_ptr())
        if not h:
            raise _lib.ScpError("scp_seqs_create: " + self.lib.scp_last_error().decode())
        return h

    def _seqs_destroy(self, h):
        if h:
            self.lib.scp_seqs_destroy(h)

    def seqs(self, offsets):
        return Seqs(offsets, self)

    @staticmethod
    def _p(view):
        t, col, _ = view
        assert t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1
        return C.c_void_p(t.data_ptr() + 4 * col), t.stride(0)

    def empty(self, rows, cols, like):
        return torch.empty((rows, cols), dtype=torch.float32, device=like.device)

    # -- operators --------------------------------------------------------------------------
    def linear(self, x, w, b, y, act="none", res=None, row_step=1, row_off=0, rows=None, engine=None):
        """y = act(x @ w.T + b) (+ res).  x, y, res are views; ``row_step/row_off`` read every
        row_step-th row of x (even/odd token split of ehem.py:113-114)."""
        xp, ldx = self._p(x)
        yp, ldy = self._p(y)
        M = y[0].shape[0] if rows is None else rows
        if row_step != 1 or row_off:
            xp = C.c_void_p(xp.value + 4 * ldx * row_off)
            ldx = ldx * row_step
        rp, ldr = (None, 0) if res is None else self._p(res)
        N, K = w.shape
        assert x[2] == K and y[2] == N, (x[2], K, y[2], N)
        _weight_is_current(self.lib, w)
        if self.check_range:
            t, col, n = x
            m = float(t[row_off::row_step, col:col + n][:M].abs().max()) if M else 0.0
            if not m < 6.0e4:
                raise _lib.ScpError(f"scp_linear: activation magnitude {m:.3g} exceeds the fp16 range of the 3xFP16 engine "
                                    f"(set SCP_AUTO_ENGINE=1 for 3xTF32)")
        with self._rec("linear", 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N * (2 if res is not None else 1))):
            _lib.check(self.lib.scp_linear(xp, ldx, _lib.ptr(w), _lib.ptr(b), rp, ldr, yp, ldy, M, N, K, ACT[act],
                                           self.engine if engine is None else ENGINE[engine], _lib.stream_ptr()),
                       "scp_linear")

    def layernorm(self, x, g, b, y, res=None, eps=1e-5):
        xp, ldx = self._p(x)
        yp, ldy = self._p(y)
        rp, ldr = (None, 0) if res is None else self._p(res)
        with self._rec("layernorm", 8.0 * x[0].shape[0] * x[2], 4.0 * x[0].shape[0] * x[2] * (3 if res is not None else 2)):
            _lib.check(self.lib.scp_layernorm(xp, ldx, rp, ldr, _lib.ptr(g), _lib.ptr(b), yp, ldy, x[0].shape[0], x[2],
                                              eps, _lib.stream_ptr()), "scp_layernorm")

    def ehem_embed(self, ctx, occ_enc, level_enc, octant_enc, y):
        yp, ldy = self._p(y)
        _lib.check(self.lib.scp_ehem_embed(_lib.ptr(ctx), ctx.shape[0], _lib.ptr(occ_enc), _lib.ptr(level_enc),
                                           level_enc.shape[0], _lib.ptr(octant_enc), yp, ldy, _lib.stream_ptr()),
                   "scp_ehem_embed")

    def ehem_embed_occ(self, ctx, occ_enc, y):
        yp, ldy = self._p(y)
        _lib.check(self.lib.scp_ehem_embed_occ(_lib.ptr(ctx), ctx.shape[0] // 2, _lib.ptr(occ_enc), yp, ldy,
                                               _lib.stream_ptr()), "scp_ehem_embed_occ")

    def knn(self, x, seqs, k):
        xp, ldx = self._p(x)
        idx = torch.empty((x[0].shape[0], k), dtype=torch.int32, device=x[0].device)
        fl = sum(2.0 * n * n * x[2] for n in seqs.lengths)
        with self._rec("knn_d%d" % x[2], fl, 4.0 * x[0].shape[0] * (x[2] + k)):
            _lib.check(self.lib.scp_knn(xp, ldx, x[2], seqs.handle, k, _lib.ptr(idx), _lib.stream_ptr()), "scp_knn")
        return idx

    def edge_gather_max(self, uv, C_, idx, bn_scale, bn_shift, y, y2=None):
        """``y2``: optional second view that receives the same values (the next layer's concatenated input)."""
        up, ldu = self._p(uv)
        yp, ldy = self._p(y)
        y2p, ldy2 = (None, 0) if y2 is None else self._p(y2)
        n = idx.shape[0]
        with self._rec("edge_gather", 2.0 * n * C_ * idx.shape[1], 4.0 * n * (C_ * (idx.shape[1] + 2) + idx.shape[1])):
            _lib.check(self.lib.scp_edge_gather_max2(up, ldu, C_, _lib.ptr(idx), idx.shape[1], n,
                                                     _lib.ptr(bn_scale), _lib.ptr(bn_shift), yp, ldy, y2p, ldy2, _lib.stream_ptr()),
                       "scp_edge_gather_max")

    def swin_attention(self, q, k, v, qb, kb, vb, relpos, heads, seqs, shift, y):
        qp, ldq = self._p(q)
        kp, ldk = self._p(k)
        vp, ldv = self._p(v)
        yp, ldy = self._p(y)
        nwin = sum((n + 511) // 512 for n in seqs.lengths)
        with self._rec("swin_attention", 4.0 * nwin * heads * 512 * 512 * 64, 4.0 * 4 * seqs.total * heads * 64):
            _lib.check(self.lib.scp_swin_attention(qp, ldq, kp, ldk, vp, ldv, _lib.ptr(qb), _lib.ptr(kb), _lib.ptr(vb),
                                                   _lib.ptr(relpos), heads, seqs.handle, shift, yp, ldy,
                                                   _lib.stream_ptr()), "scp_swin_attention")

    def pair_concat(self, x, src, dst, y):
        xp, ldx = self._p(x)
        yp, ldy = self._p(y)
        _lib.check(self.lib.scp_pair_concat(xp, ldx, src.handle, dst.handle, x[2], yp, ldy, _lib.stream_ptr()),
                   "scp_pair_concat")

    def upsample_cols(self, x, src, dst, shift, y):
        xp, ldx = self._p(x)
        t, col, _ = y
        _lib.check(self.lib.scp_upsample_cols(xp, ldx, src.handle, dst.handle, shift, x[2], _lib.ptr(t), t.stride(0),
                                              col, _lib.stream_ptr()), "scp_upsample_cols")

    def copy_cols(self, x, y, row_step=1, row_off=0, rows=None):
        xp, ldx = self._p(x)
        t, col, _ = y
        rows = t.shape[0] if rows is None else rows
        _lib.check(self.lib.scp_copy_cols(xp, ldx, row_step, row_off, rows, x[2], _lib.ptr(t), t.stride(0), col,
                                          _lib.stream_ptr()), "scp_copy_cols")

    def octattn_embed(self, ctx, ctx_pos, pos_scale, level_base, max_level, seqs, p, e, eu):
        _lib.check(self.lib.scp_octattn_embed(_lib.ptr(ctx), _lib.ptr(ctx_pos), pos_scale, level_base, max_level,
                                              seqs.handle, _lib.ptr(p["occ_enc.weight"]), _lib.ptr(p["level_enc.weight"]),
                                              _lib.ptr(p["octant_enc.weight"]), _lib.ptr(p["abs_pos_enc.weight"]),
                                              _lib.ptr(p["abs_pos_enc.bias"]),
                                              _lib.ptr(p.get("transformer_encoder.position_enc.pe")), _lib.ptr(e),
                                              _lib.ptr(eu), _lib.stream_ptr()), "scp_octattn_embed")

    def octattn_attention(self, qu, k, ku, v, vu, heads, head_dim, seqs, out, out_u):
        p = [self._p(t) for t in (qu, k, ku, v, vu)]
        ld = p[0][1]
        assert all(x[1] == ld for x in p)
        op, ldo = self._p(out)
        oup, ldo2 = self._p(out_u)
        assert ldo == ldo2
        # algorithmic flops: causal QK^T and PV of the shared score matrix (the reference materialises the full S x S
        # scores and two PV products, attention_model.py:72-93)
        fl = sum(2.0 * n * n * heads * head_dim for n in seqs.lengths)
        with self._rec("octattn_attention", fl, 4.0 * 7 * seqs.total * heads * head_dim):
            self._octattn_attention(p, ld, heads, head_dim, seqs, op, oup, ldo)

    def _octattn_attention(self, p, ld, heads, head_dim, seqs, op, oup, ldo):
        _lib.check(self.lib.scp_octattn_attention(p[0][0], p[1][0], p[2][0], p[3][0], p[4][0], ld, heads, head_dim,
                                                  seqs.handle, op, oup, ldo, _lib.stream_ptr()),
                   "scp_octattn_attention")
