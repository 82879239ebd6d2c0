"""Host wrappers: coding order (A7), softmax->CDF (A13), range coder (A14).  See include/scp_b200.h."""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib


def coding_order(level_sizes, context_size, occ, mullevel=False, level_restart=None, reference_single_node_defect=None):
    """encode.py:109-136 / encode_mullevel.py:106-133.  occ: CUDA uint8 [N] occupancy bytes 1..255.
    Returns (order int64 [N], symbols int16 [N]) on the device.

    A level that holds a single node is coded on its own (encode.py:120-124).  encode_mullevel.py:120 adds the running
    node offset; encode.py:123 forgets it, so for a single-node level BELOW the root the reference's single-octree
    encoder codes the root's row a second time and never codes that node -- a stream no decoder can invert.  That
    defect is not reproduced: the node itself is coded (what encode_mullevel.py does) and ``Decoder`` reads it back.
    ``reference_single_node_defect=True`` (or SCP_REF_SINGLE_NODE_DEFECT=1) restores the reference's order for
    byte-level comparisons with its output on such frames."""
    lib = _lib.require_device()
    n = int(sum(level_sizes))
    order = torch.zeros(n, dtype=torch.int64, device=occ.device)
    sym = torch.zeros(n, dtype=torch.int16, device=occ.device)
    sizes = (C.c_int64 * len(level_sizes))(*[int(s) for s in level_sizes])
    restart = None if level_restart is None else np.ascontiguousarray(np.asarray(level_restart, np.uint8))
    if reference_single_node_defect is None:
        reference_single_node_defect = os.environ.get("SCP_REF_SINGLE_NODE_DEFECT", "0") == "1"
    add_base = 1 if (mullevel or not reference_single_node_defect) else 0
    _lib.check(lib.scp_coding_order(sizes, _lib.ptr(restart), len(level_sizes), context_size, add_base, _lib.ptr(occ),
                                    _lib.ptr(order), _lib.ptr(sym), _lib.stream_ptr()), "scp_coding_order")
    return order, sym


def pmf_to_cdf(x, sym=None, is_logits=True, row_of=None, n_out=None, want_cdf=False, want_interval=False,
               want_pmf=False, out=None):
    """x: CUDA float32 [n,255] logits or PMFs.  ``row_of`` (int64 [n]) scatters input row i to output row
    row_of[i] (outputs then have n_out rows).  Returns a dict with the requested outputs; ``out`` may carry
    preallocated output tensors to fill (for window-by-window scatter into frame-sized buffers)."""
    lib = _lib.require_device()
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[-1] == 255
    n = x.shape[0]
    n_out = n if n_out is None else n_out
    res = dict(out or {})
    if want_cdf and "cdf" not in res:
        res["cdf"] = torch.empty((n_out, 256), dtype=torch.uint16, device=x.device)
    if want_interval and "interval" not in res:
        res["interval"] = torch.empty((n_out, 2), dtype=torch.int32, device=x.device)
    if want_pmf and "pmf" not in res:
        res["pmf"] = torch.empty((n_out, 255), dtype=torch.float32, device=x.device)
    _lib.check(lib.scp_pmf_to_cdf(_lib.ptr(x), n, int(is_logits), _lib.ptr(row_of), _lib.ptr(sym),
                                  _lib.ptr(res.get("cdf")), _lib.ptr(res.get("interval")), _lib.ptr(res.get("pmf")),
                                  _lib.stream_ptr()), "scp_pmf_to_cdf")
    return res


def range_encode(interval):
    """interval: host array-like uint32/int32 [n,2] of (c_low, c_high) in coding order -> bytes
    (byte-identical to numpyAc_backend.encode_cdf on the same CDF/symbols)."""
    lib = _lib.load()
    iv = np.ascontiguousarray(np.asarray(interval).astype(np.uint32, copy=False))
    n = iv.shape[0]
    cap = 64 + 4 * n
    buf = np.empty(cap, np.uint8)
    got = lib.scp_range_encode(_lib.ptr(iv), n, _lib.ptr(buf), cap)
    if got == -1 and b"too small" in lib.scp_last_error():
        need = lib.scp_range_encode(_lib.ptr(iv), n, None, 0)
        buf = np.empty(need, np.uint8)
        got = lib.scp_range_encode(_lib.ptr(iv), n, _lib.ptr(buf), need)
    _lib.check(got, "scp_range_encode")
    return buf[:got].tobytes()


def range_encode_cdf(cdf, sym):
    """numpyAc_backend.encode_cdf(cdf int16/uint16 [N,Lp], sym int16 [N]) -> bytes."""
    lib = _lib.load()
    cdf = np.ascontiguousarray(np.asarray(cdf).view(np.uint16))
    sym = np.ascontiguousarray(np.asarray(sym, np.int16))
    n, Lp = cdf.shape
    need = lib.scp_range_encode_cdf(_lib.ptr(cdf), _lib.ptr(sym), n, Lp, None, 0)
    _lib.check(need, "scp_range_encode_cdf")
    buf = np.empty(max(need, 1), np.uint8)
    got = lib.scp_range_encode_cdf(_lib.ptr(cdf), _lib.ptr(sym), n, Lp, _lib.ptr(buf), need)
    _lib.check(got, "scp_range_encode_cdf")
    return buf[:got].tobytes()


class RangeDecoder:
    """numpyAc.arithmeticDeCoding (numpyAc.py:139-170) on the library's host decoder: ``decode(cdf)`` takes the uint16
    CDF rows [n,256] of the next n symbols in coding order and returns them (int16 [n])."""

    def __init__(self, bitstream: bytes):
        self.lib = _lib.load()
        buf = np.frombuffer(bytes(bitstream), np.uint8)
        self.h = self.lib.scp_range_decoder_create(_lib.ptr(np.ascontiguousarray(buf)), len(buf))
        if not self.h:
            raise RuntimeError(self.lib.scp_last_error().decode())

    def decode(self, cdf) -> np.ndarray:
        cdf = np.ascontiguousarray(np.asarray(cdf).view(np.uint16))
        if cdf.ndim != 2:
            raise ValueError("cdf must be [n, Lp]")
        n, Lp = cdf.shape
        sym = np.empty(n, np.int16)
        _lib.check(self.lib.scp_range_decode(self.h, _lib.ptr(cdf), n, Lp, _lib.ptr(sym)), "scp_range_decode")
        return sym

    @property
    def count(self):
        return int(self.lib.scp_range_decoder_count(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.scp_range_decoder_destroy(self.h)
                self.h = None
        except Exception:
            pass
