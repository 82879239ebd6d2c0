"""CPU-side checks of the boundary: the library builds/loads, exports every symbol include/scp_b200.h
declares, and the host-only entry points (range coder) match the reference's numpyAc goldens."""
import ctypes as C
import os
import re

import numpy as np

from conftest import ROOT, golden
from scp_b200 import _lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "scp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) > 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/scp_b200.h but not exported"
    # and the binding table covers the header
    missing = [s for s in syms if s not in _lib.SIGNATURES]
    assert not missing, missing


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.scp_version() >= 100
    assert lib.scp_range_encode(None, 1, None, 0) < 0
    assert b"bad argument" in lib.scp_last_error()


def test_range_coder_matches_numpyac_bitstream():
    g = golden("coder.npz")
    from oracle.make_golden import coder_case
    from oracle import octree_np as onp
    pmf, sym = coder_case()
    assert np.array_equal(sym, g["sym"])
    cdf = onp.pmf_to_cdf_u16(pmf)
    lib = _lib.load()
    cap = 1 << 20
    out = np.zeros(cap, np.uint8)
    n = lib.scp_range_encode_cdf(_lib.ptr(np.ascontiguousarray(cdf)), _lib.ptr(sym), len(sym), 256, _lib.ptr(out), cap)
    assert n == len(g["bitstream"])
    assert np.array_equal(out[:n], g["bitstream"])
    # interval form gives the same bytes
    lo = cdf[np.arange(len(sym)), sym].astype(np.uint32)
    hi = np.where(sym == 254, 0x10000, cdf[np.arange(len(sym)), np.minimum(sym + 1, 255)].astype(np.int64)).astype(np.uint32)
    iv = np.ascontiguousarray(np.stack([lo, hi], 1))
    out2 = np.zeros(cap, np.uint8)
    n2 = lib.scp_range_encode(_lib.ptr(iv), len(sym), _lib.ptr(out2), cap)
    assert n2 == n and np.array_equal(out2[:n], out[:n])
    # size-only call
    assert lib.scp_range_encode(_lib.ptr(iv), len(sym), None, 0) == n


def test_range_coder_edge_cases():
    lib = _lib.load()
    out = np.zeros(16, np.uint8)
    assert lib.scp_range_encode(_lib.ptr(np.zeros((0, 2), np.uint32)), 0, _lib.ptr(out), 16) == 1   # empty stream: 1 pending bit
    sym = np.array([300], np.int16)
    cdf = np.zeros((1, 256), np.uint16)
    assert lib.scp_range_encode_cdf(_lib.ptr(cdf), _lib.ptr(sym), 1, 256, _lib.ptr(out), 16) < 0


def test_range_decoder_inverts_numpyac_bitstream():
    """The reference's own bitstream (numpyAc.encode on the golden case) decodes back to the golden symbols, in one call
    and in ragged pieces (the decoder is stateful: the entropy model hands over CDF rows a window at a time)."""
    from oracle.make_golden import coder_case
    from oracle import octree_np as onp
    from scp_b200 import coder
    g = golden("coder.npz")
    pmf, sym = coder_case()
    cdf = onp.pmf_to_cdf_u16(pmf)
    d = coder.RangeDecoder(g["bitstream"].tobytes())
    assert np.array_equal(d.decode(cdf), sym)
    d = coder.RangeDecoder(g["bitstream"].tobytes())
    got, i = [], 0
    for step in (1, 7, 500, 3, 10 ** 6):
        got.append(d.decode(cdf[i:i + step]))
        i += step
    assert np.array_equal(np.concatenate(got), sym) and d.count == len(sym)


def test_range_coder_round_trip_skewed_and_wrapped_cdfs():
    """Peaked PMFs (CDF entries above 0x7fff wrap negative as int16, numpyAc.py:104), symbol 254 (c_high = 0x10000) and
    long runs of near-certain symbols (E3 pending bits)."""
    from oracle import octree_np as onp
    from scp_b200 import coder
    rng = np.random.default_rng(5)
    n = 20000
    pmf = rng.random((n, 255)).astype(np.float32) ** 8 + 1e-6
    hot = rng.integers(0, 255, n)
    hot[:200] = 254
    pmf[np.arange(n), hot] += rng.choice([0.5, 50.0, 5000.0], n).astype(np.float32)
    pmf /= pmf.sum(1, keepdims=True)
    sym = np.where(rng.random(n) < 0.9, hot, rng.integers(0, 255, n)).astype(np.int16)
    cdf = onp.pmf_to_cdf_u16(pmf)
    bs = coder.range_encode_cdf(cdf, sym)
    assert np.array_equal(coder.RangeDecoder(bs).decode(cdf), sym)
