"""GPU parity of coding order (A7) and PMF->CDF (A13) vs the reference goldens / oracle."""
import zlib

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import octree_np as onp
from oracle.make_golden import coder_case

pytestmark = pytest.mark.gpu


def test_cdf_from_pmf_bit_exact_and_bitstream_identical():
    from scp_b200 import coder
    g = golden("coder.npz")
    pmf, sym = coder_case()
    res = coder.pmf_to_cdf(torch.from_numpy(pmf).cuda(), sym=torch.from_numpy(sym).cuda(), is_logits=False,
                           want_cdf=True, want_interval=True)
    cdf = res["cdf"].cpu().numpy()
    assert zlib.crc32(cdf.view(np.int16).tobytes()) == int(g["cdf_crc"])
    assert np.array_equal(cdf[::50], g["cdf_rows"])
    iv = res["interval"].cpu().numpy()
    bs = coder.range_encode(iv)
    assert np.array_equal(np.frombuffer(bs, np.uint8), g["bitstream"])


def test_cdf_from_logits_close_to_oracle_and_ragged_sizes():
    from scp_b200 import coder
    rng = np.random.default_rng(3)
    for n in (1, 63, 64, 65, 1000):
        logits = (rng.normal(0, 4, (n, 255))).astype(np.float32)
        res = coder.pmf_to_cdf(torch.from_numpy(logits).cuda(), is_logits=True, want_cdf=True, want_pmf=True)
        pm = res["pmf"].cpu().numpy()
        ref = onp.softmax_f32(logits)
        assert np.abs(pm - ref).max() < 1e-6
        # CDF from OUR pmf must be exactly the numpyAc CDF of that pmf
        assert np.array_equal(res["cdf"].cpu().numpy(), onp.pmf_to_cdf_u16(pm))


def test_interval_only_path_equals_full_cdf_path():
    """The encoder asks for (c_low, c_high) only (register cumsum sampled at the symbol); it must give exactly the entries of
    the full CDF table, for logits and for PMFs, with and without a scatter map."""
    from scp_b200 import coder
    rng = np.random.default_rng(5)
    for n in (1, 64, 257, 5000):
        logits = torch.from_numpy(rng.normal(0, 5, (n, 255)).astype(np.float32)).cuda()
        sym = torch.from_numpy(rng.integers(0, 255, n).astype(np.int16)).cuda()
        sym[: min(n, 3)] = torch.tensor([0, 254, 127][: min(n, 3)], dtype=torch.int16)
        perm = torch.from_numpy(rng.permutation(n)).cuda()
        for is_logits in (True, False):
            x = logits if is_logits else torch.softmax(logits, 1).contiguous()
            full = coder.pmf_to_cdf(x, sym=sym, is_logits=is_logits, want_cdf=True, want_interval=True)
            only = coder.pmf_to_cdf(x, sym=sym, is_logits=is_logits, want_interval=True)
            assert torch.equal(full["interval"], only["interval"])
            cdf = full["cdf"].cpu().numpy().astype(np.int64)
            s = sym.cpu().numpy().astype(np.int64)
            lo, hi = cdf[np.arange(n), s], np.where(s == 254, 0x10000, cdf[np.arange(n), np.minimum(s + 1, 255)])
            assert np.array_equal(only["interval"].cpu().numpy().view(np.uint32), np.stack([lo, hi], 1).astype(np.uint32))
            scat = coder.pmf_to_cdf(x, sym=sym, is_logits=is_logits, row_of=perm, out={"interval": torch.zeros((n, 2), dtype=torch.int32, device="cuda")})
            # row i of the input lands in row perm[i], coded with sym[perm[i]]
            again = coder.pmf_to_cdf(x, sym=sym[perm].contiguous(), is_logits=is_logits, want_interval=True)["interval"]
            assert torch.equal(scat["interval"][perm], again)


def test_coding_order_matches_oracle():
    from scp_b200 import coder
    for sizes, mul in (([1, 4, 9, 8192, 8193, 20000, 1, 3], False), ([1, 1, 2, 5], True)):
        occ = np.random.default_rng(1).integers(1, 256, sum(sizes)).astype(np.uint8)
        # the reference's order, including the single-node quirk of encode.py:123 (a one-node level below the root codes
        # row 0 again), on request ...
        order, sym = coder.coding_order(sizes, 8192, torch.from_numpy(occ).cuda(), mullevel=mul, reference_single_node_defect=True)
        exp = onp.coding_order(sizes, 8192, mullevel=mul)
        assert np.array_equal(order.cpu().numpy(), exp)
        assert np.array_equal(sym.cpu().numpy(), occ[exp].astype(np.int16) - 1)
        # ... and the default: every node coded exactly once (what encode_mullevel.py:120 does), a permutation
        order, sym = coder.coding_order(sizes, 8192, torch.from_numpy(occ).cuda(), mullevel=mul)
        exp = onp.coding_order(sizes, 8192, mullevel=True)
        assert np.array_equal(order.cpu().numpy(), exp) and sorted(exp.tolist()) == list(range(sum(sizes)))
        assert np.array_equal(sym.cpu().numpy(), occ[exp].astype(np.int16) - 1)
