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


def test_coding_order_matches_oracle():
    from scp_b200 import coder
    for sizes, mul in (([1, 4, 9, 8192, 8193, 20000, 1, 3], False), ([1, 1, 2, 5], True)):
        occ = np.random.default_rng(1).integers(1, 256, sum(sizes)).astype(np.uint8)
        order, sym = coder.coding_order(sizes, 8192, torch.from_numpy(occ).cuda(), mullevel=mul)
        exp = onp.coding_order(sizes, 8192, mullevel=mul)
        got = order.cpu().numpy()
        # the reference's single-node quirk can make positions collide; compare the defined entries
        assert np.array_equal(got, exp)
        assert np.array_equal(sym.cpu().numpy(), occ[exp].astype(np.int16) - 1)
