"""Oracle PMF->CDF against the reference's numpyAc output (golden), CPU only."""
import zlib

import numpy as np

from conftest import golden
from oracle import octree_np as onp
from oracle.make_golden import coder_case


def test_cdf_matches_numpyac():
    g = golden("coder.npz")
    pmf, _ = coder_case()
    cdf = onp.pmf_to_cdf_u16(pmf)
    assert np.array_equal(cdf[::50], g["cdf_rows"])
    assert zlib.crc32(cdf.view(np.int16).tobytes()) == int(g["cdf_crc"])


def test_coding_order_matches_reference_e2e():
    for name, mul in (("k12s", False), ("k16m", True)):
        g = golden(f"octree_{name}.npz")
        e = golden(f"e2e_{name}.npz")
        order = onp.coding_order(list(g["level_sizes"]), 8192, mullevel=mul)
        sym = g["ds_oct_seq"][:, -1, 0].astype(np.int16)[order]
        assert np.array_equal(sym, e["sym"])
