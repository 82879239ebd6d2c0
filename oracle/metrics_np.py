"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the distortion figures the reference prints next to bpp
(SURVEY.md section 8 row f-4).  Only tests/ may import this module; the product path is scp_b200/metrics.py (CUDA).

* ``dist_chamfer``  -- pt.py:88-95: nearest neighbours by scipy's KDTree in both directions, max of the mean distances.
  Pinned against the reference's own function on the golden frames (tests/golden/metrics.npz, oracle/make_golden.py).
* ``d1_psnr``       -- the symmetric point-to-point PSNR of MPEG's pc_error (mpeg-pcc-dmetric; the binary that
  pt.py:13-84 shells out to is NOT in the reference repository, so this figure is "parity unpinned": restated from the
  tool's published definition, psnr = 10 log10(3 peak^2 / max(mse_AB, mse_BA))).
* ``dequantise_keys`` -- Morton voxel keys -> points, data_preprocess.py:68-92 / :160-167 / :179-229.
"""
import math

import numpy as np
from scipy.spatial import cKDTree


def nn_dist(query, cand):
    return cKDTree(np.asarray(cand, np.float64)).query(np.asarray(query, np.float64), k=1)[0]


def dist_chamfer(f1, f2):
    return max(nn_dist(f2, f1).mean(), nn_dist(f1, f2).mean())


def d1_psnr(ref, deg, peak):
    mse = max((nn_dist(ref, deg) ** 2).mean(), (nn_dist(deg, ref) ** 2).mean())
    return mse, (math.inf if mse == 0 else 10 * math.log10(3 * peak * peak / mse))


def dequantise_keys(keys, steps, offset, mode):
    k = np.asarray(keys).astype(np.uint64)
    v = np.zeros((len(k), 3), np.int64)
    for b in range(21):
        for c in range(3):
            v[:, c] |= ((k >> np.uint64(3 * b + 2 - c)) & np.uint64(1)).astype(np.int64) << b
    p = v * np.asarray(steps, np.float64)[None] + np.asarray(offset, np.float64)[None]
    a, b, c = p[:, 0], p[:, 1], p[:, 2]
    if mode == "spher":
        return np.stack([a * np.sin(c) * np.cos(b), a * np.sin(c) * np.sin(b), a * np.cos(c)], 1)
    if mode == "cylin":
        return np.stack([a * np.cos(b), a * np.sin(b), c], 1)
    return p
