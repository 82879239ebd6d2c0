"""Pins the CPU oracle (oracle/octree_np.py) to outputs of the unmodified reference
(tests/golden/*.npz, produced by oracle/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

from oracle import octree_np as onp
from conftest import golden

CASES = ["k12s", "k14c", "k16m", "f17s", "k10c", "k14s"]
MPATHS = [[0, 0], [0, 1], [1]]


def oracle_rows(g):
    mode = str(g["mode"])
    pts = g["points"]
    if not bool(g["mullevel"]):
        qz = onp.quantize(pts[:, :3], float(g["qs"][0]), mode)
        return onp.tree_rows(qz["q"])["rows"], qz, None
    rows, subs = [], []
    for qs, mp in zip(g["qs"], MPATHS):
        qz = onp.quantize(pts[:, :3], float(qs), mode)
        r = onp.tree_rows(qz["q"], morton_path=mp, drop_last=True)["rows"]
        rows.append(r)
        subs.append(len(r))
    qz0 = onp.quantize(pts[:, :3], float(g["qs"][0]), mode)
    return np.vstack(rows), qz0, subs


@pytest.mark.parametrize("name", CASES)
def test_rows_match_reference(name):
    g = golden(f"octree_{name}.npz")
    rows, qz, subs = oracle_rows(g)
    assert rows.shape == g["rows"].shape
    assert np.array_equal(rows, g["rows"].astype(np.int64))
    assert qz["bin_num"] == float(g["bin_num"])
    if subs is not None:
        assert subs == list(g["sub_rows"])
    if "z_offset" in g.files:
        assert qz["offset"][2] == float(g["z_offset"])


@pytest.mark.parametrize("name", CASES)
def test_level_split_matches_reference(name):
    g = golden(f"octree_{name}.npz")
    mul = bool(g["mullevel"])
    lvl = int(g["level"])
    if not mul:
        ids, poss, pos_mm, data, oct_seq = onp.ehem_level_split(g["rows"].astype(np.int64), lvl)
    else:
        ids, poss, pos_mm, data, seqs = [], [], [], [], []
        s = 0
        for n in g["sub_rows"]:
            a = onp.ehem_level_split(g["rows"][s:s + n].astype(np.int64), lvl, mullevel=True)
            ids += a[0]; poss += a[1]; pos_mm += a[2]; data += a[3]; seqs.append(a[4])
            s += n
        oct_seq = np.vstack(seqs)
    assert [len(i) for i in ids] == list(g["level_sizes"])
    assert np.array_equal(np.concatenate(data, 0), g["ds_data"].astype(np.int64))
    got_pos = np.concatenate([p.T for p in poss], 0)
    assert got_pos.dtype == np.float32
    assert np.array_equal(got_pos, g["ds_pos"], equal_nan=True)
    assert np.array_equal(np.array(pos_mm, np.int64), g["ds_pos_mm"])
    assert np.array_equal(oct_seq, g["ds_oct_seq"].astype(np.int64))


@pytest.mark.parametrize("name", ["k12s", "k14s"])
def test_octattn_dataset_matches_reference(name):
    g = golden(f"octree_{name}.npz")
    ids, pos, data, _ = onp.octattn_dataset(g["rows"].astype(np.int64), 1024)
    assert np.array_equal(ids, g["oct_ids"])
    assert np.array_equal(data, g["oct_data"].astype(np.int64))
    assert np.array_equal(pos, g["oct_pos"])


def test_dequant_voxels_match_reference():
    """np.unique(axis=0) voxel list, checked through the reference's dequantised output."""
    g = golden("octree_k12s.npz")
    qz = onp.quantize(g["points"][:, :3], float(g["qs"][0]), "spher")
    vox = onp.voxels_unique(qz["q"])
    assert len(vox) == len(g["dequant"])
    assert onp.tree_rows(qz["q"])["n_voxels"] == len(vox)
