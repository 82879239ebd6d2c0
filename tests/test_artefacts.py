"""Artefact formats either side of the encode path: the row / meta / PLY files of
data_preproc/test_gene.py and the ``preproc_path`` ingest of the datasets."""
import numpy as np
import pytest

from conftest import golden

PLY_TEXT = """ply
format ascii 1.0
element vertex 2
property float x
property float y
property float z
property uint16 intensity
end_header
6.286511 -6.605243 32.021133 0
-116.251539 -10.939583 -62.295547 4
"""


def test_write_ply_data_format_and_read_back(tmp_path):
    """Byte-for-byte what the reference's np.savetxt-based writer produces (pt.py:114-151; text captured from it)."""
    from scp_b200.data_preproc import pt
    a = np.array([[6.2865114, -6.6052433, 32.0211332, 0], [-116.2515391, -10.9395829, -62.2955471, 4]])
    f = tmp_path / "sub" / "a.ply"
    pt.write_ply_data(f, a, ["intensity"], ["uint16"])
    assert open(f).read() == PLY_TEXT
    r = np.random.default_rng(0)
    p = r.standard_normal((70001, 3)) * 50                       # more than one formatting block
    pt.write_ply_data(str(tmp_path / "b.ply"), p)
    xyz, rest = pt.loadply(str(tmp_path / "b.ply"))
    assert xyz.dtype == np.float32 and xyz.shape == p.shape and rest.shape[1] == 0
    assert np.abs(xyz - p).max() < 1e-5
    assert np.array_equal(pt.ptread(str(tmp_path / "b.ply")), xyz)


@pytest.mark.parametrize("name", ["k12s", "k14c", "f17s"])
def test_levels_from_rows_single_level(name):
    """rows file -> the dataset tuple of the reference (golden ``ds_*`` arrays come from its ``__getitem__``)."""
    from scp_b200.dataloaders.encode_dataset_ehem import levels_from_rows
    g = golden(f"octree_{name}.npz")
    ids, poss, pos_mm, dat, seq = levels_from_rows(g["rows"].astype(np.int64), int(g["level"]))
    assert [len(i) for i in ids] == list(g["level_sizes"])
    assert np.array_equal(np.concatenate(dat, 0), g["ds_data"].astype(np.int64))
    assert np.array_equal(np.concatenate([p.T for p in poss], 0), g["ds_pos"])
    assert np.array_equal(np.array(pos_mm, np.int64), g["ds_pos_mm"])
    assert np.array_equal(seq, g["ds_oct_seq"].astype(np.int64))


def test_levels_from_rows_mullevel():
    from scp_b200.dataloaders.encode_dataset_ehem import levels_from_rows
    g = golden("octree_k16m.npz")
    rows, cuts = g["rows"].astype(np.int64), np.concatenate([[0], np.cumsum(g["sub_rows"])])
    parts = [levels_from_rows(rows[a:e], int(g["level"]), eps_last=False) for a, e in zip(cuts[:-1], cuts[1:])]
    assert sum(([len(i) for i in p[0]] for p in parts), []) == list(g["level_sizes"])
    assert np.array_equal(np.concatenate(sum((p[3] for p in parts), []), 0), g["ds_data"].astype(np.int64))
    assert np.array_equal(np.concatenate([x.T for p in parts for x in p[1]], 0), g["ds_pos"])
    assert np.array_equal(np.array(sum((p[2] for p in parts), []), np.int64), g["ds_pos_mm"])
    assert np.array_equal(np.vstack([p[4] for p in parts]), g["ds_oct_seq"].astype(np.int64))


def test_test_gene_arguments():
    from scp_b200.data_preproc import test_gene
    a = test_gene.get_args(["--ori_dir", "x/*.bin", "--out_dir", "o", "--spher", "--mullevel", "--parts", "1/4"])
    assert (a.type, a.lidar_level, a.spher, a.cylin, a.mullevel, a.parts) == ("kitti", 16, True, False, True, "1/4")
    assert test_gene._qs(a) == 400 / (2 ** 16 - 1) and test_gene._qs(a, 2) == 400 / (2 ** 18 - 1)
    a.type = "ford"
    assert test_gene._qs(a, 1) == 2 and test_gene._names("d/seq/f.ply", a)[1] == "f"
    a.type = "kitti"
    assert test_gene._names("d/seq/f.bin", a)[1] == "seqf"


# ---------------------------------------------------------------------------------------------- GPU
def _same(a, b):
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode,mul", [("k12s", "spher", False), ("k14c", "cylin", False), ("k16m", "spher", True)])
def test_gene_artefacts_feed_the_datasets(tmp_path, name, mode, mul):
    """test_gene writes rows / _loc / _quant.ply / _meta; a dataset given ``preproc_path`` returns the same tuple as one that
    runs the CUDA pre-processing itself (PSNR excepted: the artefacts do not store it)."""
    from scp_b200.data_preproc import pt, test_gene
    from scp_b200.dataloaders.encode_dataset import EncodeDataset
    from scp_b200.dataloaders.encode_dataset_ehem import EncodeEHEMDataset
    from scp_b200.dataloaders.encode_dataset_ehem_mullevel import EncodeEHEMDataset as MulDataset
    g, m = golden(f"octree_{name}.npz"), golden("metrics.npz")
    seq = tmp_path / "seq"
    seq.mkdir()
    pts = g["points"][:, :3].astype(np.float32)
    f = str(seq / f"{name}.bin")
    np.hstack([pts, np.zeros((len(pts), 1), np.float32)]).tofile(f)
    out = str(tmp_path / "pre")
    argv = ["--type", "kitti", "--ori_dir", str(seq / "*.bin"), "--out_dir", out, "--lidar_level", str(int(g["level"])),
            "--" + mode] + (["--mullevel"] if mul else [])
    test_gene.main(test_gene.get_args(argv))
    base = out + "/seq" + name
    files = [base + s for s in ("_0_0", "_0_1", "_1")] if mul else [base]
    assert np.array_equal(np.vstack([np.load(x + ".npy") for x in files]), g["rows"].astype(np.int64))
    assert all(np.array_equal(np.load(x + "_loc.npy"), pts) for x in files)
    meta = np.load(base + "_meta.npy")
    assert len(meta) == (2 if mode == "spher" and not mul else 3) and meta[0] == float(g["bin_num"])
    assert meta[1] == pytest.approx(float(m[name + "_chamfer"]), rel=1e-9 if mul else 1e-3)
    q = pt.loadply(base + "_quant.ply")[0]
    assert q.shape == m[name + "_q"].shape and np.abs(q - m[name + "_q"]).max() < 2e-6 * max(1, np.abs(q).max()) + 1e-6

    level, cyl, sph = int(g["level"]), mode == "cylin", mode == "spher"
    if mul:
        mk = lambda pre: MulDataset([f], 8192, "kitti", True, level, cyl, sph, pre)
    else:
        mk = lambda pre: EncodeEHEMDataset([f], 8192, "kitti", True, level, cyl, sph, False, False, pre)
    direct, pre = mk("")[0], mk(out + "/")[0]
    assert len(direct) == len(pre) == 11
    for i in range(9):                                            # ids, poss, pos_mm, data, oct_seq, n, pc, bin_num, z_offset
        assert _same(direct[i], pre[i]), i
    assert pre[9] == meta[1] and direct[9] == pytest.approx(meta[1], rel=1e-3) and pre[10] == 0
    if mode == "spher" and not mul:
        d, p = EncodeDataset([f], 1024, "kitti", False, level, True, "")[0], EncodeDataset([f], 1024, "kitti", False, level, True, out + "/")[0]
        assert all(_same(d[i], p[i]) for i in range(6)) and p[6] == meta[1] and p[7] == 0


@pytest.mark.parametrize("name,mul", [("k12s", False), ("k16m", True)])
def test_reconstruct_from_header_fields(name, mul):
    """decode side (host logic, no GPU): voxels + the header fields of the file name -> the reference's quantised cloud.
    The mullevel sub-octrees re-derive their own bin_num from the first one (see ``reconstruct``), hence the looser bound."""
    import types
    from oracle import metrics_np as om
    from oracle import octree_np as onp
    from scp_b200.decode_ehem import reconstruct
    g, m = golden(f"octree_{name}.npz"), golden("metrics.npz")
    pts, level = g["points"][:, :3], int(g["level"])
    vox = []
    for qs, mp in zip(g["qs"], ([0, 0], [0, 1], [1]) if mul else (None,)):
        q = np.asarray(onp.quantize(pts, float(qs), "spher")["q"], np.int64)
        if mp is not None:
            n = onp.depth_of(q)
            keep = np.ones(len(q), bool)
            for j, bit in enumerate(mp):
                keep &= ((q[:, 0] >> (n - 1 - j)) & 1) == bit
            q = q[keep]
        vox.append(onp.voxels_unique(q))
    rec = reconstruct(types.SimpleNamespace(voxels=vox), int(g["bin_num"]), 0, True, False, level, "kitti")
    want = m[name + "_q"].astype(np.float64)
    assert rec.shape == want.shape and rec.dtype == np.float64
    tol = 1e-2 if mul else 2e-4
    assert om.nn_dist(rec, want).max() < tol and om.nn_dist(want, rec).max() < tol
