"""The reference-named drop-in modules (data_preproc, dataloaders, numpyAc, encode[_mullevel]) against the
goldens of the unmodified reference."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import golden
from test_models_cpu import cfg_ehem

pytestmark = pytest.mark.gpu


def write_bin(tmp_path, name):
    g = golden(f"octree_{name}.npz")
    d = tmp_path / "seq"
    d.mkdir(exist_ok=True)
    f = d / f"{name}.bin"
    g["points"].astype(np.float32).tofile(f)
    return g, str(f)


def test_proc_pc_writes_reference_npy(tmp_path):
    from scp_b200.data_preproc.data_preprocess import proc_pc, mul_proc_pc
    g, f = write_bin(tmp_path, "k12s")
    res = proc_pc(f, str(tmp_path), "o", qs=float(g["qs"][0]), test=True, spher=True)
    assert np.array_equal(np.load(res[0] + ".npy"), g["rows"].astype(np.int64))
    assert float(res[3]) == float(g["bin_num"])
    assert np.allclose(res[1], g["dequant"], rtol=0, atol=1e-4)
    g, f = write_bin(tmp_path, "k14c")
    res = proc_pc(f, str(tmp_path), "c", qs=float(g["qs"][0]), test=True, cylin=True)
    assert np.array_equal(np.load(res[0] + ".npy"), g["rows"].astype(np.int64))
    assert float(res[4][0, 2]) == float(g["z_offset"])
    g, f = write_bin(tmp_path, "k16m")
    rows = []
    for q, mp in zip(g["qs"], ([0, 0], [0, 1], [1])):
        res = mul_proc_pc(f, str(tmp_path), "m", qs=float(q), test=True, spher=True, morton_path=mp)
        assert res[0].endswith("m_" + "_".join(map(str, mp)))
        rows.append(np.load(res[0] + ".npy"))
    assert np.array_equal(np.vstack(rows), g["rows"].astype(np.int64))


def test_gen_octree_and_k_parent_seq(tmp_path):
    from scp_b200.data_preproc.OctreeCPP.Octreewarpper import gen_octree
    from scp_b200.data_preproc.Octree import gen_K_parent_seq
    from oracle import octree_np as onp
    g = golden("octree_k12s.npz")
    q = onp.voxels_unique(onp.quantize(g["points"][:, :3], float(g["qs"][0]), "spher")["q"])
    octree = gen_octree(q)
    assert len(octree) == 11 and octree[0].node[0].oct == g["rows"][0, 3, 0]
    with pytest.raises(IndexError):
        octree[len(octree)]
    s = gen_K_parent_seq(octree, 4)
    out = np.concatenate((s["Seq"][:, :, True], s["Level"], s["Pos"]), axis=2)
    assert np.array_equal(out, g["rows"].astype(np.int64))


@pytest.mark.parametrize("name,mul", [("k12s", False), ("k14c", False), ("k16m", True)])
def test_datasets_match_reference(tmp_path, name, mul):
    g, f = write_bin(tmp_path, name)
    mode = str(g["mode"])
    if mul:
        from scp_b200.dataloaders.encode_dataset_ehem_mullevel import EncodeEHEMDataset
        ds = EncodeEHEMDataset([f], 8192, "kitti", True, int(g["level"]), mode == "cylin", mode == "spher", "")
    else:
        from scp_b200.dataloaders.encode_dataset_ehem import EncodeEHEMDataset
        ds = EncodeEHEMDataset([f], 8192, "kitti", True, int(g["level"]), mode == "cylin", mode == "spher", False, False, "")
    ids, poss, pos_mm, data, oct_seq, n, pc, bin_num, z_off, _, _ = ds[0]
    assert [len(i) for i in ids] == list(g["level_sizes"]) and n == len(g["points"])
    assert data[0].dtype == np.int64 and poss[0].dtype == np.float32 and poss[-1].shape[0] == 3
    assert np.array_equal(np.concatenate(data), g["ds_data"].astype(np.int64))
    assert np.array_equal(np.concatenate([p.T for p in poss]), g["ds_pos"], equal_nan=True)
    assert np.array_equal(np.array(pos_mm, np.int64), g["ds_pos_mm"])
    assert np.array_equal(oct_seq, g["ds_oct_seq"].astype(np.int64))
    assert bin_num == int(g["bin_num"])


def test_octattn_dataset_matches_reference(tmp_path):
    from scp_b200.dataloaders.encode_dataset import EncodeDataset
    g, f = write_bin(tmp_path, "k12s")
    ids, pos, data, oct_seq, n, bin_num, _, _ = EncodeDataset([f], 1024, "kitti", False, 12, True, "")[0]
    assert np.array_equal(ids[0], g["oct_ids"]) and np.array_equal(data[0], g["oct_data"].astype(np.int64))
    assert np.array_equal(pos[0], g["oct_pos"])


def test_numpyac_dropin_is_byte_identical_and_raises_like_the_reference():
    from scp_b200 import numpyAc
    from oracle.make_golden import coder_case
    g = golden("coder.npz")
    pmf, sym = coder_case()
    bs, bits = numpyAc.arithmeticCoding().encode(pmf, sym)
    assert bits == int(g["bits"]) and np.array_equal(np.frombuffer(bs, np.uint8), g["bitstream"])
    with pytest.raises(ValueError):
        numpyAc.arithmeticCoding().encode(pmf, sym.astype(np.int32))
    bad = sym.copy(); bad[3] = 255
    with pytest.raises(ValueError):
        numpyAc.arithmeticCoding().encode(pmf, bad)
    with pytest.raises(AssertionError):
        numpyAc.arithmeticCoding().encode(pmf[:-1], sym)


@pytest.mark.parametrize("name,mul", [("k12s", False), ("k16m", True)])
def test_compress_ehem_dropin_writes_reference_sized_stream(tmp_path, name, mul):
    from scp_b200.models import EHEM
    from scp_b200 import encode, encode_mullevel
    g = golden(f"octree_{name}.npz")
    e = golden(f"e2e_{name}.npz")
    sizes = np.cumsum(np.concatenate([[0], g["level_sizes"]]))
    ids = [torch.arange(n)[None] for n in g["level_sizes"]]
    pos = [torch.from_numpy(g["ds_pos"][a:b].T.copy())[None] for a, b in zip(sizes[:-1], sizes[1:])]
    data = [torch.from_numpy(g["ds_data"][a:b].astype(np.int64))[None] for a, b in zip(sizes[:-1], sizes[1:])]
    oct_seq = torch.from_numpy(g["ds_oct_seq"].astype(np.int64))[None]
    batch = (ids, pos, [tuple(m) for m in g["ds_pos_mm"]], data, oct_seq, torch.tensor(len(g["points"])), None,
             torch.tensor(int(g["bin_num"])), torch.tensor(0))
    model = EHEM(cfg_ehem()).cuda()
    args = types.SimpleNamespace(spher=True, cylin=False, lidar_level=int(g["level"]))
    mod = encode_mullevel if mul else encode
    bpp, _ = mod.compress_ehem(batch, str(tmp_path / "out" / name), model, args)
    fn = str(tmp_path / "out" / name) + f"_spher_{len(data)}_{int(g['bin_num'])}_0.bin"
    assert os.path.exists(fn) and os.path.exists(fn + ".dat")
    assert abs(os.path.getsize(fn) - len(e["bitstream"])) / len(e["bitstream"]) < 0.005
    assert abs(bpp - float(e["bpp"])) / float(e["bpp"]) < 0.005
    # ... and the decode drop-in reads the files back to the reference's own occupancy sequence (decode_ehem.py:184)
    from scp_b200 import decode_ehem, decode_ehem_mullevel
    label = g["ds_oct_seq"][:, -1, 0] + 1          # the dataset already subtracted 1 (encode_dataset_ehem.py:54); the .npy holds 1..255
    if mul:
        cuts = np.cumsum(np.concatenate([[0], g["sub_rows"]]))
        code, bn, zo, _, spher, cylin = decode_ehem_mullevel.decodeOct(fn, [label[a:b] for a, b in zip(cuts[:-1], cuts[1:])], model)
    else:
        code, bn, zo, _, spher, cylin = decode_ehem.decodeOct(fn, label[:, None], model)
    assert np.array_equal(np.asarray(code) + 1, label) and bn == int(g["bin_num"]) and spher and not cylin


@pytest.mark.parametrize("name,mul", [("k12s", False), ("k16m", True)])
def test_encode_and_decode_command_lines(tmp_path, monkeypatch, name, mul):
    """``python -m scp_b200.encode[_mullevel]`` then ``python -m scp_b200.decode_ehem[_mullevel]`` on a sweep file: the
    decoder finds the stream by name, checks it against the pre-generated rows and writes the reconstructed cloud."""
    from scp_b200 import decode_ehem, decode_ehem_mullevel, encode, encode_mullevel
    from scp_b200.data_preproc import pt, test_gene
    from oracle import metrics_np as om
    monkeypatch.chdir(tmp_path)                               # encode.main appends test_results_*.txt to the cwd
    g, f = write_bin(tmp_path, name)
    m = golden("metrics.npz")
    level = str(int(g["level"]))
    pre, out = str(tmp_path / "pre"), str(tmp_path / "out")
    test_gene.main(test_gene.get_args(["--ori_dir", f, "--out_dir", pre, "--lidar_level", level, "--spher"] + (["--mullevel"] if mul else [])))
    torch.manual_seed(0)
    bpps = (encode_mullevel if mul else encode).main(encode.get_args(["--test_files", f, "--out_dir", out, "--lidar_level", level, "--spher"]))
    assert len(bpps) == 1 and bpps[0] > 0
    report = open(tmp_path / f"test_results_{'mul' if mul else 'same'}_kitti_{level}.txt").read().splitlines()
    assert any(l.startswith("chamfer_dist: ") for l in report) and any(l.startswith("PSNR: ") for l in report)
    torch.manual_seed(0)                                      # same random-init weights as the encoder
    dec = decode_ehem_mullevel if mul else decode_ehem
    written = dec.main(dec.get_args(["--test_files", f, "--out_dir", out, "--lidar_level", level, "--preproc_path", pre]))
    assert written == [out + f"/{name}.ply"]
    rec, want = pt.loadply(written[0])[0].astype(np.float64), m[name + "_q"].astype(np.float64)
    tol = 1e-2 if mul else 2e-4                               # mullevel: bin_num of the finer sub-octrees is re-derived
    assert om.nn_dist(rec, want).max() < tol                  # every reconstructed point is a voxel of the encoder
    # the mullevel stream does not code the last node of each sub-octree (Octree.py:259-262 drops that row), so its
    # children (here one voxel per sub-octree) cannot be rebuilt; the single-level stream is complete
    lost = int((om.nn_dist(want, rec) >= tol).sum())
    assert len(want) - len(rec) == lost and lost <= (3 if mul else 0)
