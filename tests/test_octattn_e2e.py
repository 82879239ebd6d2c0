"""SCP-OctAttention end to end (BASELINE.json configs[3]) against the reference's OWN end-to-end runs
(tests/golden/e2e_octattn_*.npz from oracle/make_golden.py::gen_octattn_e2e: ``encode.compress``, encode.py:23-82, and
``encode_mullevel.compress``, encode_mullevel.py:23-85, of the unmodified reference with the shared seeded weights):
symbols handed to the coder bit-exact, PMF rows within 1e-3, bitstream size within 0.5 % (north_star)."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import golden
from test_models_cpu import cfg_oct

SUBS = ("_0_0", "_0_1", "_1")


def _write_mullevel_artefacts(tmp_path, name="k16m"):
    """What ``data_preproc/test_gene.py --mullevel`` leaves behind for one sweep, from the reference's golden rows."""
    g = golden(f"octree_{name}.npz")
    seq, pre = tmp_path / "seq", tmp_path / "pre"
    seq.mkdir(exist_ok=True), pre.mkdir(exist_ok=True)
    f = seq / f"{name}.bin"
    g["points"].astype(np.float32).tofile(f)
    base = str(pre) + "/seq" + name
    cuts = np.cumsum(np.concatenate([[0], g["sub_rows"]]))
    for sfx, a, b in zip(SUBS, cuts[:-1], cuts[1:]):
        np.save(base + sfx + ".npy", g["rows"][a:b].astype(np.int64))
    np.save(base + "_meta.npy", np.array([float(g["bin_num"]), 0.0125, 0.0]))
    return g, str(f), str(pre) + "/"


# ---------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("lw", [False, True])
def test_mullevel_octattn_dataset_matches_reference(tmp_path, lw):
    """dataloaders/encode_dataset_mullevel.EncodeDataset mirror == the reference's ``__getitem__`` (three row files,
    1023 pad rows per block, positions / 2^(deepest level of each file))."""
    from scp_b200.dataloaders.encode_dataset_mullevel import EncodeDataset
    e = golden(f"e2e_octattn_k16m_{'lw' if lw else 'seq'}.npz")
    g, f, pre = _write_mullevel_artefacts(tmp_path)
    ids, pos, data, oct_seq, n, bin_num, chamfer, psnr = EncodeDataset([f], 1024, "kitti", lw, 16, True, pre)[0]
    assert [len(i) for i in ids] == list(e["ds_sizes"]) and len(ids) == (48 if lw else 3)
    assert np.array_equal(np.concatenate(ids), e["ds_ids"])
    assert data[0].dtype == np.int64 and np.array_equal(np.concatenate(data), e["ds_data"].astype(np.int64))
    assert pos[0].dtype == np.float32 and np.array_equal(np.concatenate(pos), e["ds_pos"])
    assert np.array_equal(oct_seq, e["ds_oct_seq"].astype(np.int64))
    assert n == len(g["points"]) and bin_num == int(g["bin_num"]) and chamfer == 0.0125 and psnr == 0
    with pytest.raises(Exception, match="no preproc_path"):
        EncodeDataset([f], 1024, "kitti", lw, 16, True, "")


# ---------------------------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    from scp_b200.models import OctAttention
    return OctAttention(cfg_oct()).cuda()


def _check_stream(path, e):
    ref = len(e["bitstream"])
    got = os.path.getsize(path)
    print(os.path.basename(path), "bytes ours", got, "reference", ref, "dev %.4f%%" % (100 * abs(got - ref) / ref))
    assert abs(got - ref) / ref < 0.005


@gpu
@pytest.mark.parametrize("name,level", [("k12s", 12), ("k14s", 14)])
def test_compress_matches_reference_run(tmp_path, model, name, level):
    """``scp_b200.encode.compress`` on the batch the reference's EncodeDataset yields (non level-wise)."""
    from scp_b200 import encode
    g, e = golden(f"octree_{name}.npz"), golden(f"e2e_octattn_{name}.npz")
    oct_seq = g["rows"].astype(np.int64).copy()
    oct_seq[:, :, 0] -= 1
    assert np.array_equal(oct_seq[:, -1, 0], e["sym"])                     # what the reference handed to its coder
    batch = ([torch.from_numpy(g["oct_ids"])[None]], [torch.from_numpy(g["oct_pos"])[None]],
             [torch.from_numpy(g["oct_data"].astype(np.int64))[None]], torch.from_numpy(oct_seq)[None],
             torch.tensor(len(g["points"])), torch.tensor(int(g["bin_num"])))
    args = types.SimpleNamespace(sequential=False, spher=True, cylin=False)
    bpp, _ = encode.compress(batch, str(tmp_path / "out" / name), model, args)
    _check_stream(str(tmp_path / "out" / name) + ".bin", e)
    assert abs(bpp - float(e["bpp"])) / float(e["bpp"]) < 0.005
    with pytest.raises(NotImplementedError):
        encode.compress(batch, str(tmp_path / "x"), model, types.SimpleNamespace(sequential=True))


@gpu
@pytest.mark.parametrize("name,level", [("k12s", 12), ("k14s", 14)])
def test_pmf_rows_and_encoder_match_reference_run(model, name, level):
    """``Encoder(OctAttention)`` from the raw points: node count / symbols bit-exact, PMFs of every 16th node within 1e-3
    of the rows the reference's coder received, stream within 0.5 %."""
    from scp_b200 import coder
    from scp_b200.encoder import Encoder
    g, e = golden(f"octree_{name}.npz"), golden(f"e2e_octattn_{name}.npz")
    enc = Encoder(model, level, "spher", mullevel=False)
    res = enc.encode([g["points"]])[0]
    assert res.n_nodes == len(e["sym"])
    ref = len(e["bitstream"])
    print(name, "Encoder bytes", len(res.bitstream), "reference", ref)
    assert abs(len(res.bitstream) - ref) / ref < 0.005 and abs(res.bpp - float(e["bpp"])) / float(e["bpp"]) < 0.005
    # PMF rows
    data = torch.from_numpy(g["oct_data"].astype(np.int64)).cuda()
    pos = torch.from_numpy(g["oct_pos"]).cuda()
    L, cs = data.shape[0], 1024
    pm = []
    for a in range(0, L, cs):
        lg = model(data[None, a:a + cs].clone(), pos[None, a:a + cs])[0]
        pm.append(coder.pmf_to_cdf(lg.contiguous(), is_logits=True, want_pmf=True)["pmf"])
    pm = torch.cat(pm)[cs - 1:][::16].cpu().numpy()
    err = np.abs(pm - e["pdf_s16"]).max()
    print(name, "PMF max-abs vs the reference's coder input:", err)
    assert pm.shape == e["pdf_s16"].shape and err < 1e-3


@gpu
@pytest.mark.parametrize("lw", [False, True])
def test_compress_mullevel_matches_reference_run(tmp_path, model, lw):
    """``scp_b200.encode_mullevel.compress`` on the blocks of the mullevel OctAttention dataset (3 sub-octrees, or their 48
    levels): file name convention of encode_mullevel.py:65-70 and the stream size of the reference's run."""
    from scp_b200 import encode_mullevel
    from scp_b200.dataloaders.encode_dataset_mullevel import EncodeDataset
    from torch.utils.data import default_collate
    e = golden(f"e2e_octattn_k16m_{'lw' if lw else 'seq'}.npz")
    g, f, pre = _write_mullevel_artefacts(tmp_path)
    batch = default_collate([EncodeDataset([f], 1024, "kitti", lw, 16, True, pre)[0]])[:-2]
    assert np.array_equal(batch[3][0, :, -1, 0].numpy(), e["sym"])
    args = types.SimpleNamespace(sequential=False, spher=True, cylin=False)
    bpp, _ = encode_mullevel.compress(batch, str(tmp_path / "out" / "k16m"), model, args)
    _check_stream(str(tmp_path / "out" / "k16m") + f"_spher_{48 if lw else 3}_{int(g['bin_num'])}_0.bin", e)
    assert abs(bpp - float(e["bpp"])) / float(e["bpp"]) < 0.005


@gpu
def test_encoder_mullevel_matches_reference_run(model):
    """``Encoder(OctAttention, mullevel=True)`` from the raw points == the reference's three-block run."""
    from scp_b200.encoder import Encoder
    g, e = golden("octree_k16m.npz"), golden("e2e_octattn_k16m_seq.npz")
    res = Encoder(model, 16, "spher", mullevel=True).encode([g["points"]])[0]
    ref = len(e["bitstream"])
    print("k16m Encoder bytes", len(res.bitstream), "reference", ref)
    assert res.n_nodes == len(e["sym"]) and abs(len(res.bitstream) - ref) / ref < 0.005


@gpu
def test_encoder_batch_equals_frame_by_frame(model):
    from scp_b200 import synth
    from scp_b200.encoder import Encoder
    enc = Encoder(model, 12, "spher", mullevel=False)
    frames = [synth.make_frame("kitti", s, 12, "spher", guard=True, n_points=n)[0] for s, n in ((1, 3000), (2, 1200), (3, 2000))]
    together = enc.encode(frames)
    for f, r in zip(frames, together):
        single = enc.encode([f])[0]
        assert single.bitstream == r.bitstream and single.n_nodes == r.n_nodes


@gpu
def test_main_over_two_ranks_equals_single_process(tmp_path):
    """``torchrun -m scp_b200.encode`` with two ranks (frame-wise partition, scp_b200/partition.py): the union of the ranks'
    bitstream files and the summary equal the single-process run's."""
    import subprocess
    import sys
    from scp_b200 import synth
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seq = tmp_path / "seq"
    seq.mkdir()
    files = []
    for s in range(3):
        p = seq / f"{s:06d}.bin"
        synth.make_frame("kitti", 20 + s, 12, "spher", guard=True, n_points=2000)[0].astype(np.float32).tofile(p)
        files.append(str(p))
    env = dict(os.environ, PYTHONPATH=root)
    common = ["--test_files", *files, "--lidar_level", "12", "--spher"]
    for tag, launcher in (("one", [sys.executable, "-m", "scp_b200.encode"]),
                          ("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29611", "-m", "scp_b200.encode"])):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run(launcher + common + ["--out_dir", str(d / "out")], cwd=d, env=env, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    one = sorted(os.listdir(tmp_path / "one" / "out"))
    assert one == sorted(os.listdir(tmp_path / "two" / "out")) and len([f for f in one if f.endswith(".bin")]) == 3
    for f in one:
        if f.endswith(".bin"):
            assert open(tmp_path / "one" / "out" / f, "rb").read() == open(tmp_path / "two" / "out" / f, "rb").read()
    rep = [[l for l in open(tmp_path / t / "test_results_same_kitti_12.txt").read().splitlines() if l.startswith(("sample", "bpp"))]
           for t in ("one", "two")]
    assert rep[0] == rep[1] and rep[0][0] == "sample number: 3"
