"""Distortion report: Chamfer distance / D1 PSNR / dequantised cloud.

CPU part: the oracle (oracle/metrics_np.py) against the reference's own ``pt.distChamfer`` outputs on the golden frames
(tests/golden/metrics.npz, written by oracle/make_golden.py from the unmodified reference).
GPU part: ``scp_nn_dist2`` / ``scp_dequantise_keys`` through the C ABI against the oracle and the same goldens."""
import math

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import metrics_np as om
from oracle import octree_np as onp

CASES = {"k12s": "spher", "k14c": "cylin", "k16m": "spher", "f17s": "spher", "k10c": "cylin", "k14s": "spher"}


def _pc(name):
    return golden(f"octree_{name}.npz")["points"][:, :3].astype(np.float32)


# ---------------------------------------------------------------------------------------------- CPU: oracle pin
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_chamfer_matches_reference(name):
    m = golden("metrics.npz")
    assert om.dist_chamfer(_pc(name), m[name + "_q"]) == pytest.approx(float(m[name + "_chamfer"]), rel=1e-12)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_psnr_matches_pc_error(name):
    """The D1 PSNR restatement against the reference's own tool: ``pt.pcerror`` -> utils/pc_error (mpeg-pcc-dmetric 0.13.5)
    -> ``utils.get_psnr`` on (original cloud, proc_pc / mul_proc_pc cloud), oracle/make_golden.py::gen_metrics.  The tool
    prints six significant digits."""
    m = golden("metrics.npz")
    peak = 30000.0 if name == "f17s" else 59.70
    mse, psnr = om.d1_psnr(_pc(name), m[name + "_q"], peak)
    assert mse == pytest.approx(float(m[name + "_mseF"]), rel=2e-5)
    assert psnr == pytest.approx(float(m[name + "_psnr"]), abs=1.5e-4)


def test_oracle_dequantise_matches_reference_cloud():
    """proc_pc's quantised cloud (np.unique order) == oracle voxels -> keys -> dequantise, as a set."""
    g, m = golden("octree_k12s.npz"), golden("metrics.npz")
    qz = onp.quantize(g["points"][:, :3], float(g["qs"][0]), "spher")
    vox = onp.voxels_unique(qz["q"]).astype(np.uint64)
    keys = np.zeros(len(vox), np.uint64)
    for b in range(21):
        for c in range(3):
            keys |= ((vox[:, c] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + 2 - c)
    pts = om.dequantise_keys(keys, qz["steps"], np.zeros(3), "spher")
    assert len(pts) == len(m["k12s_q"])
    assert om.nn_dist(pts, m["k12s_q"]).max() < 1e-4 and om.nn_dist(m["k12s_q"], pts).max() < 1e-4


def test_oracle_psnr_definition():
    a = np.array([[0, 0, 0], [1, 0, 0]], np.float64)
    b = np.array([[0, 0, 0.5], [1, 0, 0]], np.float64)
    mse, psnr = om.d1_psnr(a, b, 10.0)
    assert mse == 0.125 and psnr == pytest.approx(10 * math.log10(300 / 0.125))


# ---------------------------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


def _brute(q, c):
    q, c = q.astype(np.float64), c.astype(np.float64)
    d = q[:, None, :] - c[None, :, :]
    return ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]).min(1)


@gpu
@pytest.mark.parametrize("nq,nc", [(1, 1), (513, 700), (100, 50000), (3000, 1025), (2048, 2048)])
def test_nn_dist2_is_bit_exact(nq, nc):
    from scp_b200 import metrics
    r = np.random.default_rng(nq * 7 + nc)
    q = (r.standard_normal((nq, 3)) * 30).astype(np.float32)
    c = (r.standard_normal((nc, 3)) * 30).astype(np.float32)
    c[: min(nq, nc) // 3] = q[: min(nq, nc) // 3]                      # exact hits (distance 0)
    got = metrics.nn_dist2(q, c).cpu().numpy()
    ref = np.concatenate([_brute(q[i:i + 256], c) for i in range(0, nq, 256)])
    assert np.array_equal(got, ref)
    assert np.allclose(np.sqrt(got), om.nn_dist(q, c), rtol=1e-13, atol=0)


@gpu
def test_nn_dist2_float64_inputs_and_errors():
    from scp_b200 import metrics
    r = np.random.default_rng(5)
    q, c = r.standard_normal((700, 3)) * 1e3, r.standard_normal((900, 3)) * 1e3
    assert np.array_equal(metrics.nn_dist2(q, c).cpu().numpy(), _brute(q, c))
    assert metrics.nn_dist2(np.zeros((0, 3)), c).shape == (0,)
    with pytest.raises(ValueError):
        metrics.nn_dist2(q, np.zeros((0, 3)))
    with pytest.raises(ValueError):
        metrics.nn_dist2(q[:, :2], c)


@gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_chamfer_matches_reference_golden(name):
    from scp_b200 import metrics
    from scp_b200.data_preproc import pt as pointCloud
    m = golden("metrics.npz")
    pc, q = _pc(name), m[name + "_q"]
    want = float(m[name + "_chamfer"])
    assert metrics.distChamfer(pc, q) == pytest.approx(want, rel=1e-12)
    assert pointCloud.distChamfer(pc, q) == pytest.approx(want, rel=1e-12)
    assert metrics.distChamfer(pc, q, scale=2.0) == pytest.approx(om.dist_chamfer(pc / np.float32(2), q / 2.0), rel=1e-6)
    peak = metrics.FORD_PEAK if name == "f17s" else metrics.KITTI_PEAK
    mse, psnr = metrics.d1_psnr(pc, q, peak)
    o_mse, o_psnr = om.d1_psnr(pc, q, peak)
    assert mse == pytest.approx(o_mse, rel=1e-12) and psnr == pytest.approx(o_psnr, rel=1e-12)
    ch, ps = metrics.distortion(pc, q, peak)
    assert ch == pytest.approx(want, rel=1e-12) and ps == pytest.approx(o_psnr, rel=1e-12)
    # ... and against the reference's pc_error run itself (six printed digits)
    assert mse == pytest.approx(float(m[name + "_mseF"]), rel=2e-5) and psnr == pytest.approx(float(m[name + "_psnr"]), abs=1.5e-4)


def _jobs(name, g):
    from scp_b200 import octree as oc
    if bool(g["mullevel"]):
        return oc.mullevel_jobs(0, int(g["level"]), "kitti")
    return [oc.JobSpec(0, float(g["qs"][0]), None)]


@gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_dequantised_cloud_of_the_octree_build(name):
    """voxel keys of the CUDA build -> scp_dequantise_keys: equal to the oracle's float64 formula, and the same point
    set (hence the same Chamfer distance) as the reference's proc_pc / mul_proc_pc output."""
    from scp_b200 import metrics, octree as oc
    g, m = golden(f"octree_{name}.npz"), golden("metrics.npz")
    pc, mode = _pc(name), CASES[name]
    b = oc.OctreeBuilder().plan(torch.from_numpy(pc).cuda(), [0, len(pc)], _jobs(name, g), mode)
    vk = b.emit(("voxel_key",))["voxel_key"]
    cloud = metrics.dequantised_cloud(b, vk, mode).cpu().numpy()
    want = np.vstack([om.dequantise_keys(vk[i.voxel_start:i.voxel_start + i.n_voxels].cpu().numpy(), i.steps,
                                         np.zeros(3) if mode == "spher" else i.offset, mode) for i in b.infos])
    assert cloud.shape == want.shape == m[name + "_q"].shape
    assert np.allclose(cloud, want, rtol=1e-13, atol=1e-12 * np.abs(want).max())
    tol = 1e-5 * max(1.0, np.abs(want).max())                # the reference's single-level cloud is float32
    assert om.nn_dist(cloud, m[name + "_q"]).max() < tol and om.nn_dist(m[name + "_q"], cloud).max() < tol
    # mul_proc_pc keeps float64 (golden equal to rounding); proc_pc rounds v*steps to float32 BEFORE the trigonometric map
    # (data_preprocess.py:70), which moves every point by ~1e-7 of its range: Chamfer agrees to that order only
    rel = 1e-9 if bool(g["mullevel"]) else 5e-3
    assert metrics.distChamfer(pc, cloud) == pytest.approx(float(m[name + "_chamfer"]), rel=rel)


@gpu
def test_datasets_report_distortion(tmp_path):
    """EncodeEHEMDataset (single level, mullevel) and EncodeDataset return (chamfer, psnr) as their last two items."""
    from scp_b200 import metrics
    from scp_b200.dataloaders.encode_dataset import EncodeDataset
    from scp_b200.dataloaders.encode_dataset_ehem import EncodeEHEMDataset
    from scp_b200.dataloaders.encode_dataset_ehem_mullevel import EncodeEHEMDataset as MulDataset
    m = golden("metrics.npz")

    def bin_file(name):
        f = tmp_path / f"{name}.bin"
        np.hstack([_pc(name), np.zeros((len(_pc(name)), 1), np.float32)]).tofile(f)
        return str(f)

    def check(item, name, psnr_zero=False):
        chamfer, psnr = item[-2], item[-1]
        assert chamfer == pytest.approx(float(m[name + "_chamfer"]), rel=1e-9 if name == "k16m" else 1e-3)
        if psnr_zero:
            assert psnr == 0
        else:
            assert psnr == pytest.approx(om.d1_psnr(_pc(name), m[name + "_q"], metrics.KITTI_PEAK)[1], rel=1e-4)

    check(EncodeEHEMDataset([bin_file("k12s")], 8192, "kitti", True, 12, False, True)[0], "k12s")
    check(EncodeEHEMDataset([bin_file("k14c")], 8192, "kitti", True, 14, True, False)[0], "k14c", psnr_zero=True)
    check(MulDataset([bin_file("k16m")], 8192, "kitti", True, 16, False, True)[0], "k16m")
    check(EncodeDataset([bin_file("k12s")], 1024, "kitti", False, 12, True)[0], "k12s")


@gpu
def test_full_size_frame_distortion():
    """BASELINE.json configs[1] size: 120 k-point sweep against its level-16 mullevel cloud, both directions, against the
    KD-tree oracle; and the size-independent properties (a cloud against itself is 0, PSNR grows ~6 dB per level)."""
    from scp_b200 import metrics, octree as oc, synth
    pc = synth.kitti_sweep(3, 120000)[:, :3].astype(np.float32)
    xyz = torch.from_numpy(pc).cuda()
    psnrs = []
    for level in (12, 16):
        b = oc.OctreeBuilder().plan(xyz, [0, len(pc)], oc.mullevel_jobs(0, level, "kitti"), "spher")
        cloud = metrics.dequantised_cloud(b, b.emit(("voxel_key",))["voxel_key"], "spher")
        ch, ps = metrics.distortion(pc, cloud, metrics.KITTI_PEAK)
        psnrs.append(ps)
    q = cloud.cpu().numpy()
    assert ch == pytest.approx(om.dist_chamfer(pc, q), rel=1e-12)
    assert ps == pytest.approx(om.d1_psnr(pc, q, metrics.KITTI_PEAK)[1], rel=1e-12)
    assert 18 < psnrs[1] - psnrs[0] < 30                               # 4 levels, ~6 dB each
    assert metrics.distChamfer(pc, pc) == 0.0 and metrics.d1_psnr(pc, pc, 1.0)[1] == math.inf


@gpu
def test_encoder_reports_distortion_per_frame():
    """``Encoder(distortion=True)``: every FrameResult of a ragged batch carries the Chamfer distance / D1 PSNR of ITS
    frame (original points against the dequantised voxels of its three sub-octrees), equal to the KD-tree oracle."""
    from test_models_cpu import cfg_ehem
    from scp_b200 import metrics, octree as oc, synth
    from scp_b200.encoder import Encoder
    from scp_b200.models import EHEM
    frames = [synth.kitti_sweep(7, 6000).astype(np.float32), synth.kitti_sweep(8, 2500).astype(np.float32)]
    model = EHEM(cfg_ehem()).cuda()
    res = Encoder(model, 16, "spher", mullevel=True, distortion=True).encode(frames)
    plain = Encoder(model, 16, "spher", mullevel=True).encode(frames)
    for fr, r, p in zip(frames, res, plain):
        assert r.bitstream == p.bitstream and p.chamfer is None and p.psnr is None
        pc = fr[:, :3]
        b = oc.OctreeBuilder().plan(torch.from_numpy(fr).cuda(), [0, len(fr)], oc.mullevel_jobs(0, 16, "kitti"), "spher")
        q = metrics.dequantised_cloud(b, b.emit(("voxel_key",))["voxel_key"], "spher").cpu().numpy()
        assert r.chamfer == pytest.approx(om.dist_chamfer(pc, q), rel=1e-12)
        assert r.psnr == pytest.approx(om.d1_psnr(pc, q, metrics.KITTI_PEAK)[1], rel=1e-12)
