"""Encode -> decode round trips through the whole GPU path (the reference's only
correctness check is its decoder's assert, decode_ehem.py:184 / decode_ehem_mullevel.py:175).

The decoder rebuilds every level from the occupancy bytes it has decoded so far, feeds the entropy model the same
context bytes and float32 positions as the encoder and must therefore recover, bit for bit, the occupancy sequence and
the voxel set of the encoder's octree.  This is also the size-independent parity property for the full-size
configurations (BASELINE.json configs[1]: level-16 mullevel, 120 k points, 529 k nodes)."""
import numpy as np
import pytest
import torch

from conftest import golden
from test_models_cpu import cfg_ehem

pytestmark = pytest.mark.gpu


def _compact3(k):
    k = k.astype(np.uint64)
    out = np.zeros_like(k)
    for b in range(21):
        out |= ((k >> np.uint64(3 * b)) & np.uint64(1)) << np.uint64(b)
    return out.astype(np.int64)


def _encoder_tree(enc, pts):
    """occupancy rows and voxel coordinates per sub-octree as the ENCODER built them (device octree kernels)."""
    xyz = torch.from_numpy(pts).cuda()
    b, t, per_frame = enc.build_context(xyz, [0, len(pts)])
    out = b.emit(("occ", "voxel_key"), finish=True)
    occ, vk = out["occ"].cpu().numpy(), out["voxel_key"].cpu().numpy().view(np.uint64)
    trees = []
    for i in b.infos:
        k = vk[i.voxel_start: i.voxel_start + i.n_voxels]
        vox = np.stack([_compact3(k >> np.uint64(2)), _compact3(k >> np.uint64(1)), _compact3(k)], 1)
        trees.append((occ[i.row_start: i.row_start + i.n_rows], vox))
    return trees


def _round_trip(level, mode, mul, pts, kind="kitti"):
    from scp_b200.decoder import Decoder
    from scp_b200.encoder import Encoder
    from scp_b200.models import EHEM
    model = EHEM(cfg_ehem()).cuda()
    enc = Encoder(model, level, mode, mullevel=mul, kind=kind)
    res = enc.encode([pts])[0]
    trees = _encoder_tree(enc, pts)
    dec = Decoder(model, level, mode, mullevel=mul, kind=kind).decode(res)
    assert dec.n_symbols == res.n_nodes == sum(len(o) for o, _ in trees)
    assert len(dec.occ) == len(trees)
    for j, ((occ, vox), d_occ, d_vox) in enumerate(zip(trees, dec.occ, dec.voxels)):
        assert np.array_equal(d_occ, occ), f"sub-octree {j}: occupancy sequence differs"
        if mul:
            # encode_mullevel drops the last node of every sub-octree (Octree.py:259-262): its <= 8 voxels are not coded
            assert 0 < len(vox) - len(d_vox) <= 8 and np.array_equal(d_vox, vox[:len(d_vox)])
        else:
            assert np.array_equal(d_vox, vox), f"sub-octree {j}: voxel set differs"
    return res, dec, trees


def test_round_trip_small_golden_frames():
    """The frames of the reference goldens (k12s: proc_pc, k16m: mul_proc_pc x3): round trip, and for k12s the decoded and
    dequantised voxels equal the points the reference's proc_pc returns (golden 'dequant', data_preprocess.py:82-92)."""
    from scp_b200.decoder import dequantise
    g = golden("octree_k12s.npz")
    res, dec, trees = _round_trip(12, "spher", False, g["points"])
    qs = float(g["qs"][0])
    bn = float(g["bin_num"])
    steps = np.array([qs, np.float32(2 * np.pi) / np.float32(bn - 1), np.float32(np.pi) / np.float32(bn - 1)], np.float64)
    p = dequantise(dec.voxels[0], steps, [0, 0, 0], "spher")
    ref = g["dequant"].astype(np.float64)
    assert p.shape == ref.shape
    d = np.abs(p[:, None, :] - ref[None, :, :]).max(-1)            # float32 vs float64 trigonometry: match as sets
    assert d.min(1).max() < 1e-4 and d.min(0).max() < 1e-4
    g = golden("octree_k16m.npz")
    res, dec, trees = _round_trip(16, "spher", True, g["points"])
    assert [len(o) for o in dec.occ] == [int(x) for x in g["sub_rows"]]


def test_round_trip_cylin_small():
    from scp_b200 import synth
    pts = synth.make_frame("kitti", 3, 12, "cylin", guard=True, n_points=20000)[0]
    _round_trip(12, "cylin", False, pts)


def test_round_trip_single_node_levels_below_the_root():
    """A tight cluster far from the sensor: the levels below the root hold ONE node each.  encode.py:123 codes the root's
    row again for such a level (no running offset) and never codes the node itself -- undecodable; here the node itself is
    coded (what encode_mullevel.py:120 does) and the round trip is lossless.  ``reference_single_node_defect=True`` still
    gives the reference's order."""
    from scp_b200 import coder
    r = np.random.default_rng(0)
    pts = (np.array([[60.0, 40.0, -1.0, 0.0]]) + np.concatenate([r.uniform(-0.4, 0.4, (300, 3)), np.zeros((300, 1))], 1)).astype(np.float32)
    res, dec, trees = _round_trip(12, "spher", False, pts)
    from scp_b200.encoder import Encoder
    from scp_b200.models import EHEM
    enc = Encoder(EHEM(cfg_ehem()).cuda(), 12, "spher", mullevel=False)
    b, t, _ = enc.build_context(torch.from_numpy(pts).cuda(), [0, len(pts)])
    level_rows = b.infos[0].level_rows
    assert level_rows[0] == 1 and 1 in level_rows[1:], level_rows          # the case under test
    occ = b.emit(("occ",))["occ"]
    fixed, _ = coder.coding_order(level_rows, 8192, occ, mullevel=False)
    ref, _ = coder.coding_order(level_rows, 8192, occ, mullevel=False, reference_single_node_defect=True)
    first = 1 + level_rows[1:].index(1)
    row = sum(level_rows[:first])
    assert int(fixed[row]) == row and int(ref[row]) == 0 and sorted(fixed.tolist()) == list(range(sum(level_rows)))


def test_round_trip_full_size_k12():
    """BASELINE.json configs[0] at full size: 120 k points, spherical level 12 (~77 k nodes): lossless."""
    from scp_b200 import synth
    pts = synth.make_frame("kitti", 0, 12, "spher", guard=True, n_points=120000)[0]
    res, dec, trees = _round_trip(12, "spher", False, pts)
    print("k12 full: nodes", res.n_nodes, "voxels", len(dec.voxels[0]), "bytes", len(res.bitstream))


def test_round_trip_full_size_k16_mullevel():
    """BASELINE.json configs[1] at full size: 120 k points, three sub-octrees of depth 15/16/17, ~529 k coded nodes."""
    from scp_b200 import synth
    pts = synth.make_frame("kitti", 0, 16, "spher", guard=True, n_points=120000)[0]
    res, dec, trees = _round_trip(16, "spher", True, pts)
    assert dec.depths == [15, 16, 17]
    print("k16m full: nodes", res.n_nodes, "bytes", len(res.bitstream), "bpp", res.bpp)


def test_round_trip_full_size_ford_l17():
    """BASELINE.json configs[2]: Ford-shaped ~80 k points (integer millimetres), spherical level 17, qs = 2 -- the deepest
    octree (depth 16) and the largest context tensors."""
    from scp_b200 import synth
    pts = synth.make_frame("ford", 0, 17, "spher", guard=True)[0]
    res, dec, trees = _round_trip(17, "spher", False, pts, kind="ford")
    print("f17 full: points", len(pts), "nodes", res.n_nodes, "depth", dec.depths, "bytes", len(res.bitstream))


def test_round_trip_full_size_k14_cylin():
    """BASELINE.json configs[4] (one frame of the batch): KITTI-shaped 120 k points, cylindrical level 14."""
    from scp_b200 import synth
    pts = synth.make_frame("kitti", 1, 14, "cylin", guard=True, n_points=120000)[0]
    res, dec, trees = _round_trip(14, "cylin", False, pts)
    print("k14c full: nodes", res.n_nodes, "depth", dec.depths, "bytes", len(res.bitstream))


@pytest.mark.parametrize("mul,level", [(False, 12), (True, 16)])
def test_decode_batch_lock_step_over_ragged_frames(mul, level):
    """Decoder.decode_batch runs the frames in lock-step (one phase-1 call per level, one phase-2 call per window index over
    all frames).  Frames of very different sizes (different numbers of windows per level, a 3-point frame whose trees are
    a single path): every frame decodes to its encoder tree."""
    from scp_b200 import synth
    from scp_b200.decoder import Decoder
    from scp_b200.encoder import Encoder
    from scp_b200.models import EHEM
    model = EHEM(cfg_ehem()).cuda()
    enc = Encoder(model, level, "spher", mullevel=mul)
    frames = [synth.make_frame("kitti", s, level, "spher", guard=True, n_points=n)[0] for s, n in ((31, 40000), (32, 3), (33, 9000), (34, 120000))]
    res = enc.encode(frames)
    dec = Decoder(model, level, "spher", mullevel=mul).decode_batch(res)
    for fr, r, d in zip(frames, res, dec):
        trees = _encoder_tree(enc, fr)
        assert d.n_symbols == r.n_nodes
        for (occ, vox), d_occ, d_vox in zip(trees, d.occ, d.voxels):
            assert np.array_equal(d_occ, occ)
            if mul:
                assert len(vox) - len(d_vox) <= 8 and np.array_equal(d_vox, vox[:len(d_vox)])
            else:
                assert np.array_equal(d_vox, vox)


def test_dequantise_matches_reference_formula():
    from scp_b200.decoder import dequantise
    rng = np.random.default_rng(0)
    v = rng.integers(0, 4096, (1000, 3))
    steps = np.array([0.0977, 2 * np.pi / 4095, np.pi / 4095])
    p = dequantise(torch.from_numpy(v).cuda(), steps, [0, 0, 0], "spher")
    q = v * steps
    ref = np.stack([q[:, 0] * np.sin(q[:, 2]) * np.cos(q[:, 1]), q[:, 0] * np.sin(q[:, 2]) * np.sin(q[:, 1]), q[:, 0] * np.cos(q[:, 2])], 1)
    assert np.allclose(p, ref, rtol=0, atol=1e-9)
