"""GPU parity of the octree pipeline (A1-A5) against the reference goldens and the CPU oracle. Bit-exact."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import octree_np as onp

pytestmark = pytest.mark.gpu

CASES = ["k12s", "k14c", "k16m", "f17s", "k10c"]
ALL = ("occ", "level", "octant", "parent", "pos", "ctx", "pos_norm", "ctx_pos", "rows_i64", "voxel_key")


def run_case(g, outputs=ALL):
    from scp_b200 import octree
    pts = torch.from_numpy(g["points"]).cuda()
    mode, lvl = str(g["mode"]), int(g["level"])
    if bool(g["mullevel"]):
        jobs = [octree.JobSpec(0, float(q), list(p), drop_last=True, lidar_level=lvl, pos_eps_last=False)
                for q, p in zip(g["qs"], octree.MULLEVEL_PATHS)]
    else:
        jobs = [octree.JobSpec(0, float(g["qs"][0]), None, lidar_level=lvl)]
    b = octree.OctreeBuilder().plan(pts, [0, len(pts)], jobs, mode)
    out = b.emit(outputs)
    torch.cuda.synchronize()
    return b, {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("name", CASES)
def test_rows_bit_exact_vs_reference(name):
    g = golden(f"octree_{name}.npz")
    b, out = run_case(g)
    ref = g["rows"].astype(np.int64)
    assert out["rows_i64"].shape == ref.shape
    assert np.array_equal(out["rows_i64"], ref)
    assert np.array_equal(out["occ"].astype(np.int64), ref[:, 3, 0])
    assert np.array_equal(out["level"].astype(np.int64), ref[:, 3, 1])
    assert np.array_equal(out["octant"].astype(np.int64), ref[:, 3, 2])
    assert np.array_equal(out["pos"].astype(np.int64), ref[:, 3, 3:])
    assert np.array_equal(out["ctx_pos"].astype(np.int64), ref[:, :, 3:])
    assert b.infos[0].bin_num == float(g["bin_num"])
    if "z_offset" in g.files:
        assert b.infos[0].offset[2] == float(g["z_offset"])
    if "sub_rows" in g.files:
        assert [i.n_rows for i in b.infos] == list(g["sub_rows"])


@pytest.mark.parametrize("name", CASES)
def test_dataset_tensors_bit_exact_vs_reference(name):
    g = golden(f"octree_{name}.npz")
    b, out = run_case(g, ("ctx", "pos_norm"))
    sizes = [n for i in b.infos for n in i.level_rows]
    assert sizes == list(g["level_sizes"])
    assert np.array_equal(out["ctx"].astype(np.int64), g["ds_data"].astype(np.int64))
    assert out["pos_norm"].dtype == np.float32
    assert np.array_equal(out["pos_norm"], g["ds_pos"], equal_nan=True)
    mm = [p for i in b.infos for p in i.pos_mm]
    assert np.array_equal(np.array(mm, np.int64), g["ds_pos_mm"])


def full_frame(kind, seed, level, mode, mul):
    from scp_b200 import synth
    pts, qs0 = synth.make_frame(kind, seed, level, mode, guard=False)
    qf = synth.KITTI_QS if kind == "kitti" else synth.FORD_QS
    for i in range(3 if mul else 1):
        pts = synth.guard_band(pts, qf(level + i), mode, margin=0.03)
    return pts


@pytest.mark.parametrize("kind,seed,level,mode,mul", [
    ("kitti", 0, 12, "spher", False), ("kitti", 1, 14, "cylin", False), ("kitti", 2, 16, "spher", True),
    ("ford", 3, 17, "spher", False)])
def test_full_size_frames_vs_oracle(kind, seed, level, mode, mul):
    from scp_b200 import octree, synth
    pts = full_frame(kind, seed, level, mode, mul)
    qf = synth.KITTI_QS if kind == "kitti" else synth.FORD_QS
    if mul:
        jobs = octree.mullevel_jobs(0, level, kind)
    else:
        jobs = [octree.JobSpec(0, qf(level), None, lidar_level=level)]
    b = octree.OctreeBuilder().plan(torch.from_numpy(pts).cuda(), [0, len(pts)], jobs, mode)
    out = b.emit(("rows_i64", "ctx", "pos_norm", "voxel_key"))
    rows = out["rows_i64"].cpu().numpy()
    exp_rows, exp_data, exp_pos = [], [], []
    for j in jobs:
        q = onp.quantize(pts[:, :3], j.qs, mode)["q"]
        r = onp.tree_rows(q, morton_path=j.morton_path, drop_last=j.drop_last)["rows"]
        exp_rows.append(r)
        _, poss, _, data, _ = onp.ehem_level_split(r, level, mullevel=mul)
        exp_data += data
        exp_pos += [p.T for p in poss]
    assert np.array_equal(rows, np.vstack(exp_rows))
    assert np.array_equal(out["ctx"].cpu().numpy().astype(np.int64), np.concatenate(exp_data, 0))
    assert np.array_equal(out["pos_norm"].cpu().numpy(), np.concatenate(exp_pos, 0), equal_nan=True)
    # voxel keys ascending and unique per job (sortedness property at full size)
    vk = out["voxel_key"].cpu().numpy()
    for i in b.infos:
        seg = vk[i.voxel_start:i.voxel_start + i.n_voxels]
        assert (np.diff(seg) > 0).all()


def test_batch_of_frames_equals_single_frames():
    from scp_b200 import octree, synth
    frames = [synth.make_frame("kitti", s, 12, "spher", guard=True, n_points=n)[0]
              for s, n in ((10, 20000), (11, 5000), (12, 33333), (13, 1), (14, 4097))]
    qs = synth.KITTI_QS(12)
    offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])])
    allp = torch.from_numpy(np.concatenate(frames, 0)).cuda()
    jobs = [octree.JobSpec(i, qs, None, lidar_level=12) for i in range(len(frames))]
    b = octree.OctreeBuilder().plan(allp, offs, jobs, "spher")
    out = b.emit(("rows_i64",))["rows_i64"].cpu().numpy()
    for i, f in enumerate(frames):
        info = b.infos[i]
        if len(f) == 1:
            assert info.depth >= 1
        r = onp.tree_rows(onp.quantize(f[:, :3], qs, "spher")["q"])["rows"]
        assert np.array_equal(out[info.row_start:info.row_start + info.n_rows], r), i


def test_cartesian_mode_and_legacy_octree_api():
    import ctypes as C
    from scp_b200 import _lib
    rng = np.random.default_rng(5)
    q = rng.integers(0, 1 << 9, (3000, 3))
    q = np.unique(q, axis=0)
    lib = _lib.require_device()
    vec = lib.new_vector()
    data = np.ascontiguousarray(q.astype(np.float64))
    codes = lib.genOctreeInterface(vec, data.ctypes.data_as(C.POINTER(C.c_double)), len(q))
    assert codes
    exp = onp.tree_rows(q)
    nlev = lib.vector_size(vec)
    assert nlev == exp["depth"]
    assert lib.int_size(codes) == len(exp["codes"])
    got_codes = np.array([lib.int_get(codes, i) for i in range(lib.int_size(codes))])
    assert np.array_equal(got_codes, exp["codes"])
    r = 0
    rows = exp["rows"]
    for L in range(nlev):
        lv = lib.vector_get(vec, L)
        assert lib.Nodes_size(lv) == exp["level_counts"][L]
        for k in (0, lib.Nodes_size(lv) - 1):
            nd = lib.Nodes_get(lv, k).contents
            row = rows[r + k, 3]
            assert (nd.oct, nd.octant, list(nd.pos)) == (row[0], row[2], list(row[3:]))
            assert nd.nodeid == r + k + 1
        r += lib.Nodes_size(lv)
    lib.delete_vector(vec)


def test_segmented_sort_vs_numpy():
    from scp_b200 import octree
    rng = np.random.default_rng(0)
    sizes = [1, 5000, 4096, 4097, 123457, 2]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for bits in (64, 40, 9):
        keys = rng.integers(0, 2 ** 63 - 1, offs[-1], dtype=np.int64)
        if bits < 64:
            keys &= (1 << bits) - 1
        t = torch.from_numpy(keys.copy()).cuda()
        octree.segmented_sort(t, offs, bits)
        got = t.cpu().numpy()
        for a, b_ in zip(offs[:-1], offs[1:]):
            assert np.array_equal(got[a:b_], np.sort(keys[a:b_])), (bits, a)


def test_plan_errors_are_reported():
    from scp_b200 import octree, _lib
    pts = torch.zeros((10, 3), device="cuda")
    pts[:, 0] = 1e9
    with pytest.raises(_lib.ScpError):
        octree.OctreeBuilder().plan(pts, [0, 10], [octree.JobSpec(0, 1e-3, None)], "cart")
    with pytest.raises(_lib.ScpError):
        octree.OctreeBuilder().plan(pts, [0, 0], [octree.JobSpec(0, 1.0, None)], "cart")


@pytest.mark.parametrize("mul,level,mode", [(False, 12, "spher"), (True, 16, "spher"), (False, 14, "cylin"), (True, 14, "cylin")])
def test_tree_builders_and_quantise_paths_are_bit_identical(mul, level, mode):
    """Three tree builders -- node records in one pass over the sorted keys (0: k_emit_nodes + k_occupancy + k_context*), node
    records level by level (1: k_level_pass), and the warp-autonomous key pass that writes occupancy and records together for
    the encoder's outputs (2, default: k_tree_occ + k_context_lean) -- and two front ends -- the fused transform + quantise + morton_path filter + compaction kernel
    with the per-job sort schedule (default) and the older quantise-all / filter / compact / sort sequence (SCP_QUANT_OLD=1)
    -- produce identical outputs, ragged batch included (a 1-point frame, frames that end inside a 32-key group, full-size
    frames)."""
    import os
    from scp_b200 import _lib, octree, synth
    lib = _lib.require_device()
    frames = [synth.make_frame("kitti", s, level, mode, guard=True, n_points=n)[0]
              for s, n in ((20, 120000), (21, 1), (22, 4097), (23, 33), (24, 60000))]
    offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])])
    allp = torch.from_numpy(np.concatenate(frames, 0)).cuda()
    if mul:
        jobs = [j for i in range(len(frames)) for j in octree.mullevel_jobs(i, level)]
    else:
        jobs = [octree.JobSpec(i, synth.KITTI_QS(level), None, lidar_level=level) for i in range(len(frames))]
    names = ("occ", "level", "octant", "parent", "pos", "ctx", "pos_norm", "ctx_pos", "voxel_key", "sym")
    res = []
    for builder, old_quant in ((0, True), (1, True), (2, True), (0, False), (2, False)):
        old = lib.scp_set_tree_builder(builder)
        if old_quant:
            os.environ["SCP_QUANT_OLD"] = "1"
        try:
            b = octree.OctreeBuilder().plan(allp, offs, jobs, mode)
            out = b.emit(names)
            lean = b.emit(("occ", "sym", "ctx", "pos_norm", "voxel_key"))
            res.append(({k: v.cpu().numpy() for k, v in out.items()}, {k: v.cpu().numpy() for k, v in lean.items()},
                        [(i.depth, i.n_rows, i.n_voxels, tuple(i.level_rows), tuple(i.pos_mm)) for i in b.infos], b.total_kept))
        finally:
            lib.scp_set_tree_builder(old)
            os.environ.pop("SCP_QUANT_OLD", None)
    for r in res[1:]:
        assert res[0][2] == r[2] and res[0][3] == r[3]
        for a, c in ((res[0][0], r[0]), (res[0][1], r[1])):
            for k in a:
                assert np.array_equal(a[k], c[k], equal_nan=True), k
