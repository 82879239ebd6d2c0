"""Host side of the batched octree builder.

Thin Python over ``scp_octree_plan`` / ``scp_octree_emit`` of the C ABI: torch is used only to own
device memory and streams.  One *job* = one octree over one frame's points, i.e. one call of the
reference's ``proc_pc`` (data_preprocess.py:13) or ``mul_proc_pc`` (:95); a batch of jobs is built by
one sequence of kernel launches.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib

MODES = {"cart": 0, "spher": 1, "cylin": 2}
MULLEVEL_PATHS = ([0, 0], [0, 1], [1])      # encode_dataset_ehem_mullevel.py:165,175,185


@dataclass
class JobSpec:
    frame: int
    qs: float
    morton_path: Optional[Sequence[int]] = None     # None = proc_pc; list = mul_proc_pc
    drop_last: bool = False
    lidar_level: int = 255
    pos_eps_last: bool = True
    cart_offset: float = 0.0


@dataclass
class JobResult:
    depth: int
    n_points: int
    n_voxels: int
    n_rows: int
    row_start: int
    voxel_start: int
    level_rows: List[int]
    bin_num: float
    steps: np.ndarray
    offset: np.ndarray
    pos_mm: List[tuple] = field(default_factory=list)


def mullevel_jobs(frame, level, kind="kitti"):
    """The three sub-octrees of encode_dataset_ehem_mullevel.py:157-186 / test_gene.py:28-60."""
    from .synth import KITTI_QS, FORD_QS
    qf = KITTI_QS if kind == "kitti" else FORD_QS
    return [JobSpec(frame, qf(level + i), list(p), drop_last=True, lidar_level=level, pos_eps_last=False)
            for i, p in enumerate(MULLEVEL_PATHS)]


class OctreeBuilder:
    """Reusable builder (keeps its device workspace between batches)."""

    ALL_OUTPUTS = ("occ", "level", "octant", "parent", "pos", "ctx", "pos_norm", "ctx_pos", "rows_i64", "voxel_key", "sym")
    _SHAPES = {"occ": ((), torch.uint8), "level": ((), torch.uint8), "octant": ((), torch.uint8),
               "parent": ((), torch.int32), "pos": ((3,), torch.int32), "ctx": ((4, 3), torch.uint8),
               "pos_norm": ((3,), torch.float32), "ctx_pos": ((4, 3), torch.int32), "rows_i64": ((4, 6), torch.int64),
               "sym": ((), torch.int16)}

    def __init__(self):
        self.lib = _lib.require_device()
        self.h = self.lib.scp_octree_create()
        self.infos: List[JobResult] = []
        self.total_rows = 0
        self.total_voxels = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.scp_octree_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def plan(self, xyz: torch.Tensor, frame_offsets: Sequence[int], jobs: Sequence[JobSpec], mode: str):
        """xyz: CUDA float32 (n, 3|4) points of all frames back to back."""
        if not (xyz.is_cuda and xyz.dtype == torch.float32 and xyz.dim() == 2 and xyz.is_contiguous()):
            raise ValueError("xyz must be a contiguous CUDA float32 (n, 3|4) tensor")
        offs = (C.c_int64 * len(frame_offsets))(*[int(o) for o in frame_offsets])
        arr = (_lib.Job * len(jobs))()
        for a, j in zip(arr, jobs):
            a.frame = j.frame
            a.qs = float(j.qs)
            a.cart_offset = float(j.cart_offset)
            a.path_len = len(j.morton_path) if j.morton_path else 0
            a.path_bits = sum(int(b) << i for i, b in enumerate(j.morton_path or []))
            a.drop_last = int(j.drop_last)
            a.lidar_level = int(j.lidar_level)
            a.pos_eps_last = int(j.pos_eps_last)
        self._keep = xyz
        _lib.check(self.lib.scp_octree_plan(self.h, _lib.ptr(xyz), xyz.shape[1], offs, len(frame_offsets) - 1, arr,
                                            len(jobs), MODES[mode], _lib.stream_ptr()), "scp_octree_plan")
        self.n_jobs = len(jobs)
        self.total_rows = int(self.lib.scp_octree_total_rows(self.h))
        self.total_voxels = int(self.lib.scp_octree_total_voxels(self.h))
        self.total_kept = int(self.lib.scp_octree_total_kept(self.h))
        self._read_infos()
        return self

    def _read_infos(self):
        self.infos = []
        info = _lib.JobInfo()
        for j in range(self.n_jobs):
            _lib.check(self.lib.scp_octree_job_info(self.h, j, C.byref(info)), "scp_octree_job_info")
            d = info.depth
            self.infos.append(JobResult(d, info.n_points, info.n_voxels, info.n_rows, info.row_start, info.voxel_start,
                                        list(info.level_rows[:d]), float(info.bin_num), np.array(info.steps[:]),
                                        np.array(info.offset[:]),
                                        [(int(info.pos_min[l]), int(info.pos_max[l])) for l in range(d)]))

    def emit(self, outputs=("occ", "ctx", "pos_norm"), finish=True):
        """Allocates the requested outputs (torch CUDA tensors) and fills them.  Returns a dict."""
        dev = self._keep.device
        N, V = self.total_rows, self.total_voxels
        out = {}
        st = _lib.OctreeOut()
        for name in outputs:
            if name == "voxel_key":
                t = torch.empty((V,), dtype=torch.int64, device=dev)
            else:
                shape, dt = self._SHAPES[name]
                t = torch.empty((N,) + shape, dtype=dt, device=dev)
            out[name] = t
            setattr(st, name, t.data_ptr())
        _lib.check(self.lib.scp_octree_emit(self.h, C.byref(st), _lib.stream_ptr()), "scp_octree_emit")
        if finish:
            _lib.check(self.lib.scp_octree_finish(self.h, _lib.stream_ptr()), "scp_octree_finish")
            self._read_infos()
        return out

    def stage_bytes(self):
        """Algorithmic bytes of each stage of the last plan+emit, STRICTLY SURVEY.md section 8d's per-unit figures times the
        units the stage processes: quantise reads 12 B per frame point once and writes 8 B per (point, job) key; the sort
        moves (1 + 2P) * 8 B per key that takes part in it (P = ceil((3*depth+1)/8) digit passes of the 8-bit model; the
        morton_path compaction in front of it is NOT credited); tree emission ("tree") 28 B per node for everything between
        the sorted keys and the node records (head levels, emission, occupancy); context gather 60 B per node.  The
        ``heads`` / ``emit`` / ``occupancy`` entries split the tree figure over the three timers for information only
        (28 B/node is charged once, to ``tree``)."""
        n_frame_pts = int(self._keep.shape[0])
        n_keys = sum(i.n_points for i in self.infos)
        kept, N = self.total_kept, self.total_rows
        P = (3 * max(i.depth for i in self.infos) + 1 + 7) // 8
        return {"quantise": 12 * n_frame_pts + 8 * n_keys, "sort": (1 + 2 * P) * 8 * kept, "tree": 28 * N, "context": 60 * N,
                "emit": 28 * N}

    def stage_ms(self):
        arr = (C.c_float * 6)()
        _lib.check(self.lib.scp_octree_stage_ms(self.h, C.byref(arr)), "scp_octree_stage_ms")
        return dict(zip(("quantise", "sort", "heads", "emit", "occupancy", "context"), [float(x) for x in arr]))


def segmented_sort(keys: torch.Tensor, seg_offsets: Sequence[int], key_bits: int = 64):
    """In-place segmented ascending sort of a CUDA int64/uint64 tensor (bit pattern order, unsigned)."""
    lib = _lib.require_device()
    tmp = torch.empty_like(keys)
    offs = (C.c_int64 * len(seg_offsets))(*[int(o) for o in seg_offsets])
    _lib.check(lib.scp_segmented_sort_u64(_lib.ptr(keys), _lib.ptr(tmp), offs, len(seg_offsets) - 1, key_bits,
                                          _lib.stream_ptr()), "scp_segmented_sort_u64")
    return keys
