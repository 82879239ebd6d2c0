"""End-to-end encode path: LiDAR points -> bitstream, for batches of frames on one GPU.

Replaces the per-frame, per-window Python loops of the reference's ``encode.py`` (compress_ehem :85-160,
compress :23-82) and ``encode_mullevel.py`` (:88-157, :23-85) together with the dataset/pre-processing they
drive (dataloaders/encode_dataset_ehem*.py, data_preproc/data_preprocess.py):

  points (host or device) --> octree rows / context bytes / normalised positions   (scp_octree_*)
                          --> all context windows of all frames as ONE ragged batch  (scp_gather_windows)
                          --> EHEM / OctAttention logits                              (models.*)
                          --> softmax -> integer CDF -> (c_low, c_high) of the true symbol   (scp_pmf_to_cdf)
                          --> coding order gather                                     (scp_coding_order, scp_gather_rows8)
                          --> range coder, one stream per frame                       (scp_range_encode, host threads)

Only 8 bytes per node leave the GPU (the reference copies 1020 B/node of PMFs, encode.py:133).
"""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _lib, coder, metrics, octree
from .synth import FORD_QS, KITTI_QS


@dataclass
class FrameResult:
    n_points: int
    n_nodes: int
    n_levels: int
    bin_num: int
    z_offset: float
    bitstream: bytes
    bpp: float
    pos_mm: list = field(default_factory=list)      # (min, max) of every level's pos block: the .dat side file (encode.py:150)
    depths: list = field(default_factory=list)      # levels per sub-octree (one entry, or three for encode_mullevel)
    chamfer: Optional[float] = None                 # distChamfer(points, dequantised voxels) and pc_error's D1 PSNR,
    psnr: Optional[float] = None                    # only with Encoder(distortion=True) (encode.py:288-291 prints them)


class Encoder:
    """``Encoder(model, lidar_level, mode, mullevel)`` then ``encode(frames)``.

    mode 'spher' | 'cylin' | 'cart'; ``mullevel`` = the three sub-octrees of encode_mullevel.py;
    ``kind`` 'kitti' | 'ford' picks the quantisation step like encode_dataset_ehem.py:164.  ``distortion=True`` adds the
    reference's per-frame distortion report (Chamfer distance, D1 PSNR; two exact nearest-neighbour passes, ~21 ms for a
    120 k-point frame) to every ``FrameResult``; it is off by default because it is not part of the bitstream path."""

    def __init__(self, model, lidar_level=12, mode="spher", mullevel=False, kind="kitti", max_tokens=1 << 19,
                 coder_threads=None, distortion=False):
        self.model = model
        self.level = lidar_level
        self.mode = mode
        self.mullevel = mullevel
        self.kind = kind
        self.distortion = distortion
        self.is_ehem = model.__class__.__name__ == "EHEM"
        self.context = model.cfg.model.context_size
        self.max_tokens = int(os.environ.get("SCP_MAX_TOKENS", max_tokens))    # tokens per ragged model call
        self.builder = octree.OctreeBuilder()
        self.lib = _lib.require_device()
        self.pool = ThreadPoolExecutor(coder_threads or min(32, os.cpu_count() or 1))
        self.timings = {}
        self._pin, self._pin_i = {}, {}

    # -- stage 1: octrees ---------------------------------------------------------------------
    def _jobs(self, n_frames):
        qf = KITTI_QS if self.kind == "kitti" else FORD_QS
        clip = self.level if self.is_ehem else 255          # encode_dataset*.py (OctAttention) do not clip levels
        if self.mullevel:
            jobs = [j for f in range(n_frames) for j in octree.mullevel_jobs(f, self.level, self.kind)]
            for j in jobs:
                j.lidar_level = clip
            return jobs, 3
        return [octree.JobSpec(f, qf(self.level), None, lidar_level=clip) for f in range(n_frames)], 1

    def build_context(self, xyz, frame_offsets):
        jobs, per_frame = self._jobs(len(frame_offsets) - 1)
        b = self.builder.plan(xyz, frame_offsets, jobs, self.mode)
        outs = ("occ", "sym", "ctx", "pos_norm") if self.is_ehem else ("occ", "sym", "ctx", "ctx_pos")
        if self.distortion:
            outs += ("voxel_key",)
        t = b.emit(outs, finish=False)
        return b, t, per_frame

    # -- stage 2: windows ---------------------------------------------------------------------
    def _ehem_windows(self, infos):
        """(row, len, token) of every context window (encode.py:112-115), level by level."""
        rows, lens = [], []
        r = infos[0].row_start if infos and hasattr(infos[0], "row_start") else 0
        sizes = [n for i in infos for n in i.level_rows] if infos and hasattr(infos[0], "level_rows") else list(infos)
        for n in sizes:
            for s in range(0, n, self.context):
                rows.append(r + s)
                lens.append(min(self.context, n - s))
            r += n
        lens = np.asarray(lens, np.int32)
        toks = np.concatenate([[0], np.cumsum(lens + (lens & 1))]).astype(np.int64)
        return np.asarray(rows, np.int64), lens, toks

    def _ehem_logits_to_intervals(self, t, infos, interval_row):
        rows, lens, toks = self._ehem_windows(infos)
        ctx, pos, sym = t["ctx"], t["pos_norm"], t["sym"]
        dev = ctx.device
        w0 = 0
        st = _lib.stream_ptr()
        # chunks of windows bounded by max_tokens, of EQUAL size: a greedy split leaves a tail chunk of a few thousand tokens
        # (a K16 frame is 529 k tokens against 2^19) whose ~850 launches run latency-bound
        n_chunks = max(1, -(-int(toks[-1]) // self.max_tokens))
        target = -(-int(toks[-1]) // n_chunks)
        while w0 < len(rows):
            w1 = w0 + 1
            while w1 < len(rows) and toks[w1 + 1] - toks[w0] <= min(self.max_tokens, target + self.context):
                w1 += 1
            T = int(toks[w1] - toks[w0])
            ctxp = torch.empty((T, 4, 3), dtype=torch.uint8, device=dev)
            posp = torch.empty((T, 3), dtype=torch.float32, device=dev)
            re = torch.empty(T // 2, dtype=torch.int64, device=dev)
            ro = torch.empty(T // 2, dtype=torch.int64, device=dev)
            tk = np.ascontiguousarray(toks[w0:w1] - toks[w0])
            _lib.check(self.lib.scp_gather_windows(_lib.ptr(ctx), _lib.ptr(pos), _lib.ptr(np.ascontiguousarray(rows[w0:w1])),
                                                   _lib.ptr(np.ascontiguousarray(lens[w0:w1])), _lib.ptr(tk), w1 - w0,
                                                   _lib.ptr(ctxp), _lib.ptr(posp), _lib.ptr(re), _lib.ptr(ro), st),
                       "scp_gather_windows")
            l1, l2 = self.model.forward_ragged(ctxp, posp, [int(x) for x in np.append(tk, T)])
            coder.pmf_to_cdf(l1, sym=sym, is_logits=True, row_of=re, out={"interval": interval_row})
            coder.pmf_to_cdf(l2, sym=sym, is_logits=True, row_of=ro, out={"interval": interval_row})
            w0 = w1

    def _octattn_logits_to_intervals(self, t, infos, per_frame, interval_row):
        """encode.py:23-82 (non level-wise) / encode_mullevel.py:23-85: every row file (the frame's octree, or each of its
        three sub-octrees) is one sequence [1023 pad rows ; nodes] cut into windows of 1024; positions are divided by
        2^(deepest level of that file) (encode_dataset.py:38,48).  All sequences of all frames of the batch are assembled by
        one launch (scp_pad_gather_seqs) and run through the model as ragged batches of windows."""
        ctx, cpos, sym = t["ctx"], t["ctx_pos"], t["sym"]
        dev = ctx.device
        cs = self.context
        lens = np.array([i.n_rows + cs - 1 for i in infos], np.int32)
        src = np.array([i.row_start for i in infos], np.int64)
        dst = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        # deepest level present in the file's rows (the mullevel files lose their last row, Octree.py:259-262)
        deepest = [max(l + 1 for l, n in enumerate(i.level_rows) if n > 0) if i.n_rows else 1 for i in infos]
        shift = np.array([21 - d for d in deepest], np.int32)
        T = int(dst[-1])
        seq_ctx = torch.empty((T, 4, 3), dtype=torch.uint8, device=dev)
        seq_pos = torch.empty((T, 4, 3), dtype=torch.int32, device=dev)
        row_of = torch.empty((T,), dtype=torch.int64, device=dev)
        _lib.check(self.lib.scp_pad_gather_seqs(_lib.ptr(ctx), _lib.ptr(cpos), _lib.ptr(dst), _lib.ptr(src), _lib.ptr(lens),
                                                _lib.ptr(shift), len(infos), cs - 1, _lib.ptr(seq_ctx), _lib.ptr(seq_pos),
                                                _lib.ptr(row_of), _lib.stream_ptr()), "scp_pad_gather_seqs")
        offs = [0]
        for a, n in zip(dst[:-1], lens):
            offs += [int(a) + min(w + cs, int(n)) for w in range(0, int(n), cs)]
        w0 = 0
        while w0 < len(offs) - 1:
            w1 = w0 + 1
            while w1 < len(offs) - 1 and offs[w1 + 1] - offs[w0] <= self.max_tokens:
                w1 += 1
            lo, hi = offs[w0], offs[w1]
            logits = self.model.forward_ragged(seq_ctx[lo:hi], seq_pos[lo:hi], [o - lo for o in offs[w0:w1 + 1]],
                                               1.0 / float(1 << 21))
            coder.pmf_to_cdf(logits, sym=sym, is_logits=True, row_of=row_of[lo:hi], out={"interval": interval_row})
            w0 = w1

    @torch.no_grad()
    def encode_context(self, ctx, pos, level_sizes, level_restart=None):
        """Entropy-model + CDF + coding order for context tensors that already exist (the ``batch`` a reference
        ``EncodeEHEMDataset`` yields): ctx uint8 [N,4,3] (level, octant, occ 0..254|255), pos float32 [N,3].
        Returns the (c_low, c_high) intervals in coding order, int32 [N,2] on the device."""
        assert self.is_ehem
        N = ctx.shape[0]
        occ = (ctx[:, 3, 2].to(torch.int16) + 1).to(torch.uint8).contiguous()
        t = {"ctx": ctx.contiguous(), "pos_norm": pos.contiguous(), "sym": ctx[:, 3, 2].to(torch.int16).contiguous()}
        interval_row = torch.empty((N, 2), dtype=torch.int32, device=ctx.device)
        self._ehem_logits_to_intervals(t, [int(n) for n in level_sizes], interval_row)
        order, _ = coder.coding_order([int(n) for n in level_sizes], self.context, occ, mullevel=self.mullevel,
                                      level_restart=level_restart)
        interval = torch.empty_like(interval_row)
        _lib.check(self.lib.scp_gather_rows8(_lib.ptr(interval_row), _lib.ptr(order), N, _lib.ptr(interval),
                                             _lib.stream_ptr()), "scp_gather_rows8")
        return interval

    # -- public API ---------------------------------------------------------------------------
    def _encode_from_context(self, b, t, per_frame, dev):
        infos = b.infos
        N = b.total_rows
        interval_row = torch.empty((N, 2), dtype=torch.int32, device=dev)
        if self.is_ehem:
            self._ehem_logits_to_intervals(t, infos, interval_row)
            sizes, restart = [], []
            for j, i in enumerate(infos):
                for l, n in enumerate(i.level_rows):
                    sizes.append(n)
                    restart.append(1 if (j % per_frame == 0 and l == 0) else 0)
            order, _ = coder.coding_order(sizes, self.context, t["occ"], mullevel=self.mullevel, level_restart=restart)
            interval = torch.empty_like(interval_row)
            _lib.check(self.lib.scp_gather_rows8(_lib.ptr(interval_row), _lib.ptr(order), N, _lib.ptr(interval),
                                                 _lib.stream_ptr()), "scp_gather_rows8")
        else:
            self._octattn_logits_to_intervals(t, infos, per_frame, interval_row)
            interval = interval_row                              # OctAttention codes in BFS order (encode.py:67-69)
        frames = []
        for f in range(len(infos) // per_frame):
            fi = infos[f * per_frame:(f + 1) * per_frame]
            frames.append((fi[0].row_start, sum(i.n_rows for i in fi)))
        return interval, frames

    @torch.no_grad()
    def encode_device(self, xyz, frame_offsets):
        """xyz: CUDA float32 (n,3|4); returns (interval tensor int32 [N,2] in coding order on the device,
        per-frame (row_start, n_rows), job infos)."""
        b, t, per_frame = self.build_context(xyz, frame_offsets)
        interval, frames = self._encode_from_context(b, t, per_frame, xyz.device)
        return interval, frames, b.infos, per_frame

    # -- host-side pipeline -----------------------------------------------------------------
    def _pinned(self, kind, shape, dtype):
        """Rotating pinned host buffers (three per kind): pinning a fresh 4-8 MB array per batch costs milliseconds."""
        ring = self._pin.setdefault(kind, [])
        n = int(np.prod(shape))
        slot = self._pin_i.get(kind, 0)
        self._pin_i[kind] = (slot + 1) % 3
        if len(ring) <= slot:
            ring.append(None)
        if ring[slot] is None or ring[slot].numel() < n or ring[slot].dtype != dtype:
            ring[slot] = torch.empty(max(n, 1), dtype=dtype).pin_memory()
        return ring[slot][:n].view(shape)

    def _gpu_stage(self, frames_xyz):
        """Enqueues everything a batch needs on the GPU (H2D, octree, entropy model, intervals, D2H) and returns a handle;
        only the octree planning synchronises (the window layout is needed on the host)."""
        offs = np.concatenate([[0], np.cumsum([len(f) for f in frames_xyz])]).astype(np.int64)
        cols = frames_xyz[0].shape[1]
        host = self._pinned("xyz", (int(offs[-1]), cols), torch.float32)
        hn = host.numpy()
        for f, a, b in zip(frames_xyz, offs[:-1], offs[1:]):
            hn[a:b] = f
        xyz = host.cuda(non_blocking=True)
        b, t, per_frame = self.build_context(xyz, offs)
        _lib.check(self.lib.scp_octree_finish(self.builder.h, _lib.stream_ptr()), "scp_octree_finish")   # level min/max: final after emit
        self.builder._read_infos()
        infos = list(self.builder.infos)
        interval, frames = self._encode_from_context(b, t, per_frame, xyz.device)
        iv = self._pinned("iv", tuple(interval.shape), torch.int32)
        iv.copy_(interval, non_blocking=True)
        dist = None
        if self.distortion:                                      # (chamfer, mse) per frame, read back with the intervals
            terms = [metrics.distortion_terms(xyz[offs[f]:offs[f + 1], :3],
                                              metrics.dequantised_cloud(infos[f * per_frame:(f + 1) * per_frame],
                                                                        t["voxel_key"], self.mode))
                     for f in range(len(offs) - 1)]
            dist = self._pinned("dist", (len(terms), 2), torch.float64)
            dist.copy_(torch.stack(terms), non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        return {"offs": offs, "infos": infos, "per_frame": per_frame, "frames": frames, "iv": iv, "done": done,
                "dist": dist, "keep": (xyz, interval, t)}

    def _launch_coder(self, st):
        st["done"].synchronize()
        ivn = st["iv"].numpy().view(np.uint32)
        st["futures"] = [self.pool.submit(coder.range_encode, ivn[r0:r0 + n]) for r0, n in st["frames"]]
        st["keep"] = None
        if st["dist"] is not None:
            st["dist"] = st["dist"].tolist()                     # the pinned block is recycled three batches later
        return st

    def _collect(self, st):
        out = []
        offs, infos, per_frame = st["offs"], st["infos"], st["per_frame"]
        for f, ((r0, n), fut) in enumerate(zip(st["frames"], st["futures"])):
            bs = fut.result()
            fi = infos[f * per_frame:(f + 1) * per_frame]
            npts = int(offs[f + 1] - offs[f])
            out.append(FrameResult(npts, n, sum(len(i.level_rows) for i in fi), int(fi[0].bin_num),
                                   float(fi[0].offset[2]), bs, 8.0 * len(bs) / npts,
                                   [p for i in fi for p in i.pos_mm], [len(i.level_rows) for i in fi]))
            if st["dist"] is not None:
                out[-1].chamfer = st["dist"][f][0]
                out[-1].psnr = metrics.psnr_of(st["dist"][f][1], metrics.KITTI_PEAK if self.kind == "kitti" else metrics.FORD_PEAK)
        return out

    @torch.no_grad()
    def encode(self, frames_xyz: List[np.ndarray]) -> List[FrameResult]:
        """frames_xyz: list of host float32 (n,3|4) arrays (KITTI .bin rows).  Returns one FrameResult per frame."""
        return self._collect(self._launch_coder(self._gpu_stage(frames_xyz)))

    @torch.no_grad()
    def encode_stream(self, batches):
        """Generator: ``batches`` yields lists of host frames; yields the list of FrameResult of each batch, in order.
        Pipelined over batches: while the GPU encodes batch i+1, the host range coder (one thread per frame, the GIL is
        released inside the C call) writes the bitstreams of batch i, so throughput is bound by the GPU alone."""
        staged, coding = None, None
        for frames_xyz in batches:
            cur = self._gpu_stage(frames_xyz)          # its octree sync also waits for the previous batch's GPU work
            if staged is not None:
                if coding is not None:
                    yield self._collect(coding)
                coding = self._launch_coder(staged)
            staged = cur
        if staged is not None:
            if coding is not None:
                yield self._collect(coding)
            coding = self._launch_coder(staged)
        if coding is not None:
            yield self._collect(coding)
