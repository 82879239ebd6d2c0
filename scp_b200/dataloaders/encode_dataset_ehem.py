"""Drop-in for dataloaders/encode_dataset_ehem.py (``EncodeEHEMDataset``, :12-105): same constructor and the same
``__getitem__`` tuple, with pre-processing + level split done by the CUDA octree pipeline instead of
``proc_pc`` + ``np.load`` + the per-node Python loop (:64-83)."""
import numpy as np
import torch
import torch.utils.data as data

from .. import metrics as _metrics
from .. import octree as _oct
from ..data_preproc import pt as pointCloud
from ..synth import FORD_QS, KITTI_QS


def _split_levels(rows, ctx, posn, infos):
    """Per-level lists in the reference's dtypes/shapes from the batch outputs."""
    ids, poss, pos_mm, dat = [], [], [], []
    for i in infos:
        r = i.row_start
        for n, mm in zip(i.level_rows, i.pos_mm):
            dat.append(ctx[r:r + n].astype(np.int64))
            poss.append(np.ascontiguousarray(posn[r:r + n].T))
            pos_mm.append(mm)
            ids.append(np.arange(n, dtype=np.int64))
            r += n
    return ids, poss, pos_mm, dat


class EncodeEHEMDataset(data.Dataset):
    def __init__(self, test_files, context_size, data_type, level_wise=True, lidar_level=12, cylin=False, spher=False,
                 circle=False, extra_pos=False, preproc_path=''):
        if extra_pos:
            raise NotImplementedError("extra_pos is dead code in the reference encode path (encode.py:164-233)")
        if not (cylin or spher):
            raise NotImplementedError("the SCP encode path is spherical/cylindrical (README.md:76-86)")
        self.test_files = test_files
        self.context_size = context_size
        self.data_type = data_type
        self.level_wise = level_wise
        self.lidar_level = lidar_level
        self.cylin = cylin
        self.spher = spher
        self.builder = None
        self._mullevel = False

    def _jobs(self):
        qf = KITTI_QS if self.data_type == 'kitti' else FORD_QS
        return [_oct.JobSpec(0, qf(self.lidar_level), None, lidar_level=self.lidar_level)], False

    def _oct_seq(self, rows, infos):
        """np.load(.npy) then ``[:, :, 0] -= 1`` and the in-place last-block level clip (:54,:86)."""
        seq = rows.copy()
        seq[:, :, 0] -= 1
        for i in infos:
            n_last = i.level_rows[-1]
            e = i.row_start + i.n_rows
            seq[e - n_last:e, :, 1] = np.minimum(seq[e - n_last:e, :, 1], self.lidar_level)
        return seq

    def _distortion(self, pc, b, voxel_key):
        """``distChamfer(pc, quantized_pc)`` and pc_error's D1 PSNR (:147, :170-171; mullevel :141-144): original cloud
        against the dequantised voxels of all jobs.  The single-level cylindrical branch reports PSNR 0 (:146-147)."""
        mode = "cylin" if self.cylin else "spher"
        q_pc = _metrics.dequantised_cloud(b, voxel_key, mode)
        peak = _metrics.KITTI_PEAK if self.data_type == 'kitti' else _metrics.FORD_PEAK
        chamfer, psnr = _metrics.distortion(pc, q_pc, peak)
        return chamfer, (0 if self.cylin and not self._mullevel else psnr)

    def __getitem__(self, index):
        pc = pointCloud.ptread(self.test_files[index])
        if self.builder is None:
            self.builder = _oct.OctreeBuilder()
        jobs, _ = self._jobs()
        xyz = torch.from_numpy(np.ascontiguousarray(pc, dtype=np.float32)).cuda()
        b = self.builder.plan(xyz, [0, len(pc)], jobs, "cylin" if self.cylin else "spher")
        out = b.emit(("rows_i64", "ctx", "pos_norm", "voxel_key"))
        rows, ctx, posn = (out[k].cpu().numpy() for k in ("rows_i64", "ctx", "pos_norm"))
        ids, poss, pos_mm, dat = _split_levels(rows, ctx, posn, b.infos)
        bin_num = int(b.infos[0].bin_num)
        z_offset = float(b.infos[0].offset[2]) if self.cylin else 0
        chamfer, psnr = self._distortion(pc, b, out["voxel_key"])
        return ids, poss, pos_mm, dat, self._oct_seq(rows, b.infos), len(pc), pc, bin_num, z_offset, chamfer, psnr

    def __len__(self):
        return len(self.test_files)
