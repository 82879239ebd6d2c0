"""Drop-in for dataloaders/encode_dataset.py (``EncodeDataset`` for OctAttention, :11-55)."""
import numpy as np
import torch
import torch.utils.data as data

from .. import metrics as _metrics
from .. import octree as _oct
from ..data_preproc import pt as pointCloud
from ..synth import FORD_QS, KITTI_QS


class EncodeDataset(data.Dataset):
    def __init__(self, test_files, context_size, data_type, level_wise=True, lidar_level=12, spher=False, preproc_path=''):
        if not spher:
            raise NotImplementedError("the SCP encode path is spherical (README.md:76-86)")
        self.test_files, self.context_size, self.data_type = test_files, context_size, data_type
        self.level_wise, self.lidar_level, self.spher = level_wise, lidar_level, spher
        self.builder = None

    def __getitem__(self, index):
        pt = pointCloud.ptread(self.test_files[index])
        if self.builder is None:
            self.builder = _oct.OctreeBuilder()
        qf = KITTI_QS if self.data_type == 'kitti' else FORD_QS
        xyz = torch.from_numpy(np.ascontiguousarray(pt, dtype=np.float32)).cuda()
        b = self.builder.plan(xyz, [0, len(pt)], [_oct.JobSpec(0, qf(self.lidar_level), None)], "spher")
        out = b.emit(("rows_i64", "voxel_key"))
        rows = out["rows_i64"].cpu().numpy()
        # distChamfer(pc, quantized_pc) and the D1 PSNR of pc_error (encode_dataset.py:98-99)
        chamfer, psnr = _metrics.distortion(pt, _metrics.dequantised_cloud(b, out["voxel_key"], "spher"),
                                            _metrics.KITTI_PEAK if self.data_type == 'kitti' else _metrics.FORD_PEAK)
        cs = self.context_size
        padding = np.zeros([cs - 1, 4, 6], np.int64)
        padding[:, :, 0] = 255
        ids_pad = -np.ones([cs - 1], np.int64)
        oct_seq = rows.copy()
        oct_seq[:, :, 0] -= 1
        max_level = oct_seq[:, -1, 1].max()
        cuts = [0, len(oct_seq)]
        if self.level_wise:
            lv = oct_seq[:, -1, 1]
            cuts = [0] + list(np.flatnonzero(lv[1:] > lv[:-1]) + 1) + [len(oct_seq)]
        dat, pos, ids = [], [], []
        for a, e in zip(cuts[:-1], cuts[1:]):
            dat.append(np.vstack((padding[:, :, :3], oct_seq[a:e, :, :3])))
            pos.append(np.vstack((padding[:, :, 3:].astype(np.float32), (oct_seq[a:e, :, 3:] / (2 ** max_level)).astype(np.float32))))
            ids.append(np.hstack((ids_pad, np.arange(e - a, dtype=np.int64))))
        return ids, pos, dat, oct_seq, len(pt), int(b.infos[0].bin_num), chamfer, psnr

    def __len__(self):
        return len(self.test_files)
