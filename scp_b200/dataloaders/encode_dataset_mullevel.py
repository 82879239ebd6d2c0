"""Drop-in for dataloaders/encode_dataset_mullevel.py (``EncodeDataset`` for OctAttention over the three sub-octrees of
a frame, :12-69).  Like the reference it only works from pre-generated artefacts (``preproc_path``, written by
``data_preproc/test_gene.py --mullevel``): the reference's own constructor raises without one (:22-23) and its
``preproc`` has no other spherical branch (:106-107)."""
from pathlib import Path

import numpy as np
import torch.utils.data as data

from ..data_preproc import pt as pointCloud
from .encode_dataset import blocks_from_rows


class EncodeDataset(data.Dataset):
    def __init__(self, test_files, context_size, data_type, level_wise=True, lidar_level=12, spher=False, preproc_path=''):
        if not preproc_path:
            raise Exception('no preproc_path!')                      # encode_dataset_mullevel.py:22-23
        if not spher:
            raise NotImplementedError("the SCP encode path is spherical (README.md:76-86)")
        self.test_files, self.context_size, self.data_type = test_files, context_size, data_type
        self.level_wise, self.lidar_level, self.spher = level_wise, lidar_level, spher
        self.preproc_path = preproc_path

    def preproc(self, ori_file):
        """:94-105: row files ``<base>_0_0 / _0_1 / _1`` and ``<base>_meta.npy`` = [bin_num, chamfer(, z_offset)] (the
        reference unpacks exactly two values and fails on the three-element file its own test_gene.py:65 writes; both are
        accepted here).  PSNR is not stored (0)."""
        stem = Path(ori_file).stem
        base = self.preproc_path + ((ori_file.split('/')[-2] + stem) if self.data_type == 'kitti' else stem)
        whole_pc = pointCloud.ptread(ori_file)
        meta = np.load(base + '_meta.npy')
        return [base + '_0_0', base + '_0_1', base + '_1'], whole_pc, meta[1], int(meta[0]), 0

    def get_data(self, npy_path):
        return blocks_from_rows(np.load(npy_path + ".npy"), self.context_size, self.level_wise)

    def __getitem__(self, index):
        npy_paths, pt, chamfer, bin_num, psnr = self.preproc(self.test_files[index])
        ids, pos, dat, oct_seq = self.get_data(npy_paths[0])
        for path in npy_paths[1:]:                                   # :37-42: lists are appended, rows stacked
            cur = self.get_data(path)
            ids += cur[0]
            pos += cur[1]
            dat += cur[2]
            oct_seq = np.vstack((oct_seq, cur[3]))
        return ids, pos, dat, oct_seq, len(pt), bin_num, chamfer, psnr

    def __len__(self):
        return len(self.test_files)
