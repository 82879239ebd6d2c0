"""Drop-in for dataloaders/encode_dataset_ehem_mullevel.py (:12-85): the three sub-octrees per frame
(morton paths [0,0], [0,1], [1] at qs, ~qs/2, ~qs/4) come out of ONE batched CUDA build."""
from .. import octree as _oct
from .encode_dataset_ehem import EncodeEHEMDataset as _Base


class EncodeEHEMDataset(_Base):
    def __init__(self, test_files, context_size, data_type, level_wise=True, lidar_level=12, cylin=False, spher=False,
                 preproc_path=''):
        super().__init__(test_files, context_size, data_type, level_wise, lidar_level, cylin, spher, False, False, preproc_path)
        self._mullevel = True

    def _jobs(self):
        return _oct.mullevel_jobs(0, self.lidar_level, self.data_type), True
