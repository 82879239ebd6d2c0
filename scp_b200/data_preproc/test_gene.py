"""Drop-in for the reference's data_preproc/test_gene.py (:11-131): pre-generates, per sweep, the artefacts that
``encode*.py --preproc_path`` and ``decode*.py --preproc_path`` read instead of re-running the pre-processing --
``<name>[_0_0|_0_1|_1].npy`` ((N,4,6) int64 rows), ``<name>…_loc.npy`` (the original points), ``<name>_quant.ply`` (the
dequantised cloud) and ``<name>_meta.npy`` = [bin_num, chamfer(, z_offset)].  Transform, octree and K-parent rows come from
the CUDA pipeline (``proc_pc`` / ``mul_proc_pc`` of this package), the Chamfer distance from ``scp_nn_dist2``.

    python -m scp_b200.data_preproc.test_gene --type kitti --ori_dir 'data/11/*.bin' --out_dir pre/ --lidar_level 16 --spher --mullevel
"""
import argparse
import glob
import os
from pathlib import Path

import numpy as np

from . import pt as pointCloud
from .data_preprocess import mul_proc_pc, proc_pc
from .pt import write_ply_data

MULLEVEL_PATHS = ([0, 0], [0, 1], [1])
args = None


def get_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--type", type=str, default="kitti", choices=["kitti", "ford"])
    parser.add_argument("--ori_dir", type=str, required=True)
    parser.add_argument("--out_dir", type=str, required=True)
    parser.add_argument("--parts", type=str, default="-1/-1")
    parser.add_argument("--lidar_level", type=int, default=16)
    parser.add_argument("--cylin", action="store_true", help="whether using cylindrical coordinate")
    parser.add_argument("--spher", action="store_true", help="whether using spherical coordinate")
    parser.add_argument("--mullevel", action="store_true", help="whether using more levels for distant area")
    return parser.parse_args(argv)


def _names(ori_file, a):
    ori_path, out_dir = Path(ori_file), Path(a.out_dir)
    out_name = str(ori_path.parent).split('/')[-1] + ori_path.stem if a.type == 'kitti' else ori_path.stem
    return out_dir, out_name


def _qs(a, extra=0):
    """test_gene.py:34,45,56: 400 / (2^level - 1) for KITTI, 2^(18 - level) for Ford."""
    level = a.lidar_level + extra
    return 400 / (2 ** level - 1) if a.type == 'kitti' else 2 ** (18 - level)


def test_multi_level(ori_file, a=None):
    """:24-65 -- the three sub-octrees at qs, ~qs/2, ~qs/4."""
    a = a or args
    out_dir, out_name = _names(ori_file, a)
    res = [mul_proc_pc(ori_file, out_dir, out_name, normalize=False, qs=_qs(a, i), test=True, spher=a.spher, cylin=a.cylin,
                       morton_path=list(mp)) for i, mp in enumerate(MULLEVEL_PATHS)]
    whole_pc, bin_num, z_offset = res[0][2], res[0][3], res[0][4]
    whole_q_pc = np.vstack([r[1] for r in res])
    write_ply_data(out_dir / (out_name + "_quant.ply"), whole_q_pc)
    np.save(out_dir / (out_name + '_meta'), [bin_num, pointCloud.distChamfer(whole_pc, whole_q_pc), z_offset])


def test(ori_file, a=None):
    """:68-87 -- single spherical octree."""
    a = a or args
    out_dir, out_name = _names(ori_file, a)
    out_file, quantized_pc, pc, bin_num = proc_pc(ori_file, out_dir, out_name, normalize=False, qs=_qs(a), test=True,
                                                  spher=a.spher)[:4]
    whole_pc = pointCloud.ptread(ori_file)
    write_ply_data(out_dir / (out_name + "_quant.ply"), quantized_pc)
    np.save(out_dir / (out_name + '_meta'), [bin_num, pointCloud.distChamfer(whole_pc, quantized_pc)])


def test_cylin(ori_file, a=None):
    """:90-106 -- single cylindrical octree; the meta file also carries the z offset."""
    a = a or args
    out_dir, out_name = _names(ori_file, a)
    out_file, quantized_pc, pc, bin_num, offset = proc_pc(ori_file, out_dir, out_name, normalize=False, qs=_qs(a), test=True,
                                                          cylin=a.cylin)
    whole_pc = pointCloud.ptread(ori_file)
    write_ply_data(out_dir / (out_name + "_quant.ply"), quantized_pc)
    np.save(out_dir / (out_name + '_meta'), [bin_num, pointCloud.distChamfer(whole_pc, quantized_pc), offset[0, 2]])


def main(a):
    """:109-131 (``--parts i/n`` processes the i-th of n slices of the file list)."""
    global args
    args = a
    if not os.path.exists(a.out_dir):
        os.mkdir(a.out_dir)
    test_files = glob.glob(a.ori_dir)
    part, total = (0, 1) if a.parts.startswith("-1") else (int(a.parts.split("/")[0]), int(a.parts.split("/")[1]))
    start = len(test_files) * part // total
    end = len(test_files) * (part + 1) // total
    for i, ori_file in enumerate(test_files[start:end]):
        if a.mullevel:
            test_multi_level(ori_file, a)
        elif a.cylin:
            test_cylin(ori_file, a)
        else:
            test(ori_file, a)
        print(f"part {part}/{total}: {i}/{end-start}")


if __name__ == '__main__':
    main(get_args())
