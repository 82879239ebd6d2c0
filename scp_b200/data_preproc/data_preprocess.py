"""Drop-in for the reference's data_preproc/data_preprocess.py on the encode path: same function names,
arguments and return values (``proc_pc`` :13-92, ``mul_proc_pc`` :95-167, coordinate helpers :171-229), with
the transform / quantise / octree / K-parent work done by the CUDA pipeline (scp_octree_plan / scp_octree_emit)
instead of numpy + Octree_python_lib.so + Python loops."""
import math
import os

import numpy as np
import torch

from .. import octree as _oct
from . import pt as pointCloud

MVUB_NAMES = ['andrew10', 'david10', 'phil10', 'phil9', 'ricardo10', 'ricardo9', 'sarah10']


def _dequant_from_keys(keys, steps, offset):
    """voxel Morton keys -> (V,3) quantised integer voxels (x,y,z per key bit triple)."""
    k = keys.astype(np.uint64)
    q = np.zeros((len(k), 3), np.int64)
    for b in range(21):
        for c in range(3):
            q[:, c] |= (((k >> np.uint64(3 * b + (2 - c))) & np.uint64(1)).astype(np.int64)) << b
    return q


def _build(p, qs, mode, morton_path, drop_last, cart_offset=0.0):
    xyz = torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32)).cuda()
    job = _oct.JobSpec(0, float(qs), list(morton_path) if morton_path is not None else None, drop_last=drop_last,
                       cart_offset=float(cart_offset))
    b = _oct.OctreeBuilder().plan(xyz, [0, len(p)], [job], mode)
    out = b.emit(("rows_i64", "voxel_key"))
    return b.infos[0], out["rows_i64"].cpu().numpy(), out["voxel_key"].cpu().numpy()


def _common(inp_path, out_dir, normalize, rotation):
    if not os.path.exists(out_dir):
        os.makedirs(out_dir)
    p = pointCloud.ptread(inp_path)
    ref_pt = p
    if normalize is True:
        p = p - np.mean(p, axis=0)
        p = p / abs(p).max()
        ref_pt = p
    if rotation:
        ref_pt = ref_pt[:, [0, 2, 1]]
        ref_pt[:, 2] = -ref_pt[:, 2]
    return ref_pt


def proc_pc(inp_path, out_dir, out_name, qs=1, offset='min', qlevel=None, rotation=False, normalize=False,
            test=False, cylin=False, spher=False):
    """data_preprocess.py:13-92.  Writes ``<out>.npy`` ((N,4,6) int64) and, with ``test``, ``<out>_loc.npy``."""
    if qlevel is not None:
        raise NotImplementedError("qlevel is not used by the encode path (encode_dataset_ehem.py:136-181)")
    ref_pt = _common(inp_path, out_dir, normalize, rotation)
    mode = "cylin" if cylin else ("spher" if spher else "cart")
    if mode == "cart":
        if isinstance(offset, str):
            raise NotImplementedError("offset='min' is not used by the encode path; pass a scalar like the callers do")
        info, rows, keys = _build(ref_pt, qs, mode, None, False, cart_offset=offset)
    else:
        info, rows, keys = _build(ref_pt, qs, mode, None, False)
    if test:
        out_file = os.path.join(out_dir, out_name)
        np.save(out_file + "_loc", ref_pt)
    else:
        out_file = os.path.join(out_dir, out_name + "_" + str(rows.shape[0]))
    np.save(out_file, rows)
    if not test:
        return
    pt = np.unique(_dequant_from_keys(keys, info.steps, info.offset), axis=0)       # np.unique order of :69
    bin_num = np.float32(info.bin_num)
    off = info.offset[None] if mode != "spher" else 0
    out_points = (pt * info.steps[None] + off).astype(np.float32)
    if cylin:
        return [out_file, cylin2cart(out_points), ref_pt, bin_num, info.offset[None]]
    if spher:
        return [out_file, spher2cart(out_points), ref_pt, bin_num]
    return [out_file, out_points, ref_pt]


def mul_proc_pc(inp_path, out_dir, out_name, qs=1, offset=0, qlevel=None, rotation=False, normalize=False, test=False,
                cylin=False, spher=False, morton_path=[0]):
    """data_preprocess.py:95-167 (multi-level: one sub-octree selected by ``morton_path``)."""
    if qlevel is not None:
        raise NotImplementedError("qlevel is not used by the encode path")
    if not (cylin or spher):
        raise NotImplementedError("mul_proc_pc is only called with cylin/spher (encode_dataset_ehem_mullevel.py:110-186)")
    ref_pt = _common(inp_path, out_dir, normalize, rotation)
    mode = "cylin" if cylin else "spher"
    info, rows, keys = _build(ref_pt, qs, mode, morton_path, True)
    if test:
        for m in morton_path:
            out_name += '_' + str(m)
        out_file = os.path.join(out_dir, out_name)
        np.save(out_file + '_loc', ref_pt)
    else:
        out_file = os.path.join(out_dir, out_name + '_' + str(rows.shape[0]))
    np.save(out_file, rows)
    pt = _dequant_from_keys(keys, info.steps, info.offset)          # DeOctree(codes) order == Morton order (:160)
    off = info.offset[None] if cylin else offset
    out_points = pt * info.steps[None] + off
    bin_num = np.float32(info.bin_num)
    if cylin:
        return [out_file, cylin2cart(out_points), ref_pt, bin_num, info.offset[2]]
    return [out_file, spher2cart(out_points), ref_pt, bin_num, offset]


def cart2cylin(points):
    """data_preprocess.py:171-177 (host helper; the encode path does this on the GPU)."""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    rho = np.sqrt(x ** 2 + y ** 2)
    phi = np.arctan2(y, x + 1e-9)
    phi[np.where(phi < 0)[0]] += 2 * math.pi
    return np.vstack((rho, phi, z)).transpose(1, 0)


def cylin2cart(points):
    rho, phi, z = points[:, 0], points[:, 1], points[:, 2]
    return np.vstack((rho * np.cos(phi), rho * np.sin(phi), z)).transpose(1, 0)


def cart2spher(points):
    """data_preprocess.py:200-207"""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    rho = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    phi = np.arctan2(y, x + 1e-9)
    phi[np.where(phi < 0)[0]] += 2 * math.pi
    theta = np.arccos(z / rho)
    return np.vstack((rho, phi, theta)).transpose(1, 0)


def spher2cart(points):
    rho, phi, theta = points[:, 0], points[:, 1], points[:, 2]
    return np.vstack((rho * np.sin(theta) * np.cos(phi), rho * np.sin(theta) * np.sin(phi), rho * np.cos(theta))).transpose(1, 0)
