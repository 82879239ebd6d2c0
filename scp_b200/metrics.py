"""Distortion report next to bpp: the dequantised cloud of an octree build and the two
nearest-neighbour statistics the reference prints per frame (encode.py:283-291) -- ``distChamfer`` (pt.py:88-95, two
scipy KDTree queries) and the symmetric D1 PSNR it parses out of MPEG's ``pc_error`` tool (pt.py:13-84,
utils/__init__.py:3-16; the binary is not part of the reference repository).  Both are exact brute-force searches on the
GPU (``scp_nn_dist2``); there is no CPU path."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .octree import MODES

KITTI_PEAK = 59.70      # "-r" argument of pc_error for KITTI (encode_dataset_ehem.py:115) ...
FORD_PEAK = 30000.0     # ... and Ford (:117)


def _dev64(a):
    """numpy / torch (n,3) points -> contiguous CUDA float64 (exact for float32 inputs)."""
    t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a)
    if t.dim() != 2 or t.shape[1] != 3:
        raise ValueError("points must be (n, 3)")
    return t.to(device="cuda", dtype=torch.float64).contiguous()


def nn_dist2(query, cand):
    """Squared distance of every ``query`` point to its nearest ``cand`` point: CUDA float64 [n_query]."""
    lib = _lib.require_device()
    q, c = _dev64(query), _dev64(cand)
    if c.shape[0] == 0:
        raise ValueError("nn_dist2: empty candidate set")
    out = torch.empty((q.shape[0],), dtype=torch.float64, device=q.device)
    if q.shape[0] == 0:
        return out
    _lib.check(lib.scp_nn_dist2(_lib.ptr(q), q.shape[0], _lib.ptr(c), c.shape[0], _lib.ptr(out), _lib.stream_ptr()),
               "scp_nn_dist2")
    return out


def distChamfer(f1, f2, scale=1.0):
    """pt.py:88-95: ``max(mean_j min_i |f2_j - f1_i|, mean_i min_j |f1_i - f2_j|)``.  Like the reference, ``scale`` divides
    both clouds first (every caller passes 1.0); unlike it, the caller's arrays are not modified in place."""
    a, b = _dev64(f1), _dev64(f2)
    if scale != 1.0:
        a, b = a / scale, b / scale
    d1 = nn_dist2(b, a).sqrt().mean()
    d2 = nn_dist2(a, b).sqrt().mean()
    return float(torch.maximum(d1, d2))


def d1_psnr(ref, deg, peak):
    """Symmetric point-to-point (D1) error of pc_error: ``mse = max(mean nn_dist2(ref->deg), mean nn_dist2(deg->ref))``,
    ``psnr = 10 log10(3 peak^2 / mse)`` -- the "mseF,PSNR (p2point)" line of its section 3 that ``get_psnr`` reads
    (utils/__init__.py:8-9).  Returns (mse, psnr)."""
    a, b = _dev64(ref), _dev64(deg)
    mse = float(torch.maximum(nn_dist2(a, b).mean(), nn_dist2(b, a).mean()))
    return mse, (math.inf if mse == 0.0 else 10.0 * math.log10(3.0 * peak * peak / mse))


def distortion_terms(ref, deg):
    """CUDA float64 [2] = (Chamfer distance, symmetric mean squared nearest-neighbour distance) from ONE pair of
    nearest-neighbour passes, without synchronising the stream (``Encoder(distortion=True)`` reads it back with the
    intervals)."""
    a, b = _dev64(ref), _dev64(deg)
    ab, ba = nn_dist2(a, b), nn_dist2(b, a)
    return torch.stack((torch.maximum(ba.sqrt().mean(), ab.sqrt().mean()), torch.maximum(ab.mean(), ba.mean())))


def psnr_of(mse, peak):
    return math.inf if mse == 0.0 else 10.0 * math.log10(3.0 * peak * peak / mse)


def distortion(ref, deg, peak):
    """Chamfer distance and D1 PSNR of a frame.  Returns (chamfer, psnr)."""
    chamfer, mse = distortion_terms(ref, deg).tolist()
    return chamfer, psnr_of(mse, peak)


def dequantise_keys(keys, steps, offset, mode):
    """Voxel Morton keys (CUDA int64 [n], ``OctreeBuilder.emit(("voxel_key",))``) -> CUDA float64 [n,3] Cartesian points:
    ``v * steps + offset`` then spher2cart / cylin2cart (data_preprocess.py:68-92, :160-167, :179-229)."""
    lib = _lib.require_device()
    if not (keys.is_cuda and keys.dtype == torch.int64 and keys.dim() == 1 and keys.is_contiguous()):
        raise ValueError("keys must be a contiguous CUDA int64 vector")
    st = (C.c_double * 3)(*[float(x) for x in np.asarray(steps).reshape(3)])
    of = (C.c_double * 3)(*[float(x) for x in np.asarray(offset).reshape(3)])
    out = torch.empty((keys.shape[0], 3), dtype=torch.float64, device=keys.device)
    if keys.shape[0] == 0:
        return out
    _lib.check(lib.scp_dequantise_keys(_lib.ptr(keys), keys.shape[0], st, of, MODES[mode], _lib.ptr(out),
                                       _lib.stream_ptr()), "scp_dequantise_keys")
    return out


def dequantised_cloud(builder, voxel_key, mode):
    """The quantised cloud of every job of ``builder`` (an ``OctreeBuilder``, or a list of its ``JobResult``s) back to back
    (the ``np.vstack`` of encode_dataset_ehem_mullevel.py:141,188): spherical jobs carry no offset, cylindrical ones their
    z offset."""
    parts = []
    for i in getattr(builder, "infos", builder):
        off = np.zeros(3) if mode == "spher" else i.offset
        parts.append(dequantise_keys(voxel_key[i.voxel_start:i.voxel_start + i.n_voxels], i.steps, off, mode))
    return torch.cat(parts, 0)
