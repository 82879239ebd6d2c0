"""Drop-in for the reference's ``decode_ehem.py``: same ``extract_info(file)`` (:20-27) and
``decodeOct(binfile, oct_data_seq, model, context_size, anc_k)`` (:56-188) contracts on the files ``compress_ehem``
writes (``<name>[_spher|_cylin]_<levels>_<bin_num>_<z_offset>.bin`` + ``.bin.dat``).  The per-node Python
(`cur_nodes` / `pre_nodes` deques, torch child expansion, numpyAc calls per window) is replaced by
scp_b200.decoder.Decoder: level-wise batches on the GPU, host range decoder in the library."""
import time
import types

import numpy as np
import torch

from .decoder import Decoder, dequantise

MULLEVEL = False


def extract_info(file):
    """decode_ehem.py:20-27: (spher, cylin, pos_mm, max_level, bin_num, z_offset) from the file name and the .dat."""
    spher = 'spher' in file
    cylin = 'cylin' in file
    rtn = list(map(lambda x: int(x), file.split('/')[-1].split('.')[0].split('_')[-3:]))
    pos_mm = torch.load(file + '.dat')
    return [spher, cylin, pos_mm] + rtn


def sub_depths(max_level, mullevel):
    """decode_ehem_mullevel.py:190-199: the file name carries the total number of levels; the three sub-octrees of
    encode_mullevel have depths (n-1, n, n+1) with n = total // 3."""
    if not mullevel:
        return [max_level]
    n = max_level // 3
    return [n - 1, n, n + 1]


def decodeOct(binfile, oct_data_seq, model, context_size=8192, anc_k=4, mullevel=MULLEVEL, lidar_level=None):
    """Returns (occupancy symbols 0..254 in BFS order, bin_num, z_offset, seconds, spher, cylin) like
    decode_ehem.py:56-188.  ``oct_data_seq`` (optional, (N,1) occupancy codes 1..255) is only used for the reference's
    own check (:184)."""
    spher, cylin, pos_mm, max_level, bin_num, z_offset = extract_info(binfile)
    depths = sub_depths(max_level, mullevel)
    if lidar_level is None:
        lidar_level = depths[1] if mullevel else max_level       # decode_ehem_mullevel.py:190 / decode_ehem.py:222
    with open(binfile, 'rb') as f:
        stream = f.read()
    mm = [(int(round(float(a))), int(round(float(b)))) for a, b in np.asarray(pos_mm, np.float64)]
    fr = types.SimpleNamespace(bitstream=stream, depths=depths, pos_mm=mm)
    dec = Decoder(model, lidar_level, 'spher' if spher else 'cylin' if cylin else 'cart', mullevel=mullevel)
    t0 = time.time()
    out = dec.decode(fr)
    torch.cuda.synchronize()
    elapsed = time.time() - t0
    code = np.concatenate(out.occ).astype(np.int64) - 1
    if oct_data_seq is not None:
        label = np.asarray(oct_data_seq).reshape(-1)
        assert len(label) == len(code) and (label == code + 1).all(), "decoded occupancy differs from the original"
    decodeOct.last = out                                             # voxels of the last call (DeOctree result)
    return code.tolist(), bin_num, z_offset, elapsed, spher, cylin


__all__ = ["extract_info", "decodeOct", "dequantise", "sub_depths"]
