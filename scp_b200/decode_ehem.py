"""Drop-in for the reference's ``decode_ehem.py``: same ``extract_info(file)`` (:20-27) and
``decodeOct(binfile, oct_data_seq, model, context_size, anc_k)`` (:56-188) contracts on the files ``compress_ehem``
writes (``<name>[_spher|_cylin]_<levels>_<bin_num>_<z_offset>.bin`` + ``.bin.dat``).  The per-node Python
(`cur_nodes` / `pre_nodes` deques, torch child expansion, numpyAc calls per window) is replaced by
scp_b200.decoder.Decoder: level-wise batches on the GPU, host range decoder in the library."""
import argparse
import math
import os
import time
import types
from pathlib import Path

import numpy as np
import torch

from .data_preproc import pt
from .decoder import Decoder, dequantise
from .synth import FORD_QS, KITTI_QS

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


def reconstruct(frame, bin_num, z_offset, spher, cylin, lidar_level, kind="kitti"):
    """Tail of ``main`` (:228-252): DeOctree voxels -> ``v * qs + offset`` -> spher2cart / cylin2cart, float64 [V,3].
    ``frame`` = ``decodeOct.last``.  For a mullevel stream every sub-octree is dequantised at its own step (the reference's
    mullevel ``main`` applies one step to the concatenated code and is marked untested, decode_ehem_mullevel.py:251-252).
    The file name only carries the first sub-octree's ``bin_num``; the finer ones are re-derived from it
    (``round(rho_max / qs_i) + 1`` with ``rho_max ~ (bin_num - 1) qs_0``), which can be off by one bin in ~10^5 -- a format
    limitation: angles of the two small far-range sub-octrees are then scaled by 1 +- 1e-5."""
    qf = FORD_QS if kind == "ford" else KITTI_QS
    parts = []
    for i, vox in enumerate(frame.voxels):
        qs = qf(lidar_level + i)
        bn = bin_num if i == 0 else int(round((bin_num - 1) * qf(lidar_level) / qs)) + 1
        if spher:
            parts.append(dequantise(vox, [qs, 2 * math.pi / (bn - 1), math.pi / (bn - 1)], [0, 0, 0], "spher"))
        elif cylin:
            parts.append(dequantise(vox, [qs, 2 * math.pi / (bn - 1), qs], [0, 0, z_offset], "cylin"))
        else:
            parts.append(dequantise(vox, [qs, qs, qs], [-200, -200, -200], "cart"))
    return np.vstack(parts)


def main(args, mullevel=MULLEVEL):
    """decode_ehem.py:190-254 without hydra: for every original sweep, find its ``.bin`` in ``--out_dir``, decode it,
    check the symbols against the pre-generated rows when ``--preproc_path`` has them (:217-221), write ``<stem>.ply``."""
    from .encode import build_model
    model = build_model(types.SimpleNamespace(model="EHEM", type=args.type, ckpt_path=args.ckpt_path))
    out_dir = args.out_dir.rstrip('/') + '/'
    if os.path.isdir(args.test_files[0]):
        d = args.test_files[0]
        args.test_files = sorted(d + x for x in os.listdir(d) if x.endswith(('.ply', '.bin')))
    elapsed, written = 0, []
    for i, ori_file in enumerate(args.test_files):
        print(f'{i}/{len(args.test_files)}')
        ori = Path(ori_file)
        name = (ori_file.split('/')[-2] + ori.stem) if args.type == 'kitti' and ori_file.count('/') >= 2 else ori.stem
        binfile = next(out_dir + f for f in sorted(os.listdir(out_dir)) if f.startswith(name + '_') and f.endswith('.bin'))
        labels = None
        if args.preproc_path:
            base = args.preproc_path.rstrip('/') + '/' + name
            tags = ('_0_0', '_0_1', '_1') if mullevel else ('',)
            labels = np.concatenate([np.load(base + t + '.npy')[:, -1, 0] for t in tags])
        code, bin_num, z_offset, t, spher, cylin = decodeOct(binfile, labels, model, 8192, 4, mullevel=mullevel)
        elapsed += t
        print("decode succeeded, time:", t)
        print("oct len:", len(code))
        print("avg dec time:", elapsed / (i + 1))
        lidar_level = args.lidar_level
        if lidar_level is None:                               # the reference's rule (:222): the level count of the name
            lidar_level = sub_depths(int(binfile.split('/')[-1].split('_')[-3]), mullevel)[1 if mullevel else 0]
        pt_rec = reconstruct(decodeOct.last, bin_num, z_offset, spher, cylin, lidar_level, args.type)
        pt.write_ply_data(out_dir + ori.stem + ".ply", pt_rec)
        print(out_dir + ori.stem + ".ply")
        written.append(out_dir + ori.stem + ".ply")
    print(elapsed / len(args.test_files))
    return written


def get_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--ckpt_path", type=str, default="", help="optional state_dict / Lightning checkpoint")
    parser.add_argument("--test_files", nargs="*", required=True, help="the ORIGINAL sweeps (names select the .bin files)")
    parser.add_argument("--out_dir", type=str, default="test_output", help="where encode.py wrote the .bin / .dat files")
    parser.add_argument("--type", type=str, default='kitti', choices=['kitti', 'ford'])
    parser.add_argument("--lidar_level", type=int, default=None,
                        help="quantisation level the stream was encoded with (default: the level count in the file name, "
                             "like the reference -- only right when the octree is as deep as the level)")
    parser.add_argument("--preproc_path", type=str, default="")
    parser.add_argument("--sequential_enc", action="store_true")
    parser.add_argument("--level_wise", action="store_true")
    return parser.parse_args(argv)


__all__ = ["extract_info", "decodeOct", "dequantise", "sub_depths", "reconstruct", "main", "get_args"]


if __name__ == "__main__":
    main(get_args())
