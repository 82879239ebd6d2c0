"""Stage timings of the octree pipeline on a batch of synthetic frames (development aid)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scp_b200 import octree, synth  # noqa: E402


def main(n_frames=64, level=16, mul=False, mode="spher"):
    base = [synth.kitti_sweep(s, 120000) for s in range(4)]
    frames = [base[i % 4] for i in range(n_frames)]
    offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])])
    xyz = torch.from_numpy(np.concatenate(frames, 0)).cuda()
    if mul:
        jobs = [j for i in range(n_frames) for j in octree.mullevel_jobs(i, level)]
    else:
        jobs = [octree.JobSpec(i, synth.KITTI_QS(level), None, lidar_level=level) for i in range(n_frames)]
    b = octree.OctreeBuilder()
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.time()
        b.plan(xyz, offs, jobs, mode)
        out = b.emit(("occ", "ctx", "pos_norm"), finish=False)
        torch.cuda.synchronize()
        wall = time.time() - t0
        ms = b.stage_ms()
    npts = int(offs[-1]) * (3 if mul else 1)
    N = b.total_rows
    P = (3 * max(i.depth for i in b.infos) + 1 + 7) // 8
    bytes_model = b.stage_bytes()
    rep = {k: {"ms": round(v, 4), "GBps": round(bytes_model[k] / v / 1e6, 1) if v > 0 else None,
                "frac": round(bytes_model[k] / v / 1e6 / 6547.2, 3) if v > 0 else None} for k, v in ms.items()}
    print(json.dumps({"frames": n_frames, "level": level, "mullevel": mul, "points": npts, "kept": b.total_kept, "rows": N, "P": P,
                      "depths": sorted(set(i.depth for i in b.infos)), "wall_ms": round(wall * 1e3, 3),
                      "device_ms": round(sum(ms.values()), 3), "stages": rep}))


if __name__ == "__main__":
    main(64, 16, False)
    main(64, 16, True)
    main(256, 16, True)
    main(256, 16, False)
    main(256, 14, False, "cylin")
