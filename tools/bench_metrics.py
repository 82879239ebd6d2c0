"""Times the distortion report of one full-size frame (120 k points, level-16 mullevel cloud) with CUDA events:
scp_dequantise_keys + the two scp_nn_dist2 passes, and the FP64 instruction rate of the brute-force search."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scp_b200 import metrics, octree as oc, synth  # noqa: E402

pc = synth.kitti_sweep(3, 120000)[:, :3].astype(np.float32)
xyz = torch.from_numpy(pc).cuda()
b = oc.OctreeBuilder().plan(xyz, [0, len(pc)], oc.mullevel_jobs(0, 16, "kitti"), "spher")
vk = b.emit(("voxel_key",))["voxel_key"]
a = xyz.double()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
res = {}
for it in range(4):
    ev[0].record()
    cloud = metrics.dequantised_cloud(b, vk, "spher")
    ev[1].record()
    ab = metrics.nn_dist2(a, cloud)
    ev[2].record()
    ba = metrics.nn_dist2(cloud, a)
    ev[3].record()
    torch.cuda.synchronize()
    res = {"dequantise_ms": ev[0].elapsed_time(ev[1]), "nn_ab_ms": ev[1].elapsed_time(ev[2]), "nn_ba_ms": ev[2].elapsed_time(ev[3])}
pairs = float(len(a)) * float(len(cloud))
res.update(points=len(a), cloud=len(cloud), pairs=pairs,
           fp64_tinstr_per_s=9 * pairs / (res["nn_ab_ms"] * 1e-3) / 1e12,
           chamfer_psnr=metrics.distortion(a, cloud, metrics.KITTI_PEAK))
print(json.dumps(res))
