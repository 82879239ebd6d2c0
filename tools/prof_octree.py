"""Workload for the octree ncu capture: 256 K16-mullevel frames, plan + emit twice."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from scp_b200 import octree, synth
base = [synth.kitti_sweep(s, 120000) for s in range(4)]
frames = [base[i % 4] for i in range(256)]
offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])])
xyz = torch.from_numpy(np.concatenate(frames, 0)).cuda()
jobs = [j for i in range(256) for j in octree.mullevel_jobs(i, 16)]
b = octree.OctreeBuilder()
for it in range(2):
    b.plan(xyz, offs, jobs, "spher")
    out = b.emit(("occ", "sym", "ctx", "pos_norm"), finish=False)
    torch.cuda.synchronize()
print(b.stage_ms())
