"""Which part of k_context_lean costs what: stage time of the context kernel per output subset (development aid)."""
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
b.plan(xyz, offs, jobs, "spher")
for outs in [("occ",), ("sym",), ("occ", "sym"), ("ctx",), ("pos_norm",), ("ctx", "pos_norm"), ("occ", "sym", "ctx", "pos_norm"),
             ("occ", "sym", "ctx", "pos_norm", "level")]:
    best = 1e9
    for it in range(3):
        out = b.emit(outs, finish=False)
        torch.cuda.synchronize()
        best = min(best, b.stage_ms()["context"])
        del out
    print(outs, round(best, 4), "ms")
