#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_octree_gpu.py -q -m gpu -x 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "octree tests failed: stop"; exit 1; }
timeout 300 python - <<'PY'
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
best = None
for it in range(5):
    b.plan(xyz, offs, jobs, "spher")
    out = b.emit(("occ", "sym", "ctx", "pos_norm"), finish=False)
    torch.cuda.synchronize()
    m = b.stage_ms()
    best = m if best is None else {k: min(best[k], v) for k, v in m.items()}
print({k: round(v, 3) for k, v in best.items()})
PY
