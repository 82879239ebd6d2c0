#!/bin/bash
# occupancy variants of k_tree_occ / k_onesweep: stage timings on the 256-frame batch
mkdir -p gpurun_out
run() {
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
}
echo "== base"; run
echo "== occ 5"; SCP_OCC_VARIANT=1 run
echo "== occ 6"; SCP_OCC_VARIANT=2 run
echo "== sort 5"; SCP_SORT_VARIANT=1 run
