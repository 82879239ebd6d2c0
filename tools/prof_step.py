"""Workload for ncu captures: one warm-up + `steps` timed encode_device() steps of one K16-mullevel frame (BASELINE config 2; no CPU baseline)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scp_b200.encoder import Encoder
from scp_b200.models import EHEM

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.cuda.set_device(0)
c = bench.CONFIGS[2]
model = EHEM(bench.model_cfg("EHEM")).cuda()
enc = Encoder(model, c["level"], c["mode"], mullevel=c["mullevel"], kind=c["kind"])
frames = bench.make_frames(c, 1, 0)
offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int64)
xyz = torch.from_numpy(np.concatenate(frames, 0)).cuda()
for _ in range(1 + steps):
    enc.encode_device(xyz, offs)
torch.cuda.synchronize()
print("done")
