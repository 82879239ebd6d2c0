"""cProfile of encode_device: where does the host block / spend time inside a step?"""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scp_b200.encoder import Encoder
from scp_b200.models import EHEM

torch.cuda.set_device(0)
model = EHEM(bench.cfg_ehem()).cuda()
enc = Encoder(model, bench.LEVEL, "spher", mullevel=True, kind="kitti")
frames = bench.make_frames(2, 0)
offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int64)
xyz = torch.from_numpy(np.concatenate(frames, 0)).cuda()
for _ in range(3):
    enc.encode_device(xyz, offs)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
for _ in range(8):
    enc.encode_device(xyz, offs)
torch.cuda.synchronize()
pr.disable()
print(f"8 steps: {(time.time()-t0)/8*1e3:.1f} ms/step")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:6000])
