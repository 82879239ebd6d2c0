"""Development aid: wall/device time of the phases of one encode step, step by step."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from bench import cfg_ehem, make_frames
from scp_b200.encoder import Encoder
from scp_b200.models import EHEM
from scp_b200 import _lib

model = EHEM(cfg_ehem()).cuda()
enc = Encoder(model, 16, "spher", mullevel=True)
frames = make_frames(2, 0)
offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int64)
xyz = torch.from_numpy(np.concatenate(frames, 0)).cuda()
orig_fwd = model.forward_ragged
orig_ctx = enc.build_context
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.time()
        r = fn(*a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.time() - t0
        return r
    return w
model.forward_ragged = timed("forward", orig_fwd)
enc.build_context = timed("octree", orig_ctx)
for step in range(7):
    T.clear()
    torch.cuda.synchronize(); t0 = time.time()
    enc.encode_device(xyz, offs)
    torch.cuda.synchronize(); tot = time.time() - t0
    print(step, "total %.1f ms" % (tot * 1e3), {k: round(v * 1e3, 1) for k, v in T.items()},
          "mem GB alloc/reserved", round(torch.cuda.memory_allocated() / 2**30, 1), round(torch.cuda.memory_reserved() / 2**30, 1),
          "num_alloc_retries", torch.cuda.memory_stats().get("num_alloc_retries"), "cudaMalloc calls", torch.cuda.memory_stats().get("num_device_alloc"))
