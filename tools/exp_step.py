"""Where does a bench step go?  Per-step GPU/CPU time over many steps + SM clocks, and a phase breakdown of one step."""
import sys, os, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scp_b200.encoder import Encoder
from scp_b200.models import EHEM

torch.cuda.set_device(0)
model = EHEM(bench.cfg_ehem()).cuda()
enc = Encoder(model, bench.LEVEL, "spher", mullevel=True, kind="kitti")
F = 2
frames = bench.make_frames(F, 0)
offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int64)
xyz = torch.from_numpy(np.concatenate(frames, 0)).cuda()
for _ in range(3):
    enc.encode_device(xyz, offs)
torch.cuda.synchronize()

clk = []
p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu",
                      "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
def rd():
    for line in p.stdout:
        clk.append((time.time(), line.strip()))
threading.Thread(target=rd, daemon=True).start()

N = 14
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
cpu = []
torch.cuda.synchronize()
ev[0].record()
t_start = time.time()
for i in range(N):
    c0 = time.time()
    enc.encode_device(xyz, offs)
    cpu.append(time.time() - c0)
    ev[i + 1].record()
torch.cuda.synchronize()
t_end = time.time()
print("gpu ms/step:", " ".join(f"{ev[i].elapsed_time(ev[i+1]):.0f}" for i in range(N)))
print("cpu ms/step:", " ".join(f"{c*1e3:.0f}" for c in cpu))
sel = [l for t, l in clk if t_start <= t <= t_end]
print("clocks (sm MHz, W, power_cap, temp) every ~0.5 s:", " | ".join(sel[::5]))
p.terminate()

# phase breakdown of one step (sync between phases)
def timed(fn):
    torch.cuda.synchronize(); t0 = time.time(); r = fn(); torch.cuda.synchronize(); return r, (time.time() - t0) * 1e3
for rep in range(2):
    (b, t, pf), t_oct = timed(lambda: enc.build_context(xyz, offs))
    infos = b.infos
    interval_row = torch.empty((b.total_rows, 2), dtype=torch.int32, device=xyz.device)
    _, t_model = timed(lambda: enc._ehem_logits_to_intervals(t, infos, interval_row))
    print(f"phase: octree+context {t_oct:.1f} ms, windows+model+cdf {t_model:.1f} ms", flush=True)
