"""kNN kernel timing: full / no accepts / no scan (SCP_KNN_DBG=0/1/2), 16 windows of 8192 tokens; splits out the kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scp_b200.ops import CudaOps, V
cu = CudaOps()
d = int(sys.argv[1]) if len(sys.argv) > 1 else 192
nw = 16
g = torch.Generator().manual_seed(0)
x = torch.cumsum(torch.randn(nw * 8192, d, generator=g) * 0.05, 0).cuda() + torch.randn(nw * 8192, d, generator=g).cuda() * 0.3
seqs = cu.seqs([i * 8192 for i in range(nw + 1)])
for dbg in (0, 1, 2):
    os.environ["SCP_KNN_DBG"] = str(dbg)
    for _ in range(2):
        cu.knn(V(x), seqs, 20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        cu.knn(V(x), seqs, 20)
    e1.record(); torch.cuda.synchronize()
    print(f"d={d} dbg={dbg}: {e0.elapsed_time(e1)/3:.2f} ms per call ({nw} windows of 8192)", flush=True)
