"""MMA-warp timeline of the nn.Linear kernel on the four layer shapes of a Swin layer (development aid)."""
import os, sys
os.environ["SCP_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scp_b200.ops import CudaOps, V
cu = CudaOps()
M = 511 * 1024
g = torch.Generator().manual_seed(0)
for N, K, act, res in ((768, 256, "none", False), (1024, 256, "gelu", False), (256, 1024, "none", True), (256, 256, "none", True)):
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).cuda() if res else None
    y = torch.empty(M, N, device="cuda")
    for _ in range(2):
        cu.linear(V(x), w, b, V(y), act=act, res=V(r) if res else None)
    torch.cuda.synchronize()
