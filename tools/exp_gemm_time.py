"""nn.Linear kernel time on the four layer shapes of a stage-0 Swin layer of a K16-mullevel frame (511k tokens); development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scp_b200.ops import CudaOps, V
cu = CudaOps()
M = 511 * 1024
g = torch.Generator().manual_seed(0)
tot = 0.0
for name, N, K, act, res in (("qkv", 768, 256, "none", False), ("mlp1+gelu", 1024, 256, "gelu", False),
                             ("mlp2+res", 256, 1024, "none", True), ("proj+res", 256, 256, "none", True)):
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).cuda() if res else None
    y = torch.empty(M, N, device="cuda")
    for _ in range(3):
        cu.linear(V(x), w, b, V(y), act=act, res=V(r) if res else None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        cu.linear(V(x), w, b, V(y), act=act, res=V(r) if res else None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tot += ms
    print(f"{name:10s} M={M} N={N} K={K}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.0f} TFLOP/s algorithmic", flush=True)
print(f"sum {tot:.3f} ms  (SCP_GEMM_DBG={os.environ.get('SCP_GEMM_DBG','0')})")
