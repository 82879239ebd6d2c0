"""one cluster-GEMM case per process (a device trap kills the context): python tools/exp_gemm_cl.py CL M N K"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scp_b200.ops import CudaOps, V
cl, M, N, K = (int(a) for a in sys.argv[1:5])
cu = CudaOps(engine="f16x3")
cu.lib.scp_set_gemm_cluster(cl)
g = torch.Generator().manual_seed(1)
x = torch.randn(M, K, generator=g).cuda()
w = (torch.randn(N, K, generator=g) * 0.1).cuda()
b = torch.randn(N, generator=g).cuda()
y = torch.empty(M, N, device="cuda")
cu.linear(V(x), w, b, V(y))
torch.cuda.synchronize()
ref = x.double() @ w.double().T + b.double()
print(f"CL={cl} M={M} N={N} K={K}: ok, max err {(y.double() - ref).abs().max().item():.2e}", flush=True)
