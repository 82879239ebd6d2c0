"""Workload for the OctAttention attention ncu capture: 64 full 1024-token windows through scp_octattn_attention."""
import sys
sys.path.insert(0, ".")
import torch
from scp_b200.ops import CudaOps, V
ops = CudaOps()
nwin = 64
T = nwin * 1024
g = torch.Generator().manual_seed(0)
a = (torch.randn(T, 3000, generator=g) * 0.3).cuda()
o, ou = torch.empty(T, 600, device="cuda"), torch.empty(T, 600, device="cuda")
seqs = ops.seqs([i * 1024 for i in range(nwin + 1)])
for _ in range(2):
    ops.octattn_attention(V(a, 1800, 600), V(a, 0, 600), V(a, 1200, 600), V(a, 600, 600), V(a, 2400, 600), 4, 150, seqs, V(o), V(ou))
torch.cuda.synchronize()
print("ok", float(ou.abs().mean()))
