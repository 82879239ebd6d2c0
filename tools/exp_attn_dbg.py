import sys, os
os.environ["SCP_ATTN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scp_b200.ops import CudaOps, V
cu = CudaOps()
T = 8192 * 16
g = torch.Generator().manual_seed(0)
qkv = torch.randn(T, 768, generator=g).cuda()
b = torch.randn(3, 256, generator=g).cuda()
rel = (torch.randn(1023, 4, generator=g) * 0.5).cuda()
y = torch.zeros(T, 256, device="cuda")
seqs = cu.seqs([i * 8192 for i in range(17)])
for _ in range(3):
    cu.swin_attention(V(qkv, 0, 256), V(qkv, 256, 256), V(qkv, 512, 256), b[0], b[1], b[2], rel, 4, seqs, 0, V(y))
torch.cuda.synchronize()
