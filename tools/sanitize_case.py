"""Small end-to-end case for compute-sanitizer (tools/gpu/sanitize.sh): one 3000-point frame through the octree kernels
(k_quantise_*, k_onesweep, k_emit_nodes, k_occupancy, k_context*), a 1024-token window through SCP-EHEM (k_gemm_x3_ts,
k_swin_attn_h, k_knn_tc, k_knn_small, ...), a 600-token OctAttention window, softmax -> CDF and the coding order."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scp_b200 import synth                      # noqa: E402
from scp_b200.encoder import Encoder            # noqa: E402
from scp_b200.models import EHEM, OctAttention  # noqa: E402

NS = types.SimpleNamespace
cfg_e = NS(model=NS(context_size=8192, token_num=255, max_level=19), train=NS(type="kitti"), data=NS(extra_pos=False))
cfg_o = NS(model=NS(max_octree_level=12, context_size=1024, token_num=255, layer_num=3, head_num=4, abs_pos_embed_dim=12,
                    occ_embed_dim=128, level_embed_dim=6, octant_embed_dim=4, hidden_dimension=300, pos_embed=True),
           train=NS(type="kitti", dropout=0.0))
which = sys.argv[1] if len(sys.argv) > 1 else "ehem"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
pts = synth.make_frame("kitti", 3, 12, "spher", guard=True, n_points=n)[0]
if which == "ehem":
    r = Encoder(EHEM(cfg_e).cuda(), 12, "spher", mullevel=False).encode([pts])[0]
elif which == "mullevel":
    r = Encoder(EHEM(cfg_e).cuda(), 16, "spher", mullevel=True).encode([pts])[0]
else:
    r = Encoder(OctAttention(cfg_o).cuda(), 12, "spher", mullevel=False).encode([pts])[0]
torch.cuda.synchronize()
print(which, "nodes", r.n_nodes, "bytes", len(r.bitstream))
