"""kNN time inside a real EHEM forward (1 K16-mullevel frame) for several compaction triggers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scp_b200.encoder import Encoder
from scp_b200.models import EHEM
torch.cuda.set_device(0)
model = EHEM(bench.cfg_ehem()).cuda()
enc = Encoder(model, bench.LEVEL, "spher", mullevel=True, kind="kitti")
frames = bench.make_frames(1, 0)
offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int64)
xyz = torch.from_numpy(np.concatenate(frames, 0)).cuda()
enc.encode_device(xyz, offs); torch.cuda.synchronize()
for trig in (224, 160, 128, 96, 64, 48):
    os.environ["SCP_KNN_TRIG"] = str(trig)
    enc.encode_device(xyz, offs)
    model.ops.reserve_events(4000); model.ops.prof = []
    enc.encode_device(xyz, offs); torch.cuda.synchronize()
    prof, model.ops.prof = model.ops.prof, None
    t = {}
    for tag, fl, by, a, b in prof:
        t[tag] = t.get(tag, 0.0) + a.elapsed_time(b)
    print(f"trig {trig}: knn_d144 {t.get('knn_d144',0):.2f} ms  knn_d192 {t.get('knn_d192',0):.2f} ms  (per frame)", flush=True)
