"""End-to-end encode path on the GPU against the reference's own end-to-end run (tests/golden/e2e_*.npz):
symbols in coding order bit-exact, bpp within 0.5 % (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from conftest import golden
from test_models_cpu import cfg_ehem, cfg_oct

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,mul,level", [("k12s", False, 12), ("k16m", True, 16)])
def test_encode_matches_reference_run(name, mul, level):
    from scp_b200.encoder import Encoder
    from scp_b200.models import EHEM
    g = golden(f"octree_{name}.npz")
    e = golden(f"e2e_{name}.npz")
    model = EHEM(cfg_ehem()).cuda()
    enc = Encoder(model, level, "spher", mullevel=mul)
    pts = g["points"]
    # symbols in coding order (bit-exact)
    xyz = torch.from_numpy(pts).cuda()
    interval, frames, infos, per_frame = enc.encode_device(xyz, [0, len(pts)])
    from scp_b200 import coder
    sizes = [n for i in infos for n in i.level_rows]
    occ = enc.builder.emit(("occ",), finish=False)["occ"]
    order, sym = coder.coding_order(sizes, 8192, occ, mullevel=mul)
    assert np.array_equal(sym.cpu().numpy(), e["sym"])
    iv = interval.cpu().numpy().view(np.uint32)
    assert (iv[:, 1] > iv[:, 0]).all() and (iv[:, 1] <= 0x10000).all()
    # bitstream size / bpp
    res = enc.encode([pts])[0]
    ref_bytes = len(e["bitstream"])
    dev = abs(len(res.bitstream) - ref_bytes) / ref_bytes
    print(name, "bytes ours", len(res.bitstream), "reference", ref_bytes, "bpp", res.bpp, "ref bpp", float(e["bpp"]), "dev", dev)
    assert dev < 0.005
    assert res.n_nodes == len(e["sym"]) and res.bin_num == int(g["bin_num"])
    assert res.n_levels == len(g["level_sizes"])


def test_batch_of_frames_equals_frame_by_frame():
    """Frames are independent (multi-GPU partition = subsets of frames): batching must not change any stream."""
    from scp_b200 import synth
    from scp_b200.encoder import Encoder
    from scp_b200.models import EHEM
    model = EHEM(cfg_ehem()).cuda()
    enc = Encoder(model, 12, "spher", mullevel=False)
    frames = [synth.make_frame("kitti", s, 12, "spher", guard=True, n_points=n)[0] for s, n in ((1, 4000), (2, 2500), (3, 6000))]
    together = enc.encode(frames)
    for f, r in zip(frames, together):
        single = enc.encode([f])[0]
        assert single.bitstream == r.bitstream and single.n_nodes == r.n_nodes


def test_octattention_encode_runs_and_codes_all_nodes():
    from scp_b200 import synth
    from scp_b200.encoder import Encoder
    from scp_b200.models import OctAttention
    model = OctAttention(cfg_oct()).cuda()
    enc = Encoder(model, 12, "spher", mullevel=False)
    pts = synth.make_frame("kitti", 5, 12, "spher", guard=True, n_points=3000)[0]
    r = enc.encode([pts])[0]
    assert r.n_nodes > 1000 and len(r.bitstream) > 100


def test_encode_stream_equals_encode():
    """The pipelined generator returns, batch by batch, exactly what encode() returns."""
    from scp_b200.encoder import Encoder
    from scp_b200.models import EHEM
    from scp_b200 import synth
    model = EHEM(cfg_ehem()).cuda()
    enc = Encoder(model, 12, "spher", mullevel=False)
    batches = [[synth.make_frame("kitti", seed=s, level=12, mode="spher", guard=True, n_points=2500)[0] for s in (b, b + 10)]
               for b in range(4)]
    ref = [enc.encode(b) for b in batches]
    got = list(enc.encode_stream(iter(batches)))
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert [x.bitstream for x in g] == [x.bitstream for x in r]
        assert [x.n_nodes for x in g] == [x.n_nodes for x in r]
        assert [x.pos_mm for x in g] == [x.pos_mm for x in r]
