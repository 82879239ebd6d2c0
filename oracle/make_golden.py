"""Generates tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container only (``python oracle/make_golden.py``); the fixtures are committed
so nothing on the GPU box needs /root/reference.  What is stored is always an *output of the
reference code itself* (C++ ``Octree_python_lib.so`` + ``gen_K_parent_seq`` via ``proc_pc``,
``mul_proc_pc``, the dataset ``__getitem__``s, ``EHEM.forward``, ``OctAttention.forward``,
``compress_ehem`` and ``numpyAc``) on small seeded synthetic inputs.
"""
import os
import sys
import tempfile
import types
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(HERE, "_ref", "torch_ext"))

import torch  # noqa: E402

from oracle import ref_shims  # noqa: E402
from scp_b200 import synth, weights  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

OCTREE_CASES = {
    # name: (kind, seed, level, mode, n_points, mullevel)
    "k12s": ("kitti", 1, 12, "spher", 1500, False),
    "k14c": ("kitti", 2, 14, "cylin", 1500, False),
    "k16m": ("kitti", 3, 16, "spher", 2500, True),
    "f17s": ("ford", 4, 17, "spher", 1200, False),
    "k10c": ("kitti", 5, 10, "cylin", 3000, False),     # many duplicate voxels
    "k14s": ("kitti", 6, 14, "spher", 1500, False),     # OctAttention config 4 (levels above 12: oct_attention.py:57-61)
}


def case_points(kind, seed, level, mode, n_points, mullevel):
    if kind == "kitti":
        pts = synth.kitti_sweep(seed, 120000)
        qs = [synth.KITTI_QS(level + i) for i in range(3 if mullevel else 1)]
    else:
        pts = synth.ford_sweep(seed, 80000)
        qs = [synth.FORD_QS(level + i) for i in range(3 if mullevel else 1)]
    rng = np.random.default_rng(seed)
    pts = pts[np.sort(rng.choice(len(pts), n_points, replace=False))]
    for _ in range(4):
        for q in qs:
            pts = synth.guard_band(pts, q, mode, margin=0.03)
    return np.ascontiguousarray(pts), qs


def pack_levels(lst):
    return np.concatenate([np.asarray(a).reshape(len(a), -1) if np.asarray(a).ndim > 1 else np.asarray(a)[:, None]
                           for a in lst], 0)


def gen_octree(ns, tmp, only=None):
    dp = ns.data_preprocess
    for name, (kind, seed, level, mode, n_points, mullevel) in OCTREE_CASES.items():
        if only and name not in only:
            continue
        pts, qs = case_points(kind, seed, level, mode, n_points, mullevel)
        binf = os.path.join(tmp, name + ".bin")
        pts.astype(np.float32).tofile(binf)
        out = dict(points=pts, qs=np.array(qs), level=level, mode=mode, mullevel=mullevel)
        kw = dict(spher=(mode == "spher"), cylin=(mode == "cylin"))
        if not mullevel:
            res = dp.proc_pc(binf, tmp, name, qs=qs[0], test=True, normalize=False, **kw)
            rows = np.load(res[0] + ".npy")
            out["rows"] = rows.astype(np.int32)
            out["bin_num"] = float(res[3])
            out["dequant"] = np.asarray(res[1], np.float32)
            if mode == "cylin":
                out["z_offset"] = float(res[4][0, 2])
            ds = ns.ds_ehem.EncodeEHEMDataset([binf], 8192, kind, True, level, mode == "cylin", mode == "spher",
                                              False, False, "")
            if mode == "cylin":
                ds.preproc = lambda f, r=res: (r[0], r[2], 0.0, r[3], r[4][0, 2], 0.0)
            else:
                ds.preproc = lambda f, r=res: (r[0], r[2], 0.0, r[3], 0.0)
            ids, poss, pos_mm, data, oct_seq = ds[0][:5]
            # OctAttention dataset on the same .npy (level_wise False as encode.py runs it)
            if mode == "spher":
                dso = ns.ds_oct.EncodeDataset([binf], 1024, kind, False, level, True, "")
                dso.preproc = lambda f, r=res: (r[0], r[2], 0.0, r[3], 0.0)
                oids, opos, odata, _ = dso[0][:4]
                out["oct_ids"] = oids[0]
                out["oct_pos"] = opos[0]
                out["oct_data"] = odata[0].astype(np.int16)
        else:
            paths = [[0, 0], [0, 1], [1]]
            files, rows_l, bn = [], [], None
            for q, mp in zip(qs, paths):
                res = dp.mul_proc_pc(binf, tmp, name, qs=q, test=True, normalize=False, morton_path=mp, **kw)
                files.append(res[0])
                rows_l.append(np.load(res[0] + ".npy"))
                bn = res[3] if bn is None else bn
            out["bin_num"] = float(bn)
            out["rows"] = np.vstack(rows_l).astype(np.int32)
            out["sub_rows"] = np.array([len(r) for r in rows_l])
            ds = ns.ds_ehem_mul.EncodeEHEMDataset([binf], 8192, kind, True, level, mode == "cylin", mode == "spher", "")
            ds.preproc = lambda f: (files, pts[:, :3], 0.0, bn, 0.0, 0.0)
            ids, poss, pos_mm, data, oct_seq = ds[0][:5]
        out["level_sizes"] = np.array([len(i) for i in ids])
        out["ds_pos"] = np.concatenate([p.T for p in poss], 0).astype(np.float32)      # (N,3)
        out["ds_data"] = np.concatenate(data, 0).astype(np.int16)                       # (N,4,3)
        out["ds_pos_mm"] = np.array(pos_mm, np.int64)
        out["ds_oct_seq"] = oct_seq.astype(np.int32)
        np.savez_compressed(os.path.join(GOLD, f"octree_{name}.npz"), **out)
        print(name, "points", len(pts), "rows", out["rows"].shape, "levels", len(ids))


def load_ehem(ns, sharpen=True, seed=0):
    cfg = ref_shims.make_cfg("ehem")
    torch.manual_seed(0)
    m = ns.ehem.EHEM(cfg)
    spec = weights.ehem_spec(cfg.model.max_level)
    sd_ref = m.state_dict()
    assert [k for k, _, _ in spec] == list(sd_ref.keys()), "state_dict names/order differ from the reference"
    for k, shape, _ in spec:
        assert tuple(sd_ref[k].shape) == tuple(shape), (k, sd_ref[k].shape, shape)
    sd = weights.synth_state_dict(spec, seed=seed, sharpen=sharpen)
    for k in sd_ref:   # buffers we regenerate must equal the reference's own
        if "relative_position_index" in k:
            assert torch.equal(sd[k], sd_ref[k]), k
    m.load_state_dict(sd, strict=True)
    return m.eval()


def load_octattn(ns, sharpen=True, seed=0):
    cfg = ref_shims.make_cfg("oct", train_type="kitti")
    m = ns.oct_attention.OctAttention(cfg)
    spec = weights.octattn_spec()
    sd_ref = m.state_dict()
    assert [k for k, _, _ in spec] == list(sd_ref.keys())
    for k, shape, _ in spec:
        assert tuple(sd_ref[k].shape) == tuple(shape), (k, sd_ref[k].shape, shape)
    sd = weights.synth_state_dict(spec, seed=seed, sharpen=sharpen)
    assert torch.equal(sd["mask"], sd_ref["mask"])
    assert torch.allclose(sd["transformer_encoder.position_enc.pe"], sd_ref["transformer_encoder.position_enc.pe"])
    m.load_state_dict(sd, strict=True)
    return m.eval()


def gen_ehem_logits(ns):
    g = np.load(os.path.join(GOLD, "octree_k12s.npz"))
    sizes = np.cumsum(np.concatenate([[0], g["level_sizes"]]))
    data_all, pos_all = g["ds_data"].astype(np.int64), g["ds_pos"]
    big = int(np.argmax(g["level_sizes"]))
    s0 = sizes[big]
    model = load_ehem(ns, sharpen=True)
    out = {}
    with torch.no_grad():
        for tag, (a, b) in {"n1": (0, 1), "n2": (sizes[1], sizes[1] + 2), "n37": (s0, s0 + 37),
                            "n600": (s0, s0 + 600), "n1100": (s0 + 100, s0 + 1200)}.items():
            d = torch.from_numpy(data_all[a:b])[None]
            p = torch.from_numpy(pos_all[a:b].T.copy())[None]
            o1, o2 = model(d, p, enc=True)
            out[f"{tag}_data"] = data_all[a:b].astype(np.int16)
            out[f"{tag}_pos"] = pos_all[a:b].T.copy()
            out[f"{tag}_logits1"] = o1[0].numpy()
            out[f"{tag}_logits2"] = o2[0].numpy()
            print("ehem", tag, o1.shape, o2.shape, float(torch.softmax(o1, 2).max()))
        # tie-free variants: same context bytes, positions replaced by seeded uniform noise so that every
        # neighbour set is unambiguous (grid positions make torch.topk's tie order part of the output)
        for tag, (a, b) in {"j600": (s0, s0 + 600), "j1100": (s0 + 100, s0 + 1200)}.items():
            d = torch.from_numpy(data_all[a:b])[None]
            pj = np.random.RandomState(b).random_sample((3, b - a)).astype(np.float32)
            o1, o2 = model(d, torch.from_numpy(pj)[None], enc=True)
            out[f"{tag}_data"] = data_all[a:b].astype(np.int16)
            out[f"{tag}_pos"] = pj
            out[f"{tag}_logits1"] = o1[0].numpy()
            out[f"{tag}_logits2"] = o2[0].numpy()
            print("ehem", tag, o1.shape, o2.shape, float(torch.softmax(o1, 2).max()))
    np.savez_compressed(os.path.join(GOLD, "ehem_logits.npz"), **out)

    # one full 8192-token window on a denser frame; logits stored at every 16th token
    pts, qs0 = synth.make_frame("kitti", seed=7, level=12, mode="spher", guard=True, n_points=30000)
    from oracle import octree_np as onp
    q = onp.quantize(pts[:, :3], qs0, "spher")["q"]
    rows = onp.tree_rows(q)["rows"]
    ids, poss, pos_mm, data, _ = onp.ehem_level_split(rows, 12)
    li = int(np.argmax([len(i) for i in ids]))
    assert len(ids[li]) >= 8192
    d = torch.from_numpy(data[li][:8192])[None]
    p = torch.from_numpy(poss[li][:, :8192].copy())[None]
    with torch.no_grad():
        o1, o2 = model(d, p, enc=True)
    pj = np.random.RandomState(8192).random_sample((3, 8192)).astype(np.float32)
    with torch.no_grad():
        j1, j2 = model(d, torch.from_numpy(pj)[None], enc=True)
    np.savez_compressed(os.path.join(GOLD, "ehem_logits_full_jit.npz"), pos=pj,
                        logits1_s16=j1[0, ::16].numpy(), logits2_s16=j2[0, ::16].numpy())
    np.savez_compressed(os.path.join(GOLD, "ehem_logits_full.npz"),
                        data=data[li][:8192].astype(np.int16), pos=poss[li][:, :8192].copy(),
                        logits1_s16=o1[0, ::16].numpy(), logits2_s16=o2[0, ::16].numpy(),
                        pmf1_max=torch.softmax(o1[0], 1).max(1)[0].numpy(),
                        pmf2_max=torch.softmax(o2[0], 1).max(1)[0].numpy())
    print("ehem full", o1.shape)


def gen_octattn_logits(ns):
    g = np.load(os.path.join(GOLD, "octree_k12s.npz"))
    data, pos = g["oct_data"].astype(np.int64), g["oct_pos"]
    model = load_octattn(ns, sharpen=True)
    out = {}
    with torch.no_grad():
        for tag, (a, b) in {"w0": (0, 1024), "w3": (3 * 1024, 4 * 1024), "tail": (len(data) - 300, len(data))}.items():
            d = torch.from_numpy(data[a:b].copy())[None]
            p = torch.from_numpy(pos[a:b].copy())[None]
            o = model(d.clone(), p)
            out[f"{tag}_data"] = data[a:b].astype(np.int16)
            out[f"{tag}_pos"] = pos[a:b]
            out[f"{tag}_logits_s2"] = o[0, ::2].numpy()
            print("octattn", tag, o.shape, float(torch.softmax(o, 2).max()))
    np.savez_compressed(os.path.join(GOLD, "octattn_logits.npz"), **out)


def import_encode(ns):
    """Imports the reference's encode.py / encode_mullevel.py (needs the numpyAc JIT build)."""
    import importlib
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())     # the reference mkdirs ./temp on import/use
    try:
        enc = importlib.import_module("encode")
        encm = importlib.import_module("encode_mullevel")
        nac = importlib.import_module("numpyAc.numpyAc")
    finally:
        os.chdir(cwd)
    return enc, encm, nac


def coder_case(n=3000, seed=11):
    """PMFs (n,255) float32 + symbols; shared with tests/test_cdf_coder.py."""
    r = np.random.RandomState(seed)
    w = r.randint(1, 64, (n, 255)).astype(np.float64) ** 3
    w[np.arange(n), r.randint(0, 255, n)] *= r.randint(1, 4000, n)
    pmf = (w / w.sum(1, keepdims=True)).astype(np.float32)
    cum = np.cumsum(pmf.astype(np.float64), 1)
    u = r.random_sample(n) * cum[:, -1]
    sym = np.minimum((cum < u[:, None]).sum(1), 254).astype(np.int16)
    return pmf, sym


def gen_coder(ns, tmp):
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    enc, encm, nac = import_encode(ns)
    # (1) numpyAc on seeded PMFs built from integer weights with IEEE-exact ops only (machine independent)
    pmf, sym = coder_case()
    cdfF = nac.pdf_convert_to_cdf_and_normalize(pmf)
    cdf = nac._convert_to_int_and_normalize(cdfF, True)
    bs, bits = nac.arithmeticCoding().encode(pmf, sym)
    np.savez_compressed(os.path.join(GOLD, "coder.npz"), sym=sym, cdf_crc=zlib.crc32(cdf.tobytes()),
                        cdf_rows=cdf.view(np.uint16)[::50], bitstream=np.frombuffer(bs, np.uint8), bits=bits)
    print("coder", bits, "bits", bits / len(sym), "bit/sym")

    # (2) end-to-end compress_ehem of the reference on the two small frames
    model = load_ehem(ns, sharpen=True)
    model.cfg = ref_shims.make_cfg("ehem")
    for name, mod in (("k12s", enc), ("k16m", encm)):
        g = np.load(os.path.join(GOLD, f"octree_{name}.npz"))
        sizes = np.cumsum(np.concatenate([[0], g["level_sizes"]]))
        ids = [torch.arange(n)[None] for n in g["level_sizes"]]
        pos = [torch.from_numpy(g["ds_pos"][a:b].T.copy())[None] for a, b in zip(sizes[:-1], sizes[1:])]
        data = [torch.from_numpy(g["ds_data"][a:b].astype(np.int64))[None] for a, b in zip(sizes[:-1], sizes[1:])]
        oct_seq = torch.from_numpy(g["ds_oct_seq"].astype(np.int64))[None]
        captured = {}
        real_encode = nac.arithmeticCoding.encode

        def spy(self, pdf, sym, binfile=None):
            captured["pdf_s16"] = np.asarray(pdf)[::16].copy()
            captured["sym"] = np.asarray(sym).copy()
            bsx, bitsx = real_encode(self, pdf, sym, binfile)
            captured["bitstream"] = np.frombuffer(bsx, np.uint8).copy()
            return bsx, bitsx

        nac.arithmeticCoding.encode = spy
        args = types.SimpleNamespace(spher=True, cylin=False)
        batch = (ids, pos, [(0, 1)] * len(ids), data, oct_seq, torch.tensor(len(g["points"])), None,
                 torch.tensor(int(g["bin_num"])), torch.tensor(0))
        bpp, _ = mod.compress_ehem(batch, os.path.join(tmp, "out", name), model, args)
        nac.arithmeticCoding.encode = real_encode
        np.savez_compressed(os.path.join(GOLD, f"e2e_{name}.npz"), bpp=bpp, n_points=len(g["points"]), **captured)
        print("e2e", name, "bpp", bpp, "bytes", len(captured["bitstream"]))
def _spy_encode(nac, captured):
    """Wraps numpyAc.arithmeticCoding.encode so that what the reference hands to its coder is recorded."""
    real_encode = nac.arithmeticCoding.encode

    def spy(self, pdf, sym, binfile=None):
        captured["pdf_s16"] = np.asarray(pdf)[::16].copy()
        captured["sym"] = np.asarray(sym).copy()
        bsx, bitsx = real_encode(self, pdf, sym, binfile)
        captured["bitstream"] = np.frombuffer(bsx, np.uint8).copy()
        return bsx, bitsx

    nac.arithmeticCoding.encode = spy
    return real_encode


def gen_octattn_e2e(ns, tmp):
    """The reference's own OctAttention end-to-end runs (SURVEY config 4): ``encode.compress`` (encode.py:23-82, one
    sequence [1023 pads ; nodes], windows of 1024) on the k12s and k14s frames, and ``encode_mullevel.compress``
    (encode_mullevel.py:23-85) on the three sub-octrees of k16m, whole-sequence and level-wise, fed by the reference's
    ``dataloaders/encode_dataset_mullevel.EncodeDataset`` (whose ``__getitem__`` output is stored for the mirror test)."""
    from torch.utils.data import default_collate
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    enc, encm, nac = import_encode(ns)
    dp = ns.data_preprocess
    model = load_octattn(ns, sharpen=True)
    model.cfg = ref_shims.make_cfg("oct", train_type="kitti")
    args = types.SimpleNamespace(spher=True, cylin=False, sequential=False)
    for name in ("k12s", "k14s"):
        kind, seed, level, mode, n_points, mullevel = OCTREE_CASES[name]
        pts, qs = case_points(kind, seed, level, mode, n_points, mullevel)
        binf = os.path.join(tmp, name + ".bin")
        pts.astype(np.float32).tofile(binf)
        res = dp.proc_pc(binf, tmp, name, qs=qs[0], test=True, normalize=False, spher=True)
        ds = ns.ds_oct.EncodeDataset([binf], 1024, kind, False, level, True, "")
        ds.preproc = lambda f, r=res: (r[0], r[2], 0.0, r[3], 0.0)
        batch = default_collate([ds[0]])[:-2]
        captured = {}
        real = _spy_encode(nac, captured)
        bpp, _ = enc.compress(batch, os.path.join(tmp, "out", name + "_oa"), model, args)
        nac.arithmeticCoding.encode = real
        np.savez_compressed(os.path.join(GOLD, f"e2e_octattn_{name}.npz"), bpp=bpp, n_points=len(pts), **captured)
        print("e2e octattn", name, "bpp", bpp, "bytes", len(captured["bitstream"]), "nodes", len(captured["sym"]))
    # mullevel: encode_mullevel.compress on the k16m sub-octrees
    name = "k16m"
    kind, seed, level, mode, n_points, mullevel = OCTREE_CASES[name]
    pts, qs = case_points(kind, seed, level, mode, n_points, mullevel)
    binf = os.path.join(tmp, name + ".bin")
    pts.astype(np.float32).tofile(binf)
    files, bn = [], None
    for q, mp in zip(qs, [[0, 0], [0, 1], [1]]):
        res = dp.mul_proc_pc(binf, tmp, name, qs=q, test=True, normalize=False, morton_path=mp, spher=True)
        files.append(res[0])
        bn = res[3] if bn is None else bn
    for lw in (False, True):
        ds = ns.ds_oct_mul.EncodeDataset([binf], 1024, kind, lw, level, True, "unused/")
        ds.preproc = lambda f: (files, pts[:, :3], 0.0, int(bn), 0.0)
        item = ds[0]
        batch = default_collate([item])[:-2]
        captured = {}
        real = _spy_encode(nac, captured)
        bpp, _ = encm.compress(batch, os.path.join(tmp, "out", name + "_oa"), model, args)
        nac.arithmeticCoding.encode = real
        tag = "lw" if lw else "seq"
        ids, pos, data, oct_seq = item[:4]
        np.savez_compressed(os.path.join(GOLD, f"e2e_octattn_{name}_{tag}.npz"), bpp=bpp, n_points=len(pts),
                            ds_sizes=np.array([len(i) for i in ids]), ds_ids=np.concatenate(ids).astype(np.int32),
                            ds_pos=np.concatenate(pos, 0), ds_data=np.concatenate(data, 0).astype(np.int16),
                            ds_oct_seq=oct_seq.astype(np.int32), **captured)
        print("e2e octattn", name, tag, "bpp", bpp, "bytes", len(captured["bitstream"]), "blocks", len(ids))


def gen_metrics(ns, tmp):
    """pt.distChamfer of the reference on (original cloud, quantised cloud of proc_pc / mul_proc_pc) of the octree cases."""
    dp = ns.data_preprocess
    out = {}
    for name, (kind, seed, level, mode, n_points, mullevel) in OCTREE_CASES.items():
        pts, qs = case_points(kind, seed, level, mode, n_points, mullevel)
        binf = os.path.join(tmp, name + ".bin")
        pts.astype(np.float32).tofile(binf)
        kw = dict(spher=(mode == "spher"), cylin=(mode == "cylin"))
        if mullevel:
            q = np.vstack([dp.mul_proc_pc(binf, tmp, name, qs=s, test=True, normalize=False, morton_path=mp, **kw)[1]
                           for s, mp in zip(qs, [[0, 0], [0, 1], [1]])])
        else:
            q = dp.proc_pc(binf, tmp, name, qs=qs[0], test=True, normalize=False, **kw)[1]
        pc = dp.pointCloud.ptread(binf)
        out[name + "_q"] = np.asarray(q)
        out[name + "_chamfer"] = float(dp.pointCloud.distChamfer(pc.copy(), np.array(q, copy=True)))
        # D1 PSNR exactly as the datasets obtain it (encode_dataset_ehem.py:170-171): pt.pcerror shells out to the
        # reference's utils/pc_error (an executable copy under oracle/_ref/, the tree itself is read-only without +x),
        # utils.get_psnr parses section 3 of its report.  The tool prints 6 significant digits.
        import importlib
        import shutil
        ref_utils = importlib.import_module("utils")
        exe = os.path.join(HERE, "_ref", "pc_error")
        if not os.path.exists(exe):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            shutil.copyfile(os.path.join(ref_shims.REF_ROOT, "utils", "pc_error"), exe)
            os.chmod(exe, 0o755)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            os.makedirs("temp/data", exist_ok=True)
            rep = os.path.join(tmp, name + "_pcerror.txt")
            peak = "59.70" if kind == "kitti" else "30000"
            dp.pointCloud.pcerror(pc, np.asarray(q), None, "-r " + peak, rep, pcerror_path=exe)
            txt = open(rep).read()
            out[name + "_mseF"] = float([l for l in txt.splitlines() if "mseF      (p2point)" in l][0].split()[-1])
            out[name + "_psnr"] = float(ref_utils.get_psnr(rep)[0])
        finally:
            os.chdir(cwd)
        print(name, "pc_error D1 PSNR", out[name + "_psnr"], "mseF", out[name + "_mseF"])
        print(name, "quantised cloud", out[name + "_q"].shape, out[name + "_q"].dtype, "chamfer", out[name + "_chamfer"])
    np.savez_compressed(os.path.join(GOLD, "metrics.npz"), **out)


if __name__ == "__main__":
    what = [a for a in sys.argv[1:] if "=" not in a] or ["octree", "ehem", "octattn", "coder", "octattn_e2e", "metrics"]
    only = [a.split("=", 1)[1].split(",") for a in sys.argv[1:] if a.startswith("only=")]
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_shims.import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        if "octree" in what:
            gen_octree(ns, tmp, only[0] if only else None)
        if "ehem" in what:
            gen_ehem_logits(ns)
        if "octattn" in what:
            gen_octattn_logits(ns)
        if "coder" in what:
            gen_coder(ns, tmp)
        if "octattn_e2e" in what:
            gen_octattn_e2e(ns, tmp)
        if "metrics" in what:
            gen_metrics(ns, tmp)
