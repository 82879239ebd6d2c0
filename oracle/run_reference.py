"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference (luoao-kddi/SCP) end to end on the host cores.

Used by ``bench.py --impl reference`` (and its ``cpu_baseline`` leg when a cached measurement exists).  Never imported by
the product (``scp_b200``).

The reference is a plain Python source tree plus one prebuilt ``Octree_python_lib.so``; it has no packaging metadata, so it
cannot be pip-installed, and ``/root/reference`` does not exist on the GPU box.  ``snapshot()`` -- called by
``__graft_entry__.build()`` in the build container -- therefore packs the tree, from where it lies, into ONE archive
``oracle/_ref/reference.zip`` (git-ignored build output, like ``oracle/_ref/torch_ext``; no reference source enters the
repository or its history).  ``materialise()`` unpacks it into a temporary directory at run time and the modules are
imported from there through ``oracle/ref_shims.py`` (stubs for the uninstalled ``pytorch_lightning`` / ``hydra`` / ``h5py`` /
``plyfile`` and two unused ``transformers`` names); ``.cuda()`` is patched to the identity, everything runs in fp32 on
``torch.get_num_threads()`` host threads exactly as ``encode.py`` / ``encode_mullevel.py`` drive it:

    config 1/3/5  proc_pc (C++ octree .so + gen_K_parent_seq)  -> EncodeEHEMDataset.__getitem__ -> encode.compress_ehem
    config 2      mul_proc_pc x3 (pure-Python mullevel octree) -> mullevel EncodeEHEMDataset     -> encode_mullevel.compress_ehem
    config 4      proc_pc                                      -> EncodeDataset.__getitem__      -> encode.compress (OctAttention)

Only the dataset's ``preproc`` hook is replaced (by the very ``proc_pc`` / ``mul_proc_pc`` calls it makes, minus the
``pc_error`` / Chamfer distortion report, which is not part of the encode metric): the reference's own hook passes
unsupported keyword arguments (encode_dataset_ehem.py:159-169 vs data_preprocess.py:13-25) or hands ``pc_error`` a ``.bin``
file it cannot read (encode_dataset_ehem_mullevel.py:189) -- SURVEY.md section 8b "known reference defects".
Weights: the seeded "random-init+" tensors of scp_b200/weights.py loaded into the reference's modules (names/shapes
asserted), the same ones the CUDA path uses, so the bitstreams of both arms are comparable (bpp parity at bench size).
"""
import importlib
import json
import os
import shutil
import sys
import tempfile
import time
import types
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ZIP = os.path.join(HERE, "_ref", "reference.zip")
CACHE = os.path.join(tempfile.gettempdir(), "scp_b200_reference_cache.json")


def snapshot(ref_root="/root/reference", out=ZIP):
    """Packs the reference tree into oracle/_ref/reference.zip (build container only).  Returns the path or None."""
    if not os.path.isfile(os.path.join(ref_root, "encode.py")):
        return None
    os.makedirs(os.path.dirname(out), exist_ok=True)
    newest = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(ref_root) for f in fs)
    if os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    with zipfile.ZipFile(out + ".tmp", "w", zipfile.ZIP_DEFLATED) as z:
        for d, _, fs in os.walk(ref_root):
            if ".git" in d.split(os.sep):
                continue
            for f in fs:
                p = os.path.join(d, f)
                z.write(p, os.path.relpath(p, ref_root))
    os.replace(out + ".tmp", out)
    return out


def materialise():
    """Directory holding the reference tree: a fresh unpack of the snapshot (GPU box), else /root/reference."""
    if os.path.exists(ZIP):
        d = tempfile.mkdtemp(prefix="scp_ref_")
        with zipfile.ZipFile(ZIP) as z:
            z.extractall(d)
        for exe in ("utils/pc_error",):
            p = os.path.join(d, exe)
            if os.path.exists(p):
                os.chmod(p, 0o755)
        return d, "oracle/_ref/reference.zip (snapshot of /root/reference taken by __graft_entry__.build)"
    if os.path.isfile("/root/reference/encode.py"):
        return "/root/reference", "/root/reference"
    return None, None


def available():
    return os.path.exists(ZIP) or os.path.isfile("/root/reference/encode.py")


def _import(ref_root):
    os.environ["SCP_REFERENCE_ROOT"] = ref_root
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_shims
    ref_shims.REF_ROOT = ref_root
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(tempfile.gettempdir(), "scp_ref_torch_ext"))
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    ns = ref_shims.import_reference()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="scp_ref_cwd_"))          # the reference mkdirs ./temp where it runs
    try:
        ns.encode = importlib.import_module("encode")          # JIT-builds numpyAc_backend with g++ (~30 s, untimed)
        ns.encode_mullevel = importlib.import_module("encode_mullevel")
    finally:
        os.chdir(cwd)
    return ns, ref_shims


def _load_model(ns, ref_shims, name):
    import torch
    from scp_b200 import weights
    if name == "EHEM":
        cfg = ref_shims.make_cfg("ehem")
        m = ns.ehem.EHEM(cfg)
        spec = weights.ehem_spec(cfg.model.max_level)
    else:
        cfg = ref_shims.make_cfg("oct", train_type="kitti")
        m = ns.oct_attention.OctAttention(cfg)
        spec = weights.octattn_spec()
    assert [k for k, _, _ in spec] == list(m.state_dict().keys()), "state_dict names differ from the reference"
    m.load_state_dict(weights.synth_state_dict(spec, seed=0, sharpen=True), strict=True)
    m.cfg = cfg
    return m.eval()


def encode_frame(cfg, points, ns=None, ref_shims=None, threads=None):
    """One frame through the reference's encode path.  cfg: dict(model, level, mode, mullevel, kind).  Returns timings (s),
    bitstream bytes, bpp, node count."""
    import torch
    from torch.utils.data import default_collate
    if ns is None:
        root, _ = materialise()
        ns, ref_shims = _import(root)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    dp = ns.data_preprocess
    kind, level, mode = cfg["kind"], cfg["level"], cfg["mode"]
    qs = (lambda L: 400.0 / (2 ** L - 1)) if kind == "kitti" else (lambda L: float(2 ** (18 - L)))
    kw = dict(spher=(mode == "spher"), cylin=(mode == "cylin"))
    tmp = tempfile.mkdtemp(prefix="scp_ref_frame_")
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        binf = os.path.join(tmp, "seq", "frame.bin")
        os.makedirs(os.path.dirname(binf))
        np.ascontiguousarray(points, np.float32).tofile(binf)
        model = _load_model(ns, ref_shims, cfg["model"])
        args = types.SimpleNamespace(spher=(mode == "spher"), cylin=(mode == "cylin"), sequential=False)
        t = {}
        t0 = time.perf_counter()
        if cfg["mullevel"]:
            res = [dp.mul_proc_pc(binf, tmp, "frame", qs=qs(level + i), test=True, normalize=False, morton_path=mp, **kw)
                   for i, mp in enumerate(([0, 0], [0, 1], [1]))]
            files, pc, bin_num = [r[0] for r in res], res[0][2], res[0][3]
            t["preproc"] = time.perf_counter() - t0
            ds = ns.ds_ehem_mul.EncodeEHEMDataset([binf], 8192, kind, True, level, mode == "cylin", mode == "spher", "")
            ds.preproc = lambda f: (files, pc, 0.0, bin_num, 0.0, 0.0)
            mod, fn = ns.encode_mullevel, "compress_ehem"
        else:
            res = dp.proc_pc(binf, tmp, "frame", qs=qs(level), test=True, normalize=False, **kw)
            t["preproc"] = time.perf_counter() - t0
            if cfg["model"] == "EHEM":
                ds = ns.ds_ehem.EncodeEHEMDataset([binf], 8192, kind, True, level, mode == "cylin", mode == "spher", False, False, "")
                if mode == "cylin":
                    ds.preproc = lambda f: (res[0], res[2], 0.0, res[3], res[4][0, 2], 0.0)
                else:
                    ds.preproc = lambda f: (res[0], res[2], 0.0, res[3], 0.0)
                mod, fn = ns.encode, "compress_ehem"
            else:
                ds = ns.ds_oct.EncodeDataset([binf], 1024, kind, False, level, True, "")
                ds.preproc = lambda f: (res[0], res[2], 0.0, res[3], 0.0)
                mod, fn = ns.encode, "compress"
        t1 = time.perf_counter()
        batch = default_collate([ds[0]])[:-2]
        t["dataset"] = time.perf_counter() - t1
        t2 = time.perf_counter()
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            bpp, model_s = getattr(mod, fn)(batch, os.path.join(tmp, "out", "frame"), model, args)
        t["compress"] = time.perf_counter() - t2
        t["model_forward"] = float(model_s)
        total = time.perf_counter() - t0
        out = [f for f in os.listdir(os.path.join(tmp, "out")) if f.endswith(".bin")]
        nbytes = os.path.getsize(os.path.join(tmp, "out", out[0]))
        n_nodes = int(batch[4].shape[1]) if cfg["model"] == "EHEM" else int(batch[3].shape[1])
        return {"seconds": total, "stages_s": t, "bytes": nbytes, "bpp": float(bpp), "n_nodes": n_nodes,
                "n_points": int(len(points)), "threads": torch.get_num_threads()}
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


def cached_measurement(key):
    try:
        return json.load(open(CACHE)).get(key)
    except Exception:
        return None


def store_measurement(key, value):
    try:
        d = json.load(open(CACHE))
    except Exception:
        d = {}
    d[key] = value
    with open(CACHE, "w") as f:
        json.dump(d, f)


if __name__ == "__main__":
    print(snapshot())
