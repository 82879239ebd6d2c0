#!/usr/bin/env python
"""bench.py -- frames/s of the SCP encode-side hot path on B200 (BASELINE.json metric).

Headline workload (config.workload): BASELINE.json configs[1] -- synthetic KITTI-shaped 120k-point sweeps, spherical
coordinates, lidar_level 16, SCP-EHEM, encode_mullevel (three sub-octrees per frame), random-init+ weights.  The other four
BASELINE configs are timed too (``other_configs`` in the line; any of them becomes the headline with ``--config N``).
A step = one pass of the whole hot path over one batch of frames:
  points -> quantise/Morton/sort/octree/context (A1-A6) -> EHEM / OctAttention forward (A8-A12) -> softmax/CDF intervals
  (A13) -> coding order (A7).
`value` times that with the points already resident in HBM; `e2e` times Encoder.encode_stream() from pinned host buffers to
per-frame bitstreams (H2D points, D2H 8 B/node intervals, host range coder A14).
Frames are independent: at N GPUs every rank encodes its own frames (weak scaling, no collective on the path); config 5
(1000-frame batch) is a fixed total split over the ranks (strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl reference]

``--impl reference`` runs the UNMODIFIED reference (oracle/run_reference.py: the snapshot of /root/reference that
__graft_entry__.build() packs into oracle/_ref/reference.zip) on the host cores: one whole frame measured once
(`measured_s`, cached in /tmp for the later runs on the same box), every step a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/s (encode path, level-16 KITTI-shape)"
# BASELINE.json configs[0..4]
CONFIGS = {
    1: dict(model="EHEM", level=12, mode="spher", mullevel=False, kind="kitti", n_points=120000, frames_per_step=16,
            workload="synthetic 120k-point 64-beam KITTI-shaped sweep, spherical octree lidar_level 12, SCP-EHEM encode (encode.py)"),
    2: dict(model="EHEM", level=16, mode="spher", mullevel=True, kind="kitti", n_points=120000, frames_per_step=4,
            workload="kitti-shaped 120k-pt sweep, spherical, lidar_level 16, SCP-EHEM encode_mullevel (3 sub-octrees/frame)"),
    3: dict(model="EHEM", level=17, mode="spher", mullevel=False, kind="ford", n_points=80000, frames_per_step=4,
            workload="Ford-shaped ~80k-point sweep (integer mm), spherical lidar_level 17, SCP-EHEM encode (encode.py)"),
    4: dict(model="OctAttention", level=14, mode="spher", mullevel=False, kind="kitti", n_points=120000, frames_per_step=4,
            workload="KITTI-shaped 120k-pt sweep, spherical lidar_level 14, SCP-OctAttention encode (encode.py compress, windows of 1024)"),
    5: dict(model="EHEM", level=14, mode="cylin", mullevel=False, kind="kitti", n_points=120000, frames_per_step=8,
            frames_total=1000,
            workload="KITTI-shaped 120k-pt sweeps, cylindrical (--cylin) lidar_level 14, SCP-EHEM encode, batch of 1000 frames"),
}
L2_NOTE = "inputs/activations per step >> 126 MB L2 (no explicit flush needed)"
PARITY_SEED = 1000


def model_cfg(name):
    NS = types.SimpleNamespace
    if name == "EHEM":
        return NS(model=NS(context_size=8192, token_num=255, max_level=19), train=NS(type="kitti"), data=NS(extra_pos=False))
    return NS(model=NS(max_octree_level=12, context_size=1024, token_num=255, layer_num=3, head_num=4, abs_pos_embed_dim=12,
                       occ_embed_dim=128, level_embed_dim=6, octant_embed_dim=4, hidden_dimension=300, pos_embed=True),
              train=NS(type="kitti", dropout=0.0))


def make_frames(c, n, seed0):
    from scp_b200 import synth
    gen = synth.kitti_sweep if c["kind"] == "kitti" else synth.ford_sweep
    return [gen(seed0 + i, c["n_points"]) for i in range(n)]


def parity_frame(c):
    """The frame both arms encode for the bpp comparison (and the reference arm times): seed 1000 of the workload, as
    generated (NOT guard-banded: numpy's SVML float32 arctan2/arccos flip the quantised bin of a few tens of the 120 000
    points against the correctly rounded CUDA path, DESIGN.md "Float stage", so node counts may differ by a few tens)."""
    return np.ascontiguousarray(make_frames(c, 1, PARITY_SEED)[0])


def public_config(cid):
    """The `config` object: identical in both arms (what is measured, not how)."""
    c = CONFIGS[cid]
    return {"workload": c["workload"], "baseline_config": cid, "model": c["model"], "lidar_level": c["level"],
            "coordinates": c["mode"], "mullevel": c["mullevel"], "points_per_frame": c["n_points"],
            "frames_total": c.get("frames_total"), "l2": L2_NOTE, "weights": "random-init+ (seeded)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, gpu):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.p.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass

    def stop(self):
        if self.p:
            self.p.terminate()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/run_reference.py) or, without it, the oracle port
# ------------------------------------------------------------------------------------------------
def port_frame_time(c, seed=PARITY_SEED, gpu_model=None):
    """Oracle port (numpy octree + plain-torch fp32 EHEM) on a bounded sample of ONE frame: the octrees in full, the model on
    one full and one quarter window; t(m) = a*m + b*m^2 (Linear/attention terms ~ tokens, the three kNN graphs ~ tokens^2)
    fitted through the two and summed over all windows of the frame.  With ``gpu_model`` the two sample windows are also
    checked against the GPU by the explained-rows criterion (tests/parity_explain.py) -> parity_maxabs."""
    import torch
    from oracle import ehem_torch as O
    from oracle import octree_np as onp
    from scp_b200 import synth, weights as W
    from scp_b200.octree import MULLEVEL_PATHS
    torch.set_num_threads(os.cpu_count() or 1)
    pts = make_frames(c, 1, seed)[0]
    qf = synth.KITTI_QS if c["kind"] == "kitti" else synth.FORD_QS
    t0 = time.time()
    levels = []
    jobs = list(enumerate(MULLEVEL_PATHS)) if c["mullevel"] else [(0, None)]
    for i, mp in jobs:
        q = onp.quantize(pts[:, :3], qf(c["level"] + i), c["mode"])["q"]
        rows = onp.tree_rows(q, morton_path=list(mp) if mp else None, drop_last=c["mullevel"])["rows"]
        ids, poss, _, data, _ = onp.ehem_level_split(rows, c["level"], mullevel=c["mullevel"])
        levels += list(zip(data, poss))
    t_tree = time.time() - t0
    windows = [(d[s:s + 8192], p[:, s:s + 8192]) for d, p in levels for s in range(0, len(d), 8192)]
    lens = np.array([len(w[0]) for w in windows], np.float64)
    full = [w for w in windows if len(w[0]) == 8192][0]
    quarter = (full[0][:2048], full[1][:, :2048])
    sd = W.synth_state_dict(W.ehem_spec(19), 0, True)
    ts, parity = [], None
    for d, p in (full, quarter):
        dt, pt = torch.from_numpy(np.ascontiguousarray(d)), torch.from_numpy(np.ascontiguousarray(p))
        t0 = time.time()
        l1, l2 = O.ehem_forward(sd, dt, pt)
        onp.pmf_to_cdf_u16(torch.softmax(torch.cat((l1, l2)), 1).numpy())
        ts.append(time.time() - t0)
        if gpu_model is not None:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from parity_explain import check_explained_parity
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                rep = check_explained_parity(gpu_model, sd, dt, pt, what="bench window")
            parity = max(parity or 0.0, rep["arith_max"])
    m1, m2 = 8192.0, 2048.0
    b = (ts[0] / m1 - ts[1] / m2) / (m1 - m2)
    a = ts[0] / m1 - b * m1
    if b < 0 or a < 0:                                       # noisy fit: fall back to proportional scaling
        a, b = ts[0] / m1, 0.0
    t_model = float((a * lens + b * lens * lens).sum())
    desc = (f"oracle port, 1 frame: numpy octree x{len(jobs)} measured in full ({t_tree:.1f}s); torch-fp32 EHEM+softmax+CDF "
            f"measured on one 8192- and one 2048-token window ({ts[0]:.2f}s, {ts[1]:.2f}s), t(m)=a*m+b*m^2 summed over the "
            f"frame's {len(windows)} windows")
    return t_tree + t_model, desc, parity


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cid = args.config
    c = CONFIGS[cid]
    cores = os.cpu_count() or 1
    from oracle import run_reference as rr
    key = f"config{cid}"
    base = {"impl": "reference", "metric": METRIC, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if cid == 5 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": public_config(cid)}
    if not rr.available():
        if c["model"] != "EHEM":
            print(json.dumps({"impl": "reference", "unavailable": "reference snapshot (oracle/_ref/reference.zip) absent"}))
            return
        sec, desc, _ = port_frame_time(c)
        v = 1.0 / sec
        base.update({"value": v, "ms_per_step": sec * 1e3,
                     "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
                     "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
        return
    import torch
    root, src = rr.materialise()
    ns, shims = rr._import(root)                                   # untimed: imports + numpyAc JIT build
    torch.set_num_threads(cores)
    cached = rr.cached_measurement(key)
    if cached is None:
        m = rr.encode_frame(c, parity_frame(c), ns, shims, threads=cores)     # ONE WHOLE FRAME, measured once
        rr.store_measurement(key, m)
    else:
        m = cached
    # the K steps: a bounded sample of the same work per step (one context window through the reference's own model),
    # so that the run stays within minutes; `value` comes from the whole-frame measurement above
    model = rr._load_model(ns, shims, c["model"])
    g = torch.Generator().manual_seed(0)
    if c["model"] == "EHEM":
        n_tok = 2048
        data = torch.stack((torch.randint(1, 12, (1, n_tok, 4), generator=g), torch.randint(1, 9, (1, n_tok, 4), generator=g),
                            torch.randint(0, 255, (1, n_tok, 4), generator=g)), -1)
        pos = torch.rand((1, 3, n_tok), generator=g)
        step = lambda: model(data, pos, enc=True)
        step_desc = "one 2048-token context window through the reference's EHEM.forward (fp32, all host threads)"
    else:
        n_tok = 1024
        data = torch.stack((torch.randint(0, 255, (1, n_tok, 4), generator=g), torch.randint(1, 12, (1, n_tok, 4), generator=g),
                            torch.randint(1, 9, (1, n_tok, 4), generator=g)), -1)
        pos = torch.rand((1, n_tok, 4, 3), generator=g)
        step = lambda: model(data.clone(), pos)
        step_desc = "one 1024-token window through the reference's OctAttention.forward (fp32, all host threads)"
    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        step_ms = (time.perf_counter() - t0) / max(args.steps, 1) * 1e3
    v = 1.0 / m["seconds"]
    sample = (f"UNMODIFIED reference ({src}) on {m['threads']} host threads: ONE WHOLE frame of the workload (seed {PARITY_SEED}, "
              f"{m['n_points']} points, {m['n_nodes']} nodes) measured once: preproc {m['stages_s']['preproc']:.1f}s "
              f"+ dataset {m['stages_s']['dataset']:.1f}s + compress {m['stages_s']['compress']:.1f}s (model forward "
              f"{m['stages_s']['model_forward']:.1f}s) = {m['seconds']:.1f}s; not extrapolated")
    base.update({"value": v, "ms_per_step": step_ms, "step_is": step_desc + "; `value` = 1 / measured_s of the whole frame",
                 "measured_s": m["seconds"], "frames_measured": 1, "extrapolated": False,
                 "cached_from_previous_run_on_this_box": cached is not None, "stages_s": m["stages_s"],
                 "parity_frame": {"seed": PARITY_SEED, "guard_banded": False, "n_points": m["n_points"], "n_nodes": m["n_nodes"],
                                  "bytes": m["bytes"], "bpp": m["bpp"]},
                 "cpu_baseline": {"value": v, "unit": "frames/s", "cores": m["threads"], "kind": "reference", "sample": sample},
                 "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class Harness:
    def __init__(self):
        import torch
        import torch.distributed as dist
        from scp_b200 import _lib
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.lib = _lib.require_device()
        self.models = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(vals, device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, vals):
        t = self.torch.tensor(vals, device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def model(self, name):
        if name not in self.models:
            from scp_b200.models import EHEM, OctAttention
            self.models[name] = (EHEM if name == "EHEM" else OctAttention)(model_cfg(name)).cuda()
        return self.models[name]

    def encoder(self, c):
        from scp_b200.encoder import Encoder
        return Encoder(self.model(c["model"]), c["level"], c["mode"], mullevel=c["mullevel"], kind=c["kind"])

    def time_config(self, cid, steps, warmup, frames_per_step=None):
        """Device-resident and end-to-end frames/s of one BASELINE config.  Returns a dict (times are max over ranks)."""
        torch = self.torch
        c = CONFIGS[cid]
        enc = self.encoder(c)
        F = frames_per_step or c["frames_per_step"]
        if c.get("frames_total"):
            return self.time_frame_batch(cid, enc, F)
        frames = make_frames(c, F, 100 * self.rank)              # every rank has its own frames (frame-wise partition)
        offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int64)
        host = torch.from_numpy(np.concatenate(frames, 0)).pin_memory()
        xyz = host.cuda()
        # L2 is flushed between steps by construction: one step streams GBs of activations (>> 126 MB L2)
        for _ in range(warmup):
            enc.encode_device(xyz, offs)
        self.barrier()
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        launches0 = self.lib.scp_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            interval, fr, infos, _ = enc.encode_device(xyz, offs)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        launches = self.lib.scp_launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        n_nodes = sum(f[1] for f in fr)
        # end to end through the public API: host points in, bitstreams out.  Encoder.encode_stream() pipelines batches
        # (the host range coder of batch i overlaps the GPU work of batch i+1); the timed region covers the WHOLE stream
        # including pipeline fill and drain, every batch paying its H2D copy, D2H of the intervals and the range coder.
        enc.encode(frames)
        self.barrier()
        e2e_steps = max(4, 2 * steps)
        t0 = time.time()
        res = None
        for res in enc.encode_stream(frames for _ in range(e2e_steps)):
            pass
        torch.cuda.synchronize()
        e2e_s = (time.time() - t0) / e2e_steps
        self.barrier()
        ms, e2e_ms = self.max_over_ranks([ms, e2e_s * 1e3])
        W = self.world
        return {"enc": enc, "frames": frames, "xyz": xyz, "offs": offs, "host": host, "res": res, "F": F, "clocks": clocks,
                "n_nodes": n_nodes, "launches": int(launches), "ms": ms,
                "line": {"value": W * F * steps / (ms / 1e3), "unit": "frames/s", "ms_per_step": ms / steps, "steps": steps,
                         "warmup": warmup, "frames_per_step_per_gpu": F, "nodes_per_step": int(n_nodes),
                         "gpu_launches": int(launches), "scaling": "weak",
                         "e2e": {"value": W * F / (e2e_ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(host.numel() * 4),
                                 "d2h_bytes_per_step": int(n_nodes * 8), "bpp_mean": float(np.mean([r.bpp for r in res])),
                                 "api": "Encoder.encode_stream (pipelined batches)", "batches_timed": e2e_steps}}}

    def time_frame_batch(self, cid, enc, F):
        """Config 5: a FIXED batch of ``frames_total`` frames split frame-wise over the ranks (strong scaling), encoded end to
        end (host points -> bitstreams) in pipelined batches of F; frames cycle through a pool of 16 synthetic seeds."""
        torch = self.torch
        c = CONFIGS[cid]
        from scp_b200 import partition
        total = c["frames_total"]
        pool = make_frames(c, 16, 0)
        mine = partition.frames_for_rank(total, self.rank, self.world)
        batches = [[pool[i % 16] for i in mine[a:a + F]] for a in range(0, len(mine), F)]
        for _ in range(3):
            enc.encode(batches[0])
        self.barrier()
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        launches0 = self.lib.scp_launch_count()
        t0 = time.time()
        nbytes = npts = nodes = 0
        for res in enc.encode_stream(iter(batches)):
            nbytes += sum(len(r.bitstream) for r in res)
            npts += sum(r.n_points for r in res)
            nodes += sum(r.n_nodes for r in res)
        torch.cuda.synchronize()
        s = time.time() - t0
        launches = self.lib.scp_launch_count() - launches0
        self.barrier()
        clocks = sampler.stop() if sampler else None
        (s,) = self.max_over_ranks([s])
        nbytes, npts, nodes = self.sum_over_ranks([nbytes, npts, nodes])
        fps = total / s
        h2d = int(sum(f.size * 4 for f in batches[0]))
        return {"enc": enc, "frames": batches[0], "res": res, "F": F, "clocks": clocks, "n_nodes": int(nodes), "ms": s * 1e3,
                "launches": int(launches), "xyz": None,
                "line": {"value": fps, "unit": "frames/s", "ms_per_step": s * 1e3 / max(len(batches), 1), "steps": len(batches),
                         "warmup": 3, "frames_per_step_per_gpu": F, "frames_total": total, "nodes_total": int(nodes),
                         "gpu_launches": int(launches), "scaling": "strong",
                         "timing": "wall clock around the whole pipelined stream (host points -> bitstreams), max over ranks",
                         "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(8 * nodes / max(len(batches), 1) / self.world),
                                 "bpp_mean": 8.0 * nbytes / npts, "api": "Encoder.encode_stream (pipelined batches)",
                                 "batches_timed": len(batches)}}}


def kernel_profile(h, r, steps):
    """Per-kernel durations: the same steps again with every model operator launch bracketed by CUDA events on the launching
    stream.  Kept out of the headline loop because ~1700 timing events per step cost 15-20 % of the step."""
    torch = h.torch
    enc, xyz, offs = r["enc"], r["xyz"], r["offs"]
    ops = enc.model.ops
    ops.reserve_events(2 * 900 * r["F"] * steps + 4096)
    ops.prof = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(steps):
        enc.encode_device(xyz, offs)
    p1.record()
    h.barrier()
    prof_ms = p0.elapsed_time(p1)
    prof, ops.prof = ops.prof, None
    agg = {}
    for tag, fl, by, a, b in prof:
        d = agg.setdefault(tag, [0.0, 0.0, 0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += fl; d[2] += by; d[3] += 1
    return agg, prof_ms


def octree_stage_report(h, r, hbm_peak):
    """Octree / gather kernels alone on a batch large enough to fill the GPU (256 frames, >> L2; SURVEY 8d: "measure on
    batches of >= 100 frames per launch" -- one frame is 1.4 MB of points)."""
    torch = h.torch
    enc, xyz, frames, F = r["enc"], r["xyz"], r["frames"], r["F"]
    OB = 256
    reps = max(1, OB // F)
    big = torch.cat([xyz] * reps)
    boffs = np.concatenate([[0], np.cumsum([len(f) for f in frames] * reps)]).astype(np.int64)
    octree_ms = None
    for _ in range(4):
        bb, tt_, _pf = enc.build_context(big, boffs)
        torch.cuda.synchronize()
        m = enc.builder.stage_ms()
        octree_ms = m if octree_ms is None else {k: min(octree_ms[k], v) for k, v in m.items()}
    strict = bb.stage_bytes()                 # SURVEY 8d per-unit figures x units, nothing else
    rep = {}
    for k, v in octree_ms.items():
        if k in strict and v > 0:
            rep[k] = {"ms": round(v, 4), "bytes": strict[k], "GBps": round(strict[k] / v / 1e6, 1),
                      "frac_of_hbm_peak": round(strict[k] / v / 1e6 / hbm_peak, 3)}
        else:
            rep[k] = {"ms": round(v, 4)}
    tree_ms = sum(octree_ms.get(k, 0.0) for k in ("heads", "emit", "occupancy"))
    total_ms = sum(octree_ms.values())
    total_b = strict["quantise"] + strict["sort"] + strict["tree"] + strict["context"]
    rep["tree_emission(heads+emit+occupancy)"] = {"ms": round(tree_ms, 4), "bytes": strict["tree"],
                                                  "GBps": round(strict["tree"] / tree_ms / 1e6, 1) if tree_ms else None,
                                                  "frac_of_hbm_peak": round(strict["tree"] / tree_ms / 1e6 / hbm_peak, 3) if tree_ms else None}
    rep["all"] = {"ms": round(total_ms, 4), "bytes": total_b, "frac_of_hbm_peak": round(total_b / total_ms / 1e6 / hbm_peak, 3)}
    out = {"batch_frames": reps * F, "point_job_pairs": int(sum(i.n_points for i in bb.infos)), "sorted_keys": bb.total_kept,
           "nodes": bb.total_rows, "byte_model": "SURVEY 8d strict: quantise 12 B/frame point + 8 B/key; sort (1+2P)*8 B per "
           "sorted key, P = digit passes; tree emission 28 B/node; context gather 60 B/node", "stages": rep}
    del big, tt_
    return out


def run_ours(args):
    h = Harness()
    torch = h.torch
    cid = args.config
    c = CONFIGS[cid]
    r = h.time_config(cid, args.steps, args.warmup, args.frames_per_step)
    line = r["line"]
    hbm_peak, tf_peak, src = peaks()
    extra = {}
    if r["xyz"] is not None:
        agg, prof_ms = kernel_profile(h, r, args.steps)
        if c["model"] == "EHEM":
            extra["octree_stages"] = octree_stage_report(h, r, hbm_peak)
    else:
        agg, prof_ms = {}, 0.0

    # the other BASELINE configs, each through the same two timings (short runs; --config N makes one the headline)
    others = {}
    if args.other_configs and h.world == 1:
        for oc in sorted(CONFIGS):
            if oc == cid:
                continue
            try:
                if CONFIGS[oc].get("frames_total"):
                    saved = CONFIGS[oc]["frames_total"]
                    CONFIGS[oc]["frames_total"] = args.sweep_frames
                    o = h.time_config(oc, 3, 3)
                    CONFIGS[oc]["frames_total"] = saved
                    o["line"]["note"] = (f"{args.sweep_frames} of the config's {saved} frames in this default run; "
                                         f"`--config {oc}` runs all {saved}")
                else:
                    o = h.time_config(oc, 3, 3)
                o["line"]["workload"] = CONFIGS[oc]["workload"]
                others[str(oc)] = o["line"]
                del o
                torch.cuda.empty_cache()
            except Exception as e:                                   # never lose the headline over a side config
                others[str(oc)] = {"error": repr(e)}

    if h.rank != 0:
        if h.world > 1:
            h.dist.destroy_process_group()
        return

    enc, res = r["enc"], r["res"]
    # roofline of the dominant kernel class
    roof = None
    kernels = {}
    if agg:
        top = max(agg.items(), key=lambda kv: kv[1][0])
        tag, (tms, tfl, tby, cnt) = top
        tensor_kernel = tag in ("linear", "swin_attention", "octattn_attention") or tag.startswith("knn_d1")
        f16 = os.environ.get("SCP_AUTO_ENGINE", "2") == "2" and os.environ.get("SCP_GEMM", "auto") == "auto"
        if tensor_kernel:
            ach = tfl / tms / 1e9
            roof = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                    "note": "ALGORITHMIC flops (2MNK of the fp32-class product) / CUDA-event time of the kernel's launches, vs the "
                            "measured sustained dense bf16 rate (kernel timed inside a long step)",
                    "pipe_util": (3.0 * ach / (tf_peak if f16 else tf_peak / 2.0)),
                    "pipe_util_note": "tensor-pipe view: the contraction runs error-compensated (x = hi + lo, 3 tcgen05.mma per "
                                      "product: " + ("3xFP16 on the kind::f16 pipe" if f16 else "3xTF32, half the bf16 rate") +
                                      "), so issued MMA flops = 3 x algorithmic"}
        else:
            roof = {"bound": "hbm", "achieved": tby / tms / 1e6, "peak": hbm_peak, "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof.update({"kernel": tag, "launches": cnt, "avg_ms": tms / cnt, "share_of_step": tms / prof_ms, "peak_source": src,
                     "traffic": None, "timed_in": "instrumented repeat of the timed steps (events around every launch)"})
        for tp in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
            p = os.path.join(ROOT, "profiles", tp)
            if os.path.exists(p):
                tr = json.load(open(p)).get(tag)
                if tr:
                    roof["traffic"] = tr["dram_bytes_per_launch_mean"]
                    roof["traffic_note"] = tr["note"] + f" ({tp})"
                    break
        kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[3] // args.steps,
                       "TFLOPs": v[1] / v[0] / 1e9 if v[0] else None, "GBps": v[2] / v[0] / 1e6 if v[0] else None,
                       "frac_of_peak": (v[1] / v[0] / 1e9 / tf_peak) if (v[0] and (k in ("linear", "swin_attention", "octattn_attention") or k.startswith("knn_d1")))
                       else ((v[2] / v[0] / 1e6 / hbm_peak) if v[0] else None)}
                   for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}

    # decode path, EHEM only
    if c["model"] == "EHEM" and r["xyz"] is not None:
        from scp_b200.decoder import Decoder
        dcd = Decoder(enc.model, c["level"], c["mode"], mullevel=c["mullevel"], kind=c["kind"])
        torch.cuda.synchronize()
        t0 = time.time()
        dres_all = dcd.decode_batch(res)                      # all frames of the last e2e batch in lock-step
        torch.cuda.synchronize()
        dec_s = (time.time() - t0) / len(res)
        dres = dres_all[0]
        xyz, offs = r["xyz"], r["offs"]
        chk_b, _t, _pf = enc.build_context(xyz[: int(offs[1])], offs[:2])
        chk_occ = chk_b.emit(("occ",), finish=False)["occ"].cpu().numpy()
        extra["decode"] = {"frames_per_s": 1.0 / dec_s, "s_per_frame": dec_s, "symbols": int(dres.n_symbols),
                           "round_trip_exact": bool(np.array_equal(np.concatenate(dres.occ), chk_occ)), "frames_in_batch": len(res),
                           "api": "Decoder.decode_batch (frames in lock-step: phase 1 per level, phase 2 + host range decoders per window index)"}
        try:                                                    # distortion report (row f-4)
            from scp_b200 import metrics
            vk = chk_b.emit(("voxel_key",), finish=True)["voxel_key"]
            pts64 = xyz[: int(offs[1]), :3].double()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                d0.record()
                terms = metrics.distortion_terms(pts64, metrics.dequantised_cloud(chk_b, vk, c["mode"]))
                d1.record()
                torch.cuda.synchronize()
            ch, mse = terms.tolist()
            extra["distortion"] = {"ms_per_frame": d0.elapsed_time(d1), "chamfer_m": ch,
                                   "d1_psnr_db": metrics.psnr_of(mse, metrics.KITTI_PEAK if c["kind"] == "kitti" else metrics.FORD_PEAK)}
        except Exception as e:
            extra["distortion"] = {"error": repr(e)}

    # parity at the bench's own size: the frame the reference arm encodes and times (seed 1000)
    pf = parity_frame(c)
    pr = enc.encode([pf])[0]
    parity = {"frame": {"seed": PARITY_SEED, "guard_banded": False, "n_points": int(pr.n_points), "n_nodes": int(pr.n_nodes),
                        "bytes": len(pr.bitstream), "bpp": pr.bpp}}
    ref_m = None
    try:
        from oracle import run_reference as rr
        ref_m = rr.cached_measurement(f"config{cid}")
    except Exception:
        pass
    if ref_m:
        parity["reference_run_on_this_box"] = {"n_nodes": ref_m["n_nodes"], "bytes": ref_m["bytes"], "bpp": ref_m["bpp"]}
        parity["nodes_equal"] = ref_m["n_nodes"] == pr.n_nodes
        if not parity["nodes_equal"]:
            parity["nodes_note"] = ("the frame is NOT guard-banded: numpy's float32 arctan2 / arccos (SVML, not correctly rounded) put a "
                                    "handful of the 120 000 points into the neighbouring bin of the correctly rounded CUDA transform "
                                    "(DESIGN.md 'Float stage'); on guard-banded inputs every integer stage is bit-exact (tests/test_octree_gpu.py)")
        parity["bpp_dev"] = abs(pr.bpp - ref_m["bpp"]) / ref_m["bpp"]
    cores = os.cpu_count() or 1
    if ref_m:
        cpu = {"value": 1.0 / ref_m["seconds"], "unit": "frames/s", "cores": ref_m["threads"], "kind": "reference",
               "sample": f"whole frame through the unmodified reference, measured once by `bench.py --impl reference` on this box "
                         f"({ref_m['seconds']:.1f}s; cached in /tmp)"}
        if c["model"] == "EHEM" and args.cpu_parity:
            _s, _d, pm = port_frame_time(c, gpu_model=enc.model)
            parity["pmf_maxabs_vs_oracle_2_windows"] = pm
    elif c["model"] == "EHEM":
        cpu_s, cpu_desc, pm = port_frame_time(c, gpu_model=enc.model if args.cpu_parity else None)
        cpu = {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": cores, "kind": "port", "sample": cpu_desc}
        parity["pmf_maxabs_vs_oracle_2_windows"] = pm
    else:
        cpu = {"value": None, "unit": "frames/s", "cores": cores, "kind": "port", "sample": "no port timing for OctAttention; run --impl reference first"}
    if "pmf_maxabs_vs_oracle_2_windows" in parity:
        parity["pmf_criterion"] = ("explained rows (tests/parity_explain.py): max-abs PMF error over EVERY row of one 8192- and one "
                                   "2048-token window of the bench frame vs the oracle run on the device's kNN sets; sets verified "
                                   "k-nearest up to float32 ties")

    out = {"metric": METRIC, "value": line["value"], "unit": "frames/s", "n_gpus": h.world, "steps": line["steps"],
           "warmup": line["warmup"], "ms_per_step": line["ms_per_step"], "higher_is_better": True, "scaling": line["scaling"],
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": public_config(cid),
           "setup": {"frames_per_step_per_gpu": line["frames_per_step_per_gpu"], "nodes_per_step": line.get("nodes_per_step", line.get("nodes_total")),
                     "gemm_engine": os.environ.get("SCP_GEMM", "auto"),
                     "auto_engine": {"0": "fp32 simt", "1": "3xTF32 tcgen05", "2": "3xFP16 tcgen05"}[os.environ.get("SCP_AUTO_ENGINE", "2")]},
           "e2e": line["e2e"], "gpu_launches": line["gpu_launches"], "clocks": r["clocks"], "roofline": roof, "kernels": kernels,
           "parity": parity, "cpu_baseline": cpu, "other_configs": others}
    out.update(extra)
    print(json.dumps(out))
    if h.world > 1:
        h.dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (1-based); 2 = headline")
    ap.add_argument("--frames-per-step", type=int, default=None)
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false", help="skip the short runs of the other configs")
    ap.add_argument("--sweep-frames", type=int, default=96, help="frames of config 5 in the default (side) run")
    ap.add_argument("--no-cpu-parity", dest="cpu_parity", action="store_false", help="skip the oracle-vs-GPU window check")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
