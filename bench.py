#!/usr/bin/env python
"""bench.py -- frames/s of the SCP encode-side hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] -- synthetic KITTI-shaped 120k-point sweeps, spherical
coordinates, lidar_level 16, SCP-EHEM, encode_mullevel (three sub-octrees per frame), random-init+ weights.
A step = one pass of the whole hot path over one batch of frames:
  points -> quantise/Morton/sort/octree/context (A1-A6) -> EHEM forward (A8-A11) -> softmax/CDF intervals (A13)
  -> coding order (A7).
`value` times that with the points already resident in HBM; `e2e` times Encoder.encode() from pinned host
buffers to per-frame bitstreams (H2D points, D2H 8 B/node intervals, host range coder A14) .
Frames are independent: at N GPUs every rank encodes its own frames (weak scaling, no collective on the path).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
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

LEVEL = 16
N_POINTS = 120000
FRAMES_PER_STEP = 4
WORKLOAD = "kitti-shaped 120k-pt sweep, spherical, lidar_level 16, SCP-EHEM encode_mullevel (3 sub-octrees/frame)"


def cfg_ehem():
    NS = types.SimpleNamespace
    return NS(model=NS(context_size=8192, token_num=255, max_level=19), train=NS(type="kitti"), data=NS(extra_pos=False))


def make_frames(n, seed0):
    from scp_b200 import synth
    return [synth.kitti_sweep(seed0 + i, N_POINTS) for i in range(n)]


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
# CPU baseline: the oracle port (numpy octree + plain-torch fp32 EHEM) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_frame_time(n_sample_windows=2, seed=1000):
    """Times the CPU restatement of the reference path on a bounded sample of ONE frame of the workload and
    scales to the whole frame.  Returns (seconds_per_frame, description)."""
    import torch
    from oracle import ehem_torch as O
    from oracle import octree_np as onp
    from scp_b200 import synth, weights as W
    from scp_b200.octree import MULLEVEL_PATHS
    torch.set_num_threads(os.cpu_count() or 1)
    pts = synth.kitti_sweep(seed, N_POINTS)
    t0 = time.time()
    levels = []
    for i, mp in enumerate(MULLEVEL_PATHS):
        q = onp.quantize(pts[:, :3], synth.KITTI_QS(LEVEL + i), "spher")["q"]
        rows = onp.tree_rows(q, morton_path=list(mp), drop_last=True)["rows"]
        ids, poss, _, data, _ = onp.ehem_level_split(rows, LEVEL, mullevel=True)
        levels += list(zip(data, poss))
    t_tree = time.time() - t0
    windows = [(d[s:s + 8192], p[:, s:s + 8192]) for d, p in levels for s in range(0, len(d), 8192)]
    total_tok = sum(len(w[0]) for w in windows)
    full = [w for w in windows if len(w[0]) == 8192][:n_sample_windows]
    sd = W.synth_state_dict(W.ehem_spec(19), 0, True)
    t0 = time.time()
    tok = 0
    for d, p in full:
        l1, l2 = O.ehem_forward(sd, torch.from_numpy(np.ascontiguousarray(d)), torch.from_numpy(np.ascontiguousarray(p)))
        pm = torch.softmax(torch.cat((l1, l2)), 1).numpy()
        onp.pmf_to_cdf_u16(pm)
        tok += len(d)
    t_model = (time.time() - t0) * total_tok / max(tok, 1)
    desc = (f"1 frame: numpy octree x3 sub-octrees measured in full ({t_tree:.1f}s); torch-fp32 EHEM+softmax+CDF measured on "
            f"{len(full)} of {len(windows)} windows ({tok} of {total_tok} tokens) and scaled by tokens")
    return t_tree + t_model, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    desc = ""
    for i in range(args.warmup + args.steps):
        t, desc = cpu_reference_frame_time(1, seed=1000 + i)
        if i >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    v = 1.0 / sec
    cores = os.cpu_count() or 1
    print(json.dumps({
        "impl": "reference", "metric": "frames/s (encode path, level-16 KITTI-shape)", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": 1},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from scp_b200 import _lib
    from scp_b200.encoder import Encoder
    from scp_b200.models import EHEM

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.require_device()
    model = EHEM(cfg_ehem()).cuda()
    enc = Encoder(model, LEVEL, "spher", mullevel=True, kind="kitti")
    F = args.frames_per_step
    frames = make_frames(F, 100 * rank)                   # every rank has its own frames (frame-wise partition)
    offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int64)
    host = torch.from_numpy(np.concatenate(frames, 0)).pin_memory()
    xyz = host.cuda()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # L2 is flushed between steps by construction: one step streams > 10 GB of activations (>> 126 MB L2)
    for _ in range(args.warmup):
        enc.encode_device(xyz, offs)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.scp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        interval, fr, infos, _ = enc.encode_device(xyz, offs)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.scp_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    n_nodes = sum(f[1] for f in fr)

    # per-kernel durations: the same K steps again with every operator launch bracketed by CUDA events on the launching
    # stream.  Kept out of the headline loop because ~1700 timing events per step cost 15-20 % of the step.
    model.ops.reserve_events(2 * 700 * F * args.steps + 4096)
    model.ops.prof = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        enc.encode_device(xyz, offs)
    p1.record()
    barrier()
    prof_ms = p0.elapsed_time(p1)
    prof, model.ops.prof = model.ops.prof, None

    # octree / gather kernels alone on a batch large enough to fill the GPU (256 frames = 768 jobs, >> L2; SURVEY 8d:
    # "measure on batches of >= 100 frames per launch" -- one frame is 1.4 MB of points)
    OB = 256
    big = torch.cat([xyz] * (OB // F + 1))[: 0 + sum(len(f) for f in frames) * (OB // F)] if F <= OB else xyz
    reps = OB // F if F <= OB else 1
    boffs = np.concatenate([[0], np.cumsum([len(f) for f in frames] * reps)]).astype(np.int64)
    octree_ms = None
    for _ in range(4):
        bb, tt_, _pf = enc.build_context(big, boffs)
        torch.cuda.synchronize()
        m = enc.builder.stage_ms()
        octree_ms = m if octree_ms is None else {k: min(octree_ms[k], v) for k, v in m.items()}
    o_pts = int(boffs[-1]) * 3
    o_kept = bb.total_kept
    o_bytes = bb.stage_bytes()
    o_nodes = bb.total_rows
    o_depth = max(i.depth for i in bb.infos)
    del big, tt_

    # end to end through the public API: host points in, bitstreams out.  Encoder.encode_stream() pipelines batches (the
    # host range coder of batch i overlaps the GPU work of batch i+1); the timed region covers the WHOLE stream including
    # pipeline fill and drain, every batch paying its H2D copy, D2H of the intervals and the range coder.
    enc.encode(frames)
    barrier()
    e2e_steps = max(4, 2 * args.steps)
    t0 = time.time()
    res = None
    for res in enc.encode_stream(frames for _ in range(e2e_steps)):
        pass
    torch.cuda.synchronize()
    e2e_s = (time.time() - t0) / e2e_steps
    barrier()

    tt = torch.tensor([ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, e2e_ms = tt.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # per-kernel totals from CUDA events recorded around every launch of the model operators
    agg = {}
    for tag, fl, by, a, b in prof:
        d = agg.setdefault(tag, [0.0, 0.0, 0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += fl; d[2] += by; d[3] += 1
    top = max(agg.items(), key=lambda kv: kv[1][0])
    hbm_peak, tf_peak, src = peaks()
    tag, (tms, tfl, tby, cnt) = top
    # Dense contractions run error-compensated (x = x_hi + x_lo, three tcgen05.mma per fp32-class product), so the tensor-pipe
    # work is 3x the nominal 2MNK.  nn.Linear runs the split on the FP16 pipe (kind::f16: peak = measured sustained bf16
    # rate; SCP_AUTO_ENGINE=1 selects the 3xTF32 form, half that rate); attention and kNN use the 3xTF32 form.
    tensor_kernel = tag in ("linear", "swin_attention") or tag.startswith("knn_d1")
    f16 = tag == "linear" and os.environ.get("SCP_AUTO_ENGINE", "2") == "2" and os.environ.get("SCP_GEMM", "auto") == "auto"
    if tensor_kernel:
        roof = {"bound": "tensor", "achieved": 3.0 * tfl / tms / 1e9, "peak": tf_peak if f16 else tf_peak / 2.0, "unit": "TFLOP/s",
                "note": ("FP16 MMA flops issued (3 per fp32-class product, 3xFP16 split) vs measured sustained bf16 peak" if f16 else
                         "TF32 MMA flops issued (3 per fp32-class product) vs TF32 peak = measured sustained bf16 / 2")}
    else:
        roof = {"bound": "hbm", "achieved": tby / tms / 1e6, "peak": hbm_peak, "unit": "GB/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof.update({"kernel": tag, "launches": cnt, "avg_ms": tms / cnt, "share_of_step": tms / prof_ms, "peak_source": src,
                 "traffic": None, "timed_in": "instrumented repeat of the timed steps (events around every launch)"})
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/r01_ncu_traffic.json):
    # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches of that kernel
    tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tp):
        tr = json.load(open(tp)).get(tag)
        if tr:
            roof["traffic"] = tr["dram_bytes_per_launch_mean"]
            roof["traffic_note"] = tr["note"]
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[3] // args.steps,
                   "TFLOPs": v[1] / v[0] / 1e9 if v[0] else None, "GBps": v[2] / v[0] / 1e6 if v[0] else None}
               for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    bm = o_bytes          # OctreeBuilder.stage_bytes(): SURVEY 8d figures x the units each stage really processes
    oct_rep = {k: {"ms": round(v, 4), "GBps": round(bm[k] / v / 1e6, 1) if v > 0 else None,
                   "frac_of_hbm_peak": round(bm[k] / v / 1e6 / hbm_peak, 3) if v > 0 else None} for k, v in octree_ms.items()}

    # decode path: one frame of the last e2e batch back through Decoder, checked against the encoder's
    # own octree (lossless round trip) -- reported next to the encode numbers, not part of `value`
    from scp_b200.decoder import Decoder
    dcd = Decoder(model, LEVEL, "spher", mullevel=True, kind="kitti")
    torch.cuda.synchronize()
    t0 = time.time()
    dres_all = dcd.decode_batch(res)                      # all frames of the last e2e batch in lock-step
    torch.cuda.synchronize()
    dec_s = (time.time() - t0) / len(res)
    dres = dres_all[0]
    chk_b, _t, _pf = enc.build_context(xyz[: int(offs[1])], offs[:2])
    chk_occ = chk_b.emit(("occ",), finish=False)["occ"].cpu().numpy()
    dec_ok = bool(np.array_equal(np.concatenate(dres.occ), chk_occ))
    decode_rep = {"frames_per_s": 1.0 / dec_s, "s_per_frame": dec_s, "symbols": int(dres.n_symbols), "round_trip_exact": dec_ok,
                  "frames_in_batch": len(res),
                  "api": "Decoder.decode_batch (frames in lock-step: phase 1 per level, phase 2 + host range decoders per window index)"}

    # distortion report: Chamfer distance / D1 PSNR of the first frame against its dequantised voxels,
    # timed with CUDA events; reported next to the encode numbers, not part of `value`
    try:
        from scp_b200 import metrics
        vk = chk_b.emit(("voxel_key",), finish=True)["voxel_key"]
        pts64 = xyz[: int(offs[1]), :3].double()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            d0.record()
            terms = metrics.distortion_terms(pts64, metrics.dequantised_cloud(chk_b, vk, "spher"))
            d1.record()
            torch.cuda.synchronize()
        ch, mse = terms.tolist()
        dist_rep = {"ms_per_frame": d0.elapsed_time(d1), "chamfer_m": ch, "d1_psnr_db": metrics.psnr_of(mse, metrics.KITTI_PEAK),
                    "api": "metrics.dequantised_cloud + metrics.distortion_terms (scp_dequantise_keys, 2 x scp_nn_dist2, exact FP64)"}
    except Exception as e:                                 # never lose the bench line over the side report
        dist_rep = {"error": repr(e)}

    cpu_s, cpu_desc = cpu_reference_frame_time(2)
    fps = world * F * args.steps / (ms / 1e3)
    e2e_fps = world * F / (e2e_ms / 1e3)
    out = {
        "metric": "frames/s (encode path, level-16 KITTI-shape)", "value": fps, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": F, "nodes_per_step": n_nodes,
                   "l2": "inputs/activations per step >> 126 MB L2 (no explicit flush needed)",
                   "weights": "random-init+ (seeded)", "gemm_engine": os.environ.get("SCP_GEMM", "auto"), "auto_engine": {"0": "fp32 simt", "1": "3xTF32 tcgen05", "2": "3xFP16 tcgen05"}[os.environ.get("SCP_AUTO_ENGINE", "2")]},
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(host.numel() * 4),
                "d2h_bytes_per_step": int(n_nodes * 8), "bpp_mean": float(np.mean([r.bpp for r in res])),
                "api": "Encoder.encode_stream (pipelined batches)", "batches_timed": e2e_steps},
        "gpu_launches": int(launches),
        "decode": decode_rep,
        "distortion": dist_rep,
        "clocks": clocks,
        "roofline": roof,
        "kernels": kernels,
        "octree_stages": {"batch_frames": reps * F, "point_job_pairs": o_pts, "sorted_keys": o_kept, "nodes": o_nodes,
                          "algorithmic_bytes": o_bytes, "stages": oct_rep},
        "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": cpu_desc},
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=FRAMES_PER_STEP)
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
