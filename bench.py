#!/usr/bin/env python
"""
bench.py — output Mpix/s of IMP's fused decoded-pixel chain on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic frames. The default workload is
BASELINE.json configs[1] (cfg2): 256 requests of a 3840x2160 BGRA frame -> crop 3600x2025 -> INTER_AREA
to 800x450 -> 256x64 watermark at opacity 60, i.e. one fused kernel launch over 256 jobs.

  value      whole-job Mpix/s with the frames already resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the public C ABI with HOST (pinned) buffers: H2D of every request's
             crop window + kernel + D2H of every result inside the timed region
  roofline   algorithmic bytes / measured launch time / MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's CPU path (its own C for filters/compositing compiled unmodified + cv2 for
             the OpenCV calls; or the C port when those are absent) on this box's host cores

Multi-GPU: one process per GPU (torchrun), every rank runs the same per-GPU workload on its own frames
(weak scaling; frames are independent, there is no collective on the data path).
`--impl reference` times the reference CPU implementation only (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------------------
# Workloads. Each returns: list of (src_shape(h,w,c), request kwargs, n_jobs_with_this_shape), cfg kwargs
# ------------------------------------------------------------------------------------------------------------
def watermark(seed, h, w):
    rng = np.random.default_rng(seed)
    wm = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    wm[:, :, 3] = np.linspace(0, 255, w).astype(np.uint8)[None, :]
    return wm


def workload(name: str, scale: int = 1):
    if name == "cfg1":
        return dict(desc="1920x1080x3 -> resize=640,360 (AREA 3x3), batch 256", jobs=[((1080, 1920, 3), dict(resize="640,360"), 256 // scale)], cfg=dict())
    if name == "cfg1l":   # cfg1 "as worded": bilinear (shim-level extension, not a reference call site)
        return dict(desc="1920x1080x3 -> resize=640,360 INTER_LINEAR (extension), batch 256", jobs=[((1080, 1920, 3), dict(resize="640,360", interp=1), 256 // scale)], cfg=dict())
    if name == "cfg3nn":  # cfg3 reference-faithful: GIF output forces INTER_NN (bridge.c:594)
        return dict(desc="200 frames 480x270x4 -> resize=960,540,up (NN, GIF output) + modulate=0,0,100 + colorize=704214,0.6",
                    jobs=[((270, 480, 4), dict(resize="960,540,up", simple=True, filters=["modulate=0,0,100", "colorize=704214,0.6"]), 200 // scale)], cfg=dict())
    if name == "cfg2":
        return dict(desc="256 x [3840x2160x4 -> crop=3600px,2025px,c,c -> resize=800,450 (AREA 4.5x) -> 256x64 watermark r,b,10,10 opacity 60]",
                    jobs=[((2160, 3840, 4), dict(crop="3600px,2025px,c,c", resize="800,450"), 256 // scale)],
                    cfg=dict(watermark=watermark(3, 64, 256), wm_gravity_x="r", wm_gravity_y="b", wm_offset_x=10, wm_offset_y=10, wm_opacity=60))
    if name == "cfg3":
        return dict(desc="200 frames 480x270x4 -> resize=960,540,up (CUBIC 2x) + modulate=0,0,100 + colorize=704214,0.6",
                    jobs=[((270, 480, 4), dict(resize="960,540,up", filters=["modulate=0,0,100", "colorize=704214,0.6"]), 200 // scale)], cfg=dict())
    if name == "cfg4":
        return dict(desc="8 x [4000x3000x3 -> blur=2.3 + vignette=0.8 + rotate=90]",
                    jobs=[((3000, 4000, 3), dict(filters=["blur=2.3", "vignette=0.8", "rotate=90"]), max(1, 8 // scale))], cfg=dict(allow_experiments=True))
    if name == "cfg5":
        rng = np.random.default_rng(6)
        shapes = [(int(rng.integers(240, 1537)), int(rng.integers(320, 2049)), 3 if rng.random() < 0.75 else 4) for _ in range(512)]
        n = 65536 // scale
        per = [n // 512 + (1 if i < n % 512 else 0) for i in range(512)]
        return dict(desc=f"{n} jobs from 512 mixed-size sources -> resize=256,256 (AREA) + 64x64 watermark r,b,8,8",
                    jobs=[(s, dict(resize="256,256"), k) for s, k in zip(shapes, per) if k],
                    cfg=dict(max_w=0, max_h=0, watermark=watermark(61, 64, 64), wm_gravity_x="r", wm_gravity_y="b", wm_offset_x=8, wm_offset_y=8, wm_opacity=100))
    raise SystemExit(f"unknown config {name}")


# ------------------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One nginx-worker-like process: runs `count` requests of the workload single-threaded."""
    name, scale, count, seed, use_ref = args
    from oracle import oracle as O
    wl = workload(name, scale)
    cfg = O.OracleConfig(**wl["cfg"])
    rng = np.random.default_rng(seed)
    shape, rq, _ = wl["jobs"][0]
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    ref_ok = False
    if use_ref and O.Ref.available():
        ref_ok = O.Ref.use_cv2(True)
    query = "&".join(([f"crop={rq['crop']}"] if rq.get("crop") else []) + ([f"resize={rq['resize']}"] if rq.get("resize") else []) +
                     [f"filter-{f}" for f in rq.get("filters", [])])
    t0 = time.perf_counter()
    opix = 0
    for _ in range(count):
        if ref_ok:
            code, step, out = O.Ref.run_job(query, img, cfg)
        else:
            code, step, out = O.run_chain(img, rq.get("crop"), None, rq.get("resize"), rq.get("filters", []), cfg)
        assert code == 0
        opix += out.shape[0] * out.shape[1]
    return time.perf_counter() - t0, opix, ref_ok


def cpu_run(name, scale, per_worker, workers, use_ref=True):
    """`workers` independent single-threaded processes (how nginx worker_processes scales). Returns Mpix/s, kind."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(_cpu_worker, [(name, scale, per_worker, 1000 + i, use_ref) for i in range(workers)])
    wall = max(r[0] for r in res)
    opix = sum(r[1] for r in res)
    kind = "reference" if all(r[2] for r in res) else "port"
    return opix / wall / 1e6, kind, wall


# ------------------------------------------------------------------------------------------------------------
def sample_clocks(stop, out):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    idx = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", idx], capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.15)


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = [float(s[0]) for s in samples if s[0].replace(".", "").isdigit()]
    mx = [float(s[1]) for s in samples if s[1].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(samples)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _bind_to_gpu_cpus(index: int) -> bool:
    """Multi-GPU runs: pin this rank to the CPUs NVML reports as local to its GPU before any pinned host memory is
    allocated, so the e2e leg's staging buffers are first-touched on the GPU's own NUMA node (the H2D copies of 8 ranks
    otherwise cross the socket interconnect). Best effort: silently skipped where NVML or the cpuset says no."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--scale", type=int, default=1, help="divide the job count (debug only; 1 = the named workload)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(a.config, a.scale)
    metric = "output Mpix/s (fused resize+filter chain)"

    # ---------------- reference arm: the CPU path, rank 0 only ------------------------------------------
    if a.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        per_worker = {"cfg2": 8, "cfg4": 1}.get(a.config, 16)
        for _ in range(max(0, min(a.warmup, 1))):
            cpu_run(a.config, a.scale, 1, cores)
        vals, kind = [], "port"
        t_total = 0.0
        for _ in range(a.steps):
            v, kind, wall = cpu_run(a.config, a.scale, per_worker, cores)
            vals.append(v); t_total += wall
        v = float(np.mean(vals))
        sample = f"{per_worker} request(s) per worker x {cores} single-threaded worker processes per step, same request as '{a.config}'"
        print(json.dumps({"metric": metric, "value": v, "unit": "Mpix/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * t_total / max(1, a.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "u8", "data": "synthetic", "impl": "reference", "config": {"workload": a.config + ": " + wl["desc"]},
                          "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    # ---------------- our arm -----------------------------------------------------------------------------
    import torch
    import ngx_http_imgproc_b200 as M
    from ngx_http_imgproc_b200 import api

    torch.cuda.set_device(local)
    numa_bound = _bind_to_gpu_cpus(local) if world > 1 else False
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = M.library()
    L.init(local)
    dev = torch.device("cuda", local)
    cfg = api.Config(**wl["cfg"])

    # device-resident frames: every job has its own source (inputs >> the 126 MB L2) and its own destination
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    plans, srcs, dsts, jobs = [], [], [], []
    out_pix = 0
    for shape, rq, count in wl["jobs"]:
        h, w, c = shape
        plan = L.plan(w, h, c, cfg, **rq)
        plans.append(plan)
        pitch = (w * c + 15) & ~15
        n_src = count if a.config != "cfg5" else 1
        src = torch.randint(0, 256, (n_src, h, pitch), dtype=torch.uint8, device=dev, generator=gen)
        opitch = (plan.out_w * plan.out_c + 15) & ~15
        dst = torch.empty((count, plan.out_h, opitch), dtype=torch.uint8, device=dev)
        srcs.append(src); dsts.append(dst)
        for k in range(count):
            jobs.append((plan, src[k % n_src].data_ptr(), pitch, dst[k].data_ptr(), opitch))
        out_pix += count * plan.out_w * plan.out_h
    batch = api.Batch(L)
    for j in jobs:
        batch.add(*j)
    # a real (non-default) stream: the kernels, and the CUDA events that time them, are issued on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    batch.launch(stream)                      # compiles the job table
    torch.cuda.synchronize()
    algo_bytes = batch.algorithmic_bytes
    launches_per_step = batch.launches_per_run

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        batch.launch(stream)
    barrier()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, samples), daemon=True)
    th.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s, e in evs:
        s.record(); batch.launch(stream); e.record()
    e1.record()
    barrier()
    launches = L.launch_count() - l0
    total_ms = e0.elapsed_time(e1)
    step_ms = [s.elapsed_time(e) for s, e in evs]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * out_pix * a.steps / (total_ms_max / 1e3) / 1e6

    # ---------------- end to end: host (pinned) buffers through the C ABI ---------------------------------
    # 32 distinct pinned sources per shape, cycled; every request still copies its own crop window H2D and its
    # result D2H inside the timed region.
    h_plans, h_srcs, h_dsts = [], [], []
    h2d = d2h = 0
    e2e_jobs = jobs if a.config != "cfg5" else jobs[:4096]
    hs_cache = {}
    ji = 0
    for (shape, rq, count), plan in zip(wl["jobs"], plans):
        h, w, c = shape
        cnt = count if a.config != "cfg5" else min(count, max(1, 4096 // len(wl["jobs"])))
        pool = [torch.randint(0, 256, (h, w, c), dtype=torch.uint8).pin_memory() for _ in range(min(32 if a.config != "cfg5" else 1, cnt))]
        outs = [torch.empty((plan.out_h, plan.out_w, plan.out_c), dtype=torch.uint8).pin_memory() for _ in range(min(32, cnt))]
        for k in range(cnt):
            h_plans.append(plan); h_srcs.append(pool[k % len(pool)].numpy()); h_dsts.append(outs[k % len(outs)].numpy())
            x, y, ww, hh = plan.window
            h2d += ww * hh * c; d2h += plan.out_w * plan.out_h * plan.out_c
    e2e_pix = sum(p.out_w * p.out_h for p in h_plans)
    api.run_host_batch(L, h_plans, h_srcs, h_dsts, n_streams=4)            # warm-up: allocates the lanes
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        api.run_host_batch(L, h_plans, h_srcs, h_dsts, n_streams=4)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_pix * a.e2e_steps / float(t.item()) / 1e6
    # one request at a time through the C ABI (imp_gpu_run_host: H2D of the crop window, kernel(s), D2H, sync): what a single
    # nginx request pays once its frame is decoded
    lat = []
    if rank == 0:
        same = [s_ for p_, s_ in zip(h_plans, h_srcs) if p_ is h_plans[0]][:32]
        for k in range(60):
            t1 = time.perf_counter()
            h_plans[0].run_host(same[k % len(same)], h_dsts[0])
            lat.append((time.perf_counter() - t1) * 1e3)
        lat = sorted(lat[10:])
    stop.set(); th.join(timeout=2)            # clocks were sampled across the device-timed and the end-to-end regions

    if rank != 0:
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json), scaled to
    # this launch's job count; null for workloads that have no capture yet.
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(a.config)
        if tj:
            traffic = tj["traffic_bytes_per_job"] * len(jobs)
    except Exception:
        traffic = None
    achieved = algo_bytes / (float(np.mean(step_ms)) / 1e3) / 1e9
    line = {
        "metric": metric, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": total_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 pixels, f32/i32 arithmetic", "data": "synthetic",
        "config": {"workload": a.config + ": " + wl["desc"], "jobs_per_gpu": len(jobs), "l2": "every job has its own source frame; inputs per step >> 126 MB L2" if a.config != "cfg5" else "512 sources (1.7 GB) cycled, > L2",
                   "e2e_inputs": "32 distinct pinned host frames per shape, cycled; each request copies its crop window H2D and its result D2H",
                   "rank_cpu_binding": "NVML cpu affinity of the rank's GPU" if numa_bound else "none"},
        "e2e": {"value": e2e_val, "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "jobs_per_step": len(h_plans), "steps": a.e2e_steps},
        "gpu_launches": int(launches),
        "single_request_latency_ms": {"p50": lat[len(lat) // 2], "p99": lat[-1], "n": len(lat),
                                      "what": "imp_gpu_run_host on the workload's first request shape: H2D of the crop window + kernel(s) + D2H + sync, pinned host frames"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_step": int(algo_bytes), "kernel_launches_per_step": launches_per_step,
                     "ms_per_launch_group": float(np.mean(step_ms)), "ms_min": float(np.min(step_ms))},
        "clocks": clocks_summary(samples),
    }
    if world == 1 and not a.no_cpu:
        cores = os.cpu_count() or 1
        per_worker = {"cfg2": 24, "cfg4": 2, "cfg1": 100, "cfg3": 100}.get(a.config, 8)
        try:
            v, kind, wall = cpu_run(a.config, a.scale, per_worker, cores)
            line["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": kind,
                                    "sample": f"{per_worker} request(s) x {cores} single-threaded worker processes ({wall:.1f} s), same request as '{a.config}'"}
        except Exception as ex:  # the baseline must not take the GPU numbers down with it
            line["cpu_baseline"] = {"value": None, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
