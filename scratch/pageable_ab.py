"""Batch e2e with PAGEABLE host frames (what cvDecodeImage hands RunJob) against pinned ones, cfg2 and cfg1 shapes."""
import sys, time, statistics, numpy as np, torch
sys.path.insert(0, ".")
import ngx_http_imgproc_b200 as M
from ngx_http_imgproc_b200 import api
import bench
L = M.library(); L.init(0)
for name, n in (("cfg2", 64), ("cfg1", 256), ("cfg4", 8)):
    wl = bench.workload(name, 1)
    (h, w, c), rq, _ = wl["jobs"][0]
    cfg = api.Config(**wl["cfg"])
    p = L.plan(w, h, c, cfg, **rq)
    rng = np.random.default_rng(1)
    for kind in ("pinned", "pageable"):
        if kind == "pinned":
            keep = [torch.randint(0, 256, (h, w, c), dtype=torch.uint8).pin_memory() for _ in range(min(n, 16))]
            srcs = [t.numpy() for t in keep]
            keepo = [torch.empty((p.out_h, p.out_w, p.out_c), dtype=torch.uint8).pin_memory() for _ in range(min(n, 16))]
            dsts = [t.numpy() for t in keepo]
        else:
            srcs = [rng.integers(0, 256, (h, w, c), dtype=np.uint8) for _ in range(min(n, 16))]
            dsts = [np.empty((p.out_h, p.out_w, p.out_c), np.uint8) for _ in range(min(n, 16))]
        hj = api.HostJobs([p] * n, [srcs[k % len(srcs)] for k in range(n)], [dsts[k % len(dsts)] for k in range(n)])
        hj.run(L); ts = []
        for _ in range(6):
            t0 = time.perf_counter(); hj.run(L); ts.append(time.perf_counter() - t0)
        ms = statistics.median(ts) * 1e3
        x, y, ww, hh = p.window
        print(f"{name} {kind}: {ms:.1f} ms for {n} jobs = {n*p.out_w*p.out_h/ms/1e3:.0f} Mpix/s, h2d {n*ww*hh*c/ms/1e6:.1f} GB/s")
