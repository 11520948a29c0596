"""End-to-end rates of requests whose results have rows that are not a multiple of 16 bytes (2-D D2H copies), pinned in and out.
Used to A/B a dense-packing kernel + linear D2H (no gain: 26.8 vs 26.1, 22.0 vs 22.6, 32.4 vs 32.0 GB/s per direction; dropped)."""
import sys, time, statistics, numpy as np, torch
sys.path.insert(0, ".")
import ngx_http_imgproc_b200 as M
from ngx_http_imgproc_b200 import api
L = M.library(); L.init(0)
cfg = api.Config(max_w=0, max_h=0)
for (h, w, c, rq, n) in [(1000, 1001, 3, dict(filters=["flip=10"]), 64), (500, 501, 3, dict(filters=["flip=10"]), 256), (3000, 4000, 3, dict(filters=["rotate=90"]), 8),
                         (2160, 3840, 4, dict(resize="1001,563"), 64)]:
    p = L.plan(w, h, c, cfg, **rq)
    srcs = [torch.randint(0, 256, (h, w, c), dtype=torch.uint8).pin_memory() for _ in range(min(n, 16))]
    dsts = [torch.empty((p.out_h, p.out_w, p.out_c), dtype=torch.uint8).pin_memory() for _ in range(min(n, 16))]
    hj = api.HostJobs([p] * n, [srcs[k % len(srcs)].numpy() for k in range(n)], [dsts[k % len(dsts)].numpy() for k in range(n)])
    hj.run(L); ts = []
    for _ in range(8):
        t0 = time.perf_counter(); hj.run(L); ts.append(time.perf_counter() - t0)
    ms = statistics.median(ts) * 1e3
    print(f"{h}x{w}x{c} {rq} n={n}: {ms:.2f} ms  out row {p.out_w*p.out_c} B x {p.out_h}  d2h {n*p.out_w*p.out_c*p.out_h/ms/1e6:.1f} GB/s  h2d {n*h*w*c/ms/1e6:.1f} GB/s")
