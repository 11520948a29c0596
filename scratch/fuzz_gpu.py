"""Random request fuzz against the oracle (GPU box). Usage: python scratch/fuzz_gpu.py [n] [seed]"""
import sys, numpy as np, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import ngx_http_imgproc_b200 as M
from ngx_http_imgproc_b200 import api
from oracle import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
L = M.library(); L.init(0)
FILT = ["flip=10", "flip=01", "rotate=90", "rotate=180", "rotate=270", "modulate=30,120,90", "colorize=336699,0.4", "gamma=0.7", "contrast=1.3",
        "gradmap=102030,f0e0d0", "vignette=0.7", "vignette=1.5,0.6", "gotham=1", "lomo=1", "kelvin=1", "rainbow=mid", "scanline=0.6,0.2,2,1",
        "blur=0.6", "blur=1.4", "blur=2.3", "blur=3.9", "blur=5.5"]
bad = 0; t0 = time.time()
for it in range(n):
    c = int(rng.choice([1, 3, 4])); h = int(rng.integers(1, 400)); w = int(rng.integers(1, 500))
    if rng.random() < 0.1: h, w = int(rng.integers(400, 1500)), int(rng.integers(400, 2200))
    img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    rq = {}
    if rng.random() < 0.4:
        cw, ch = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
        rq["crop"] = f"{cw}px,{ch}px,{int(rng.integers(0, w - cw + 1))}px,{int(rng.integers(0, h - ch + 1))}px"
    if rng.random() < 0.7:
        dw, dh = int(rng.integers(1, 600)), int(rng.integers(1, 500))
        rq["resize"] = f"{dw},{dh}" + (",up" if rng.random() < 0.5 else "")
        if rng.random() < 0.15: rq["simple"] = True
        if rng.random() < 0.15: rq["interp"] = 1
    k = int(rng.integers(0, 5))
    rq["filters"] = [str(f) for f in rng.choice(FILT, k)]
    if rng.random() < 0.3: rq["flatten"] = True
    if rng.random() < 0.2: rq["pack"] = int(rng.choice([24, 32]))
    kw = dict(allow_experiments=True, max_filters=8, max_w=0, max_h=0)
    if rng.random() < 0.4:
        wc = int(rng.choice([3, 4]))
        kw.update(watermark=rng.integers(0, 256, (int(rng.integers(1, 60)), int(rng.integers(1, 80)), wc), dtype=np.uint8),
                  wm_gravity_x=str(rng.choice(list("lcr"))), wm_gravity_y=str(rng.choice(list("tcb"))),
                  wm_offset_x=int(rng.integers(-20, 30)), wm_offset_y=int(rng.integers(-20, 30)), wm_opacity=int(rng.integers(1, 101)))
    o = dict(rq); interp = o.pop("interp", 0)
    c2, s2, ref = O.run_chain(img, o.get("crop"), None, o.get("resize"), o.get("filters", []), O.OracleConfig(**kw), o.get("simple", False),
                              o.get("flatten", False), linear=bool(interp), pack=o.get("pack", 0))
    code, step, plan = L.try_plan(w, h, c, api.Config(**kw), **rq)
    if code != c2 or (code and step != s2):
        bad += 1; print("CODE", it, (h, w, c), rq, code, step, c2, s2); continue
    if code: continue
    out = plan.run_host(img); plan.close()
    tol = 1 if any("vignette" in f for f in rq["filters"]) else 0
    d = np.abs(out.astype(int) - ref.astype(int)) if out.shape == ref.shape else None
    if d is None or d.max() > tol:
        bad += 1; print("PIX", it, (h, w, c), rq, {k: v for k, v in kw.items() if k != "watermark"}, None if d is None else (int(d.max()), int((d > tol).sum())))
print("fuzz done", n, "bad", bad, "%.1fs" % (time.time() - t0))
