"""GPU tests of the host-facing machinery round 2 added under the C ABI: the plan cache, the upload-once overlay, the
chunked host batches (one grouped launch per chunk), the asynchronous submit/wait pair, the size-aware farm, and the
strip kernels' stage hand-off under slow rotated stores (the race round 2 found). Every result is compared with the
oracle or with a single-request run of the same plan."""
import numpy as np
import pytest

from conftest import rnd_image, smooth_image
from ngx_http_imgproc_b200 import api
from test_planner_host import _oracle

pytestmark = pytest.mark.gpu


def _same(a, b, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(a, b), (what, int(np.abs(a.astype(int) - b.astype(int)).max()), int((a != b).sum()))


def test_plan_cache_serves_repeated_requests(gpu, orc):
    """bridge.c:346-372 re-parses every request; here an identical (request, geometry, config) is lowered once."""
    gpu.plan_cache_clear()
    wm = rnd_image(3, 16, 40, 4)
    kw = dict(watermark=wm, wm_gravity_x="r", wm_gravity_y="b", wm_offset_x=3, wm_offset_y=2, wm_opacity=60)
    rq = dict(crop="200px,100px,c,c", resize="50,25", filters=["gamma=1.2"])
    s0 = gpu.plan_cache_stats()
    p1 = gpu.plan(320, 200, 4, api.Config(**kw), **rq)
    p2 = gpu.plan(320, 200, 4, api.Config(**kw), **rq)          # a NEW Config object with equal content
    assert p1.h.value == p2.h.value                              # the same immutable plan, reference counted
    s1 = gpu.plan_cache_stats()
    assert s1["hits"] == s0["hits"] + 1 and s1["misses"] == s0["misses"] + 1
    # anything that changes the lowering is a different key
    others = [gpu.plan(320, 200, 4, api.Config(**dict(kw, wm_opacity=61)), **rq),
              gpu.plan(320, 200, 4, api.Config(**dict(kw, watermark=rnd_image(4, 16, 40, 4))), **rq),
              gpu.plan(320, 200, 3, api.Config(**kw), **rq),
              gpu.plan(320, 200, 4, api.Config(**kw), **dict(rq, filters=["gamma=1.3"])),
              gpu.plan(320, 200, 4, api.Config(**kw), **dict(rq, flatten=True))]
    assert len({p.h.value for p in others} | {p1.h.value}) == 6
    img = smooth_image(1, 200, 320, 4)
    ref = _oracle(orc, img, rq, kw)[2]
    _same(p1.run_host(img), ref)
    p1.close()
    _same(p2.run_host(img), ref, "the second reference keeps the plan alive")
    gpu.plan_cache_clear()
    _same(p2.run_host(img), ref, "and so does a caller's reference after the cache dropped its own")
    p2.close()
    for p in others:
        p.close()


def test_cache_eviction_keeps_live_plans_valid(gpu, orc):
    gpu.plan_cache_clear()
    img = smooth_image(2, 90, 120, 3)
    first = gpu.plan(120, 90, 3, api.Config(max_w=0, max_h=0), resize="33,21")
    for k in range(140):                                         # more than the 128 entries of the cache
        gpu.plan(120, 90, 3, api.Config(max_w=0, max_h=0), resize=f"{20 + k},21").close()
    assert gpu.plan_cache_stats()["entries"] <= 128
    _same(first.run_host(img), _oracle(orc, img, dict(resize="33,21"), dict(max_w=0, max_h=0))[2])
    first.close()


def test_overlay_is_uploaded_once_and_shared(gpu, orc):
    """PrepareWatermark decodes once per configuration (bridge.c:199-237): every plan with the same overlay content
    shares one device copy, registered explicitly or on first use; a different overlay at a recycled host address
    must NOT alias it (the registry is keyed by content)."""
    kw = dict(wm_gravity_x="c", wm_gravity_y="c", wm_offset_x=2, wm_offset_y=-1, wm_opacity=80)
    img = smooth_image(3, 120, 160, 4)
    buf = np.zeros((20, 30, 4), np.uint8)
    for seed in (11, 12, 13):
        buf[...] = rnd_image(seed, 20, 30, 4)                    # same address, new content
        cfg = api.Config(watermark=buf, **kw)
        gpu.upload_watermark(cfg)
        for rq in (dict(resize="80,60"), dict(filters=["rotate=90"]), dict(resize="200,150,up")):
            out = gpu.run(img, cfg, **rq)
            _same(out, _oracle(orc, img, rq, dict(kw, watermark=buf.copy()))[2], (seed, rq))


def test_chunked_host_batches_group_their_launches(gpu, orc):
    """A GIF album / a burst of requests through imp_gpu_batch_run_host: chunks of jobs run as one launch per kernel variant
    (counted), results equal single runs; pinned and pageable buffers, odd steps, mixed plans, several passes."""
    import torch
    kw = dict(allow_experiments=True, max_filters=8, max_w=0, max_h=0)
    frames = [smooth_image(40 + i, 135, 240, 4) for i in range(48)]
    plan = gpu.plan(240, 135, 4, api.Config(**kw), resize="480,270,up", filters=["modulate=0,0,100", "colorize=704214,0.6"])
    dsts = [np.zeros((plan.out_h, plan.out_w, plan.out_c), np.uint8) for _ in frames]
    l0 = gpu.launch_count()
    api.run_host_batch(gpu, [plan] * len(frames), frames, dsts, n_streams=4)
    launches = gpu.launch_count() - l0
    assert launches <= 8, launches                               # 48 frames: 4 chunks of 12, one launch each (was 48)
    ref = _oracle(orc, frames[7], dict(resize="480,270,up", filters=["modulate=0,0,100", "colorize=704214,0.6"]), kw)[2]
    _same(dsts[7], ref)
    for f, d in zip(frames, dsts):
        _same(d, plan.run_host(f))
    # mixed plans (several passes, different kernels), pinned sources with short rows / wide rows, a strided destination
    rqs = [dict(resize="100,60"), dict(filters=["blur=1.5", "gamma=1.2"]), dict(crop="150px,100px,10px,5px"), dict(resize="64,64", filters=["blur=6", "kelvin=1"]),
           dict(filters=["rotate=270"]), dict(resize="333,200,up", interp=1)]
    imgs, plans, outs, refs = [], [], [], []
    for i in range(23):
        rq = rqs[i % len(rqs)]
        h, w, c = 120 + 3 * i, 200 + 5 * i, 3 if i % 2 else 4
        im = smooth_image(i, h, w, c)
        if i % 3 == 0:
            t = torch.from_numpy(im).pin_memory(); im = t.numpy(); imgs.append((im, t))
        else:
            imgs.append((im, None))
        p = gpu.plan(w, h, c, api.Config(**kw), **rq)
        plans.append(p)
        o = np.zeros((p.out_h, p.out_w * p.out_c + (8 if i % 4 == 1 else 0)), np.uint8)
        outs.append(o)
        refs.append(_oracle(orc, im, rq, kw)[2])
    api.run_host_batch(gpu, plans, [i[0] for i in imgs], outs, n_streams=3)
    for p, o, r, rq in zip(plans, outs, refs, rqs * 4):
        _same(o[:, :p.out_w * p.out_c].reshape(p.out_h, p.out_w, p.out_c), r, rq)
    for p in plans:
        p.close()
    plan.close()


def test_submit_wait_is_asynchronous_and_exact(gpu, orc):
    kw = dict(max_w=0, max_h=0)
    imgs = [smooth_image(i, 400, 600, 3) for i in range(12)]
    rq = dict(resize="150,100", filters=["contrast=1.4"])
    plan = gpu.plan(600, 400, 3, api.Config(**kw), **rq)
    dsts = [np.zeros((plan.out_h, plan.out_w, 3), np.uint8) for _ in imgs]
    jobs = api.HostJobs([plan] * len(imgs), imgs, dsts)
    t1 = jobs.submit(gpu, n_streams=2)
    dsts2 = [np.zeros_like(d) for d in dsts]
    jobs2 = api.HostJobs([plan] * len(imgs), imgs, dsts2)
    t2 = jobs2.submit(gpu, n_streams=2)                          # two batches in flight (they serialise on the lanes)
    api.HostJobs.wait(gpu, t1)
    api.HostJobs.wait(gpu, t2)
    ref = [_oracle(orc, im, rq, kw)[2] for im in imgs]
    for d, d2, r in zip(dsts, dsts2, ref):
        _same(d, r); _same(d2, r)
    plan.close()


def test_farm_policies_match_single_runs(gpu, orc):
    """imp_gpu_farm_run_host_policy on every GPU the box has (the driver's box has one; the 2+ GPU run is
    test_farm_on_all_gpus): round-robin and size-aware assignments give the same bytes as single runs."""
    kw = dict(max_w=0, max_h=0)
    rng = np.random.default_rng(5)
    imgs = [smooth_image(i, int(rng.integers(100, 500)), int(rng.integers(120, 700)), 3 if i % 4 else 4) for i in range(21)]
    rq = dict(resize="96,96")
    plans = [gpu.plan(im.shape[1], im.shape[0], im.shape[2], api.Config(**kw), **rq) for im in imgs]
    ref = [p.run_host(im) for p, im in zip(plans, imgs)]
    _same(ref[3], _oracle(orc, imgs[3], rq, kw)[2])
    for n_gpus in sorted({1, gpu.device_count()}):
        for policy in (api.FARM_ROUND_ROBIN, api.FARM_SIZE_AWARE):
            dsts = [np.zeros_like(r) for r in ref]
            api.run_host_batch(gpu, plans, imgs, dsts, n_streams=2, n_gpus=n_gpus, policy=policy)
            for d, r in zip(dsts, ref):
                _same(d, r, (n_gpus, policy))
    for p in plans:
        p.close()


def test_farm_on_all_gpus(gpu, orc):
    """Product-level multi-GPU test: skipped below 2 devices."""
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs 2+ GPUs on the box")
    kw = dict(max_w=0, max_h=0, watermark=rnd_image(9, 12, 12, 4), wm_gravity_x="r", wm_gravity_y="b", wm_offset_x=1, wm_offset_y=1)
    imgs = [smooth_image(i, 300 + 11 * i, 500 + 7 * i, 3 if i % 3 else 4) for i in range(40)]
    rq = dict(resize="128,128")
    plans = [gpu.plan(im.shape[1], im.shape[0], im.shape[2], api.Config(**kw), **rq) for im in imgs]
    ref = [p.run_host(im) for p, im in zip(plans, imgs)]
    for policy in (api.FARM_ROUND_ROBIN, api.FARM_SIZE_AWARE):
        dsts = [np.zeros_like(r) for r in ref]
        api.run_host_batch(gpu, plans, imgs, dsts, n_streams=2, n_gpus=n, policy=policy)
        for d, r in zip(dsts, ref):
            _same(d, r, policy)
    gpu.set_device(0)
    for p in plans:
        p.close()


def test_rotated_three_byte_stores_do_not_outrun_the_ring(gpu, orc):
    """Round-2 regression: with rotate=90/270 and 3-byte pixels the strip kernels' epilogue issues scattered byte stores;
    the plain-gather consumers then released a ring stage with their own loads still queued, and the TMA refill of 8 tiles
    later overtook them (tiles 5 .. n-9 of a column came out wrong, differently on every run). Needs >= 14 tile rows."""
    for shape in [(256, 256, 3), (512, 384, 3), (300, 200, 1)]:
        img = rnd_image(1, *shape)
        for f in (["rotate=90"], ["rotate=270"], ["flip=01", "rotate=270"]):
            for rq in (dict(filters=f), dict(resize=f"{shape[1] * 2},{shape[0] * 2},up", simple=True, filters=f), dict(resize=f"{shape[1] - 7},{shape[0] - 5}", interp=1, filters=f)):
                ref = _oracle(orc, img, rq, dict(max_w=0, max_h=0))[2]
                for rep in range(3):
                    _same(gpu.run(img, api.Config(max_w=0, max_h=0), **rq), ref, (shape, rq, rep))


def test_bounds_assertions_stay_silent_under_the_fuzz(orc):
    """compute-sanitizer is closed on the pool, so the tile kernels carry their own bounds assertions
    (-DIMP_DEBUG_BOUNDS, libimp_gpu_dbg.so built by __graft_entry__.build()): the request matrix, the resize fuzz, odd sizes
    under all orientations and the BASELINE shapes run over the debug build in a child process; no assertion may fire and
    the results still equal the oracle's."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dbg = os.path.join(root, "ngx_http_imgproc_b200", "libimp_gpu_dbg.so")
    if not os.path.exists(dbg):
        pytest.skip("debug library not built (python -m ngx_http_imgproc_b200.build --debug)")
    code = r"""
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import ctypes as C
import ngx_http_imgproc_b200 as M
from ngx_http_imgproc_b200 import api
from oracle import oracle as O
from conftest import rnd_image, smooth_image
from test_planner_host import REQS, _oracle
L = M.library(); L.init(0)
L.lib.imp_gpu_debug_flags.restype = C.c_uint
wm = rnd_image(7, 12, 20, 4)
kw = dict(allow_experiments=True, max_filters=8, max_w=0, max_h=0, watermark=wm, wm_gravity_x="r", wm_gravity_y="b", wm_offset_x=3, wm_offset_y=2, wm_opacity=60)
n = 0
def check(img, rq, kwx=kw):
    global n
    code, step, p = L.try_plan(img.shape[1], img.shape[0], img.shape[2], api.Config(**kwx), **rq)
    c2, s2, ref = _oracle(O, img, rq, kwx)
    assert code == c2, (rq, code, c2)
    if code: return
    out = p.run_host(img); p.close()
    d = np.abs(out.astype(int) - ref.astype(int)).max()
    assert d <= (1 if any("vignette" in f for f in rq.get("filters", [])) else 0), (rq, int(d))
    n += 1
for (h, w, c) in [(60, 80, 3), (45, 64, 4), (33, 47, 1), (97, 33, 3), (131, 259, 4)]:
    img = smooth_image(h + w, h, w, c)
    for rq in REQS:
        check(img, rq)
orient = [[], ["flip=10"], ["rotate=90"], ["rotate=270"], ["flip=01", "rotate=270"]]
rng = np.random.default_rng(3)
for it in range(120):
    sw, sh, c = int(rng.integers(1, 300)), int(rng.integers(1, 200)), int(rng.choice([1, 3, 4]))
    dw, dh = int(rng.integers(1, 400)), int(rng.integers(1, 300))
    img = rnd_image(it, sh, sw, c)
    fl = orient[it % len(orient)] + (["blur=%.1f" % (0.5 + (it % 7) * 0.6)] if it % 4 == 0 and c > 1 else [])
    check(img, dict(resize=f"{dw},{dh},up", filters=fl, simple=bool(it % 5 == 0), interp=int(it % 3 == 1)), dict(max_w=0, max_h=0))
check(smooth_image(1, 1080, 1920, 3), dict(resize="640,360"))
check(smooth_image(2, 2160, 3840, 4), dict(crop="3600px,2025px,c,c", resize="800,450"))
check(smooth_image(3, 270, 480, 4), dict(resize="960,540,up", filters=["modulate=0,0,100", "colorize=704214,0.6"]))
check(smooth_image(4, 750, 1000, 3), dict(filters=["blur=2.3", "vignette=0.8", "rotate=90"]))
flags = L.lib.imp_gpu_debug_flags()
print("BOUNDS", n, hex(flags))
assert flags == 0, hex(flags)
"""
    env = dict(os.environ, IMP_GPU_LIB=dbg)
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "BOUNDS" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    assert int(r.stdout.split("BOUNDS")[1].split()[0]) > 300


def test_workers_forked_before_cuda_init(orc):
    """nginx forks its workers after the configuration (and the module's shared library) is loaded, and OnEnvStart runs in
    each worker (module.c:100-107): libimp_gpu.so must tolerate being loaded — but not initialised — before fork(), and
    every child must be able to create its own CUDA context, serve requests and shut down. (A CUDA context itself can not
    cross fork(), which is why imp_gpu_init is not called at load time.)"""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import os, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import ngx_http_imgproc_b200 as M
from ngx_http_imgproc_b200 import api
from oracle import oracle as O
from conftest import smooth_image
L = M.library()                                   # master process: library loaded, plans validated, NO imp_gpu_init
cfg = api.Config(max_w=0, max_h=0, watermark=smooth_image(9, 8, 12, 4), wm_gravity_x="r", wm_gravity_y="b", wm_opacity=70)
assert L.try_plan(64, 48, 3, cfg, resize="0,0")[0] == 50
img = smooth_image(1, 240, 320, 4)
rq = dict(crop="300px,200px,c,c", resize="150,100", filters=["gamma=1.2"])
ref = O.run_chain(img, rq["crop"], None, rq["resize"], rq["filters"], O.OracleConfig(max_w=0, max_h=0, watermark=cfg.watermark, wm_gravity_x="r", wm_gravity_y="b", wm_opacity=70))[2]
pids = []
for w in range(3):                                # three "worker processes"
    pid = os.fork()
    if pid == 0:
        try:
            L.init(0)                             # OnEnvStart
            for k in range(3):
                out = L.run(img, cfg, **rq)
                assert np.array_equal(out, ref)
            L.shutdown()                          # OnEnvDestroy
            os._exit(0)
        except BaseException as e:
            sys.stderr.write("worker %d: %r\n" % (w, e)); os._exit(1)
    pids.append(pid)
bad = [os.waitpid(p, 0)[1] for p in pids]
assert all(os.WIFEXITED(s) and os.WEXITSTATUS(s) == 0 for s in bad), bad
print("FORK OK")
"""
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FORK OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_large_frames_and_many_tiny_jobs(gpu):
    """Maximum-size edge of the path: frames of 100-200 MP through every tile kernel (byte offsets beyond 2^31 in the
    destination, thousands of tiles per job), checked against size-independent properties (numpy index maps, flat fields),
    and a 3000-job host batch of tiny mixed shapes (more live plans than the plan cache holds)."""
    cfg = api.Config(max_w=0, max_h=0, max_filters=8, allow_experiments=True)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (12000, 16000, 3), dtype=np.uint8)
    assert np.array_equal(gpu.run(img, cfg, filters=["rotate=90"]), np.rot90(img, -1))
    assert np.array_equal(gpu.run(img, cfg, filters=["flip=11"]), img[::-1, ::-1])
    assert np.array_equal(gpu.run(img, cfg, crop="15000px,11000px,777px,555px"), img[555:11555, 777:15777])
    small = np.ascontiguousarray(img[:3000, :4000])
    assert np.array_equal(gpu.run(small, cfg, resize="8000,6000,up", simple=True), np.repeat(np.repeat(small, 2, 0), 2, 1))
    del img
    flat = np.full((12000, 18000, 4), 93, np.uint8)
    out = gpu.run(flat, cfg, resize="1800,1200")
    assert out.shape == (1200, 1800, 4) and (out == 93).all()
    assert (gpu.run(np.ascontiguousarray(flat[:9000, :9000]), cfg, filters=["blur=3"]) == 93).all()
    out = gpu.run(np.ascontiguousarray(flat[:4000, :6000]), cfg, resize="12000,8000,up")
    assert out.shape == (8000, 12000, 4) and (out == 93).all()
    del flat, out
    imgs = [rng.integers(0, 256, (17 + i % 13, 23 + i % 7, 3 + (i % 2)), dtype=np.uint8) for i in range(3000)]
    plans = [gpu.plan(im.shape[1], im.shape[0], im.shape[2], cfg, resize="9,7") for im in imgs]
    dsts = [np.zeros((p.out_h, p.out_w, p.out_c), np.uint8) for p in plans]
    api.run_host_batch(gpu, plans, imgs, dsts, n_streams=4)
    for d, p, im in list(zip(dsts, plans, imgs))[::97]:
        assert np.array_equal(d, p.run_host(im))
    for p in plans:
        p.close()


def test_results_with_unaligned_rows(gpu, orc):
    """Results whose rows are not a multiple of 16 bytes (the device pitch is): pageable destinations padded like an IplImage,
    dense and strided pinned ones, a one-pixel-wide column of 90000 rows — the 2-D copies back deliver the oracle's bytes.
    (Packing such rows densely on the device for ONE linear D2H was measured on B200: no gain, the D2H engine takes pitched
    rows at full rate — unlike H2D of short rows — and same-size requests are bound by the duplex link, scratch/dense_ab.py.)"""
    import torch
    kw = dict(allow_experiments=True, max_filters=8, max_w=0, max_h=0)
    cases = [((400, 1001, 3), dict(filters=["flip=10"])), ((611, 333, 3), dict(filters=["rotate=90", "gamma=1.3"])),
             ((900, 1203, 3), dict(resize="601,450")), ((300, 500, 4), dict(crop="333px,290px,3px,3px")),
             ((120, 77, 3), dict(filters=["flip=01"])), ((1000, 999, 3), dict(filters=["blur=1.1"]))]
    imgs, plans, outs, keep, refs = [], [], [], [], []
    for k, (shape, rq) in enumerate(cases * 2):
        im = smooth_image(50 + k, *shape)
        p = gpu.plan(shape[1], shape[0], shape[2], api.Config(**kw), **rq)
        row = p.out_w * p.out_c
        if k < len(cases):                                        # pageable destination (an IplImage: rows padded to 4 bytes)
            o = np.zeros((p.out_h, (row + 3) & ~3), np.uint8)
        else:                                                     # pinned: dense for even k, strided (2-D copy) for odd k
            t = torch.zeros((p.out_h, row + (0 if k % 2 == 0 else 5)), dtype=torch.uint8).pin_memory()
            keep.append(t); o = t.numpy()
        imgs.append(im); plans.append(p); outs.append(o); refs.append(_oracle(orc, im, rq, kw)[2])
    tall = np.ascontiguousarray(smooth_image(9, 90000, 8, 3)[:, :1])          # 1 x 90000 column: 3-byte rows, 270 KB
    tp = gpu.plan(1, 90000, 3, api.Config(**kw), filters=["flip=01"])
    tout = np.zeros((90000, 4), np.uint8)
    api.run_host_batch(gpu, plans + [tp], imgs + [tall], outs + [tout], n_streams=3)
    for p, o, r, (shape, rq) in zip(plans, outs, refs, cases * 2):
        _same(o[:, :p.out_w * p.out_c].reshape(p.out_h, p.out_w, p.out_c), r, rq)
    _same(tout[:, :3].reshape(90000, 1, 3), _oracle(orc, tall, dict(filters=["flip=01"]), kw)[2], "column")
    for p in plans + [tp]:
        p.close()


def test_concurrent_host_threads_share_one_device(gpu, orc):
    """Several host threads (ctypes drops the GIL inside the library) drive the same device at once through every host entry
    point — single requests, chunked batches, GIF albums, the operator layer, plan creation with cache hits, misses and
    evictions: every result equals the one the same request gives alone."""
    import threading
    kw = dict(allow_experiments=True, max_filters=8, max_w=0, max_h=0)
    cfg = api.Config(**kw)
    rqs = [dict(resize="100,60"), dict(filters=["blur=1.5", "gamma=1.2"]), dict(crop="150px,100px,10px,5px", filters=["vignette=0.7"]),
           dict(resize="64,64", filters=["blur=6", "kelvin=1"]), dict(filters=["rotate=270", "modulate=0,0,100", "colorize=704214,0.6"]),
           dict(resize="333,200,up"), dict(resize="333,200,up", simple=True), dict(resize="90,70", filters=["gotham=1"])]
    imgs = [smooth_image(70 + i, 120 + 7 * i, 200 + 9 * i, 3 + (i % 2)) for i in range(8)]
    want = {}
    for i, im in enumerate(imgs):
        for j, rq in enumerate(rqs):
            want[(i, j)] = gpu.run(im, cfg, **rq)
    rng = np.random.default_rng(3)
    pages = []
    for f in range(12):
        pages.append(dict(indices=np.ascontiguousarray(rng.integers(0, 16, (40, 64), dtype=np.uint8)), left=0, top=0, dispose=int(rng.integers(0, 4)),
                          key=int(rng.choice([-1, 0, 3])), palette=rng.integers(0, 256, (256, 4), dtype=np.uint8)))
    gplan_rq = dict(resize="32,20", filters=["flip=10"])
    gp = gpu.plan(64, 40, 4, cfg, **gplan_rq)
    gif_want = [o.copy() for o in gpu.gif_album(pages, 64, 40, True, gp)]
    gp.close()
    errors = []

    def worker(t):
        try:
            r = np.random.default_rng(100 + t)
            ops = api.OpsLayer(gpu) if t == 0 else None                 # the operator layer's allocator hooks are process-wide: one user
            for it in range(25):
                i, j = int(r.integers(0, len(imgs))), int(r.integers(0, len(rqs)))
                kind = (it + t) % 4
                if kind == 0:
                    got = gpu.run(imgs[i], cfg, **rqs[j])
                    assert np.array_equal(got, want[(i, j)]), ("run", t, it, i, j)
                elif kind == 1:
                    sel = [(int(r.integers(0, len(imgs))), int(r.integers(0, len(rqs)))) for _ in range(9)]
                    plans = [gpu.plan(imgs[a].shape[1], imgs[a].shape[0], imgs[a].shape[2], cfg, **rqs[b]) for a, b in sel]
                    outs = [np.zeros((p.out_h, p.out_w, p.out_c), np.uint8) for p in plans]
                    api.run_host_batch(gpu, plans, [imgs[a] for a, _ in sel], outs, n_streams=2 + t % 3)
                    for (a, b), o in zip(sel, outs):
                        assert np.array_equal(o, want[(a, b)]), ("batch", t, it, a, b)
                    for p in plans:
                        p.close()
                elif kind == 2:
                    p = gpu.plan(64, 40, 4, cfg, **gplan_rq)
                    outs = gpu.gif_album(pages, 64, 40, True, p)
                    for a, b in zip(outs, gif_want):
                        assert np.array_equal(a, b), ("gif", t, it)
                    p.close()
                elif ops is not None:
                    code, outs = ops.request([imgs[i]], cfg, **{k: v for k, v in rqs[j].items()})
                    assert code == 0 and np.array_equal(outs[0], want[(i, j)]), ("ops", t, it, i, j)
                else:
                    # plans of throw-away geometries: cache misses and evictions while the other threads run
                    for k in range(40):
                        q = gpu.plan(50 + k + 41 * t, 37 + it, 3, cfg, resize=f"{20 + k},{15 + it % 7}")
                        q.close()
        except Exception as e:                                           # noqa: BLE001 — reported by the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in threads: th.start()
    for th in threads: th.join(timeout=300)
    assert not errors, errors[:3]
    assert not any(th.is_alive() for th in threads)


def test_pageable_frames_at_odd_addresses_and_strides(gpu, orc):
    """The copy workers' streaming copies (16-byte stores after an unaligned head, memcpy tail): pageable sources and
    destinations that start at odd addresses, rows of 4 KB and more with odd strides, single requests (row slices) and
    batches, against numpy identities."""
    cfg = api.Config(max_w=0, max_h=0)
    rng = np.random.default_rng(12)
    for (h, w, c, off, pad) in [(700, 1501, 3, 3, 5), (2500, 2300, 4, 1, 7), (3000, 4000, 3, 13, 0), (64, 5000, 3, 7, 11)]:
        step = w * c + pad
        raw = rng.integers(0, 256, off + step * h + 64, dtype=np.uint8)
        src = np.lib.stride_tricks.as_strided(raw[off:], shape=(h, w, c), strides=(step, c, 1))
        for rq, ref in ((dict(filters=["flip=11"]), src[::-1, ::-1]), (dict(crop=f"{w - 7}px,{h - 5}px,4px,3px"), src[3:h - 2, 4:w - 3])):
            p = gpu.plan(w, h, c, cfg, **rq)
            orow = p.out_w * p.out_c
            ostep = orow + pad + 2
            oraw = np.zeros(off + 5 + ostep * p.out_h + 64, np.uint8)
            dst = np.lib.stride_tricks.as_strided(oraw[off + 5:], shape=(p.out_h, p.out_w, p.out_c), strides=(ostep, p.out_c, 1))
            # one request (staged and uploaded in row slices when the window is 16 MB or more) ...
            gpu.check(gpu.lib.imp_gpu_run_host(p.h, src.ctypes.data, step, dst.ctypes.data, ostep))
            assert np.array_equal(dst, ref), (h, w, c, rq)
            # ... and a batch of three into fresh destinations
            oraws = [np.zeros_like(oraw) for _ in range(3)]
            dsts = [np.lib.stride_tricks.as_strided(o[off + 5:], shape=(p.out_h, p.out_w, p.out_c), strides=(ostep, p.out_c, 1)) for o in oraws]
            api.run_host_batch(gpu, [p] * 3, [src] * 3, dsts, n_streams=2)
            for d, o in zip(dsts, oraws):
                assert np.array_equal(d, ref), (h, w, c, rq, "batch")
                gap = np.lib.stride_tricks.as_strided(o[off + 5 + orow:], shape=(p.out_h - 1, ostep - orow), strides=(ostep, 1))
                assert not gap.any() and not o[:off + 5].any()              # nothing written between or before the rows
            p.close()
