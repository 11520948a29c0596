import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def rnd_image(seed, h, w, c):
    return np.random.default_rng(seed).integers(0, 256, (h, w, c), dtype=np.uint8)


def smooth_image(seed, h, w, c):
    """gradient + noise, alpha in {0,255}-ish blocks: a less pathological input than white noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, c), np.float64)
    for k in range(c):
        img[:, :, k] = 127 + 100 * np.sin(xx / (7.0 + 3 * k)) * np.cos(yy / (11.0 + 2 * k)) + rng.normal(0, 12, (h, w))
    if c == 4:
        img[:, :, 3] = np.where(((xx // 8) + (yy // 8)) % 3 == 0, 0, 255)
    return np.clip(img, 0, 255).astype(np.uint8)


# ---- hostsim: the planner + the kernels' per-pixel headers compiled for the CPU (tests only) -------------
class HostSim:
    def __init__(self):
        from ngx_http_imgproc_b200 import api
        self.api = api
        d = os.path.join(ROOT, "tests", "hostsim")
        so = os.path.join(d, "libimp_hostsim.so")
        csrc = os.path.join(ROOT, "ngx_http_imgproc_b200", "csrc")
        deps = [os.path.join(d, "hostsim.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
        if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
            cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
            subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I" + cuda_inc,
                            "-o", so, os.path.join(d, "hostsim.cpp"), os.path.join(csrc, "imp_planner.cpp"), "-lm"], check=True)
        self.lib = C.CDLL(so)
        self.lib.hostsim_plan_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        self.lib.hostsim_plan_destroy.argtypes = [C.c_void_p]
        self.lib.hostsim_plan_output.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
        self.lib.hostsim_plan_passes.argtypes = [C.c_void_p]
        self.lib.hostsim_plan_pass_kind.argtypes = [C.c_void_p, C.c_int]
        self.lib.hostsim_plan_bytes.argtypes = [C.c_void_p]
        self.lib.hostsim_plan_bytes.restype = C.c_ulonglong
        self.lib.hostsim_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]

    def run(self, img, cfg=None, want_info=False, **req):
        """Returns (code, step, image-or-None[, info])."""
        api = self.api
        img = np.ascontiguousarray(img if img.ndim == 3 else img[:, :, None], dtype=np.uint8)
        cfg = cfg or api.Config()
        ccfg, k1 = cfg.to_c()
        creq, k2 = api.make_request(**req)
        plan, step = C.c_void_p(), C.c_int(-1)
        rc = self.lib.hostsim_plan_create(C.byref(creq), C.byref(ccfg), img.shape[1], img.shape[0], img.shape[2], C.byref(plan), C.byref(step))
        if rc:
            return (rc, step.value, None, None) if want_info else (rc, step.value, None)
        w, h, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.hostsim_plan_output(plan, C.byref(w), C.byref(h), C.byref(c))
        out = np.zeros((h.value, w.value, c.value), np.uint8)
        assert self.lib.hostsim_run(plan, img.ctypes.data, img.strides[0], out.ctypes.data, out.strides[0]) == 0
        n = self.lib.hostsim_plan_passes(plan)
        info = dict(passes=n, kinds=[self.lib.hostsim_plan_pass_kind(plan, k) for k in range(n)], bytes=int(self.lib.hostsim_plan_bytes(plan)))
        self.lib.hostsim_plan_destroy(plan)
        return (0, step.value, out, info) if want_info else (0, step.value, out)


@pytest.fixture(scope="session")
def hostsim():
    return HostSim()


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def gpu():
    """The product library on cuda:0. Fails loudly (never falls back) when it cannot run."""
    import ngx_http_imgproc_b200 as m
    L = m.library()
    L.init(int(os.environ.get("LOCAL_RANK", "0")))
    return L
