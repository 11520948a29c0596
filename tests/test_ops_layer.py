"""The reference-signature operator layer (include/imp_ops.h): Crop/Resize/Filter/Watermark/BlendWithPaper
record + validate (CPU-testable: codes and header fix-ups), imp_Flush runs the fused plan (GPU)."""
import ctypes as C

import numpy as np
import pytest

import ngx_http_imgproc_b200 as M
from ngx_http_imgproc_b200 import api
from conftest import rnd_image, smooth_image
from oracle.oracle import IplImage          # layout-identical ctypes mirror of OpenCV's IplImage (test infrastructure)


def _lib():
    L = M.library().lib
    L.imp_Crop.argtypes = [C.POINTER(C.POINTER(IplImage)), C.c_char_p, C.c_char_p]
    L.imp_Resize.argtypes = [C.POINTER(C.POINTER(IplImage)), C.c_char_p, C.POINTER(api.CConfig), C.c_int]
    L.imp_Filter.argtypes = [C.POINTER(C.POINTER(IplImage)), C.c_char_p, C.c_int]
    L.imp_Watermark.argtypes = [C.POINTER(IplImage), C.POINTER(api.CConfig)]
    L.imp_BlendWithPaper.argtypes = [C.POINTER(IplImage)]
    L.imp_Flush.argtypes = [C.POINTER(C.POINTER(IplImage))]
    L.imp_FlushAll.argtypes = [C.POINTER(C.POINTER(IplImage)), C.c_int]
    L.imp_Discard.argtypes = [C.POINTER(IplImage)]
    L.imp_ops_pending.argtypes = [C.POINTER(IplImage)]
    L.imp_ops_set_image_allocator.argtypes = [_CREATE_T, _RELEASE_T]
    L.imp_ops_set_image_allocator(*_CB)
    return L


# Frames in these tests live in numpy buffers, so the layer gets an allocator that keeps them Python-owned
# (in nginx the two callbacks are thin wrappers of cvCreateImage / cvReleaseImage).
_KEEP = []
_CREATE_T = C.CFUNCTYPE(C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int)
_RELEASE_T = C.CFUNCTYPE(None, C.POINTER(C.c_void_p))


def _create(w, h, depth, c):
    im, buf = _ipl(np.zeros((h, w, c), np.uint8))
    _KEEP.append((im, buf))
    return C.addressof(im)


def _release(pp):
    pp[0] = None


_CB = (_CREATE_T(_create), _RELEASE_T(_release))


def _ipl(img):
    """IplImage header over a numpy buffer with cvCreateImage's 4-byte row alignment."""
    h, w, c = img.shape
    step = (w * c + 3) & ~3
    buf = np.zeros((h, step), np.uint8)
    buf[:, :w * c] = img.reshape(h, w * c)
    im = IplImage()
    im.nSize = C.sizeof(IplImage); im.nChannels = c; im.depth = 8; im.width = w; im.height = h
    im.widthStep = step; im.imageSize = step * h
    im.imageData = buf.ctypes.data; im.imageDataOrigin = buf.ctypes.data
    return im, buf


def _to_np(p):
    im = p.contents
    raw = (C.c_ubyte * (im.widthStep * im.height)).from_address(im.imageData)
    a = np.frombuffer(raw, np.uint8).reshape(im.height, im.widthStep)[:, :im.width * im.nChannels]
    return a.reshape(im.height, im.width, im.nChannels).copy()


def test_operators_validate_like_the_reference_and_fix_up_the_header():
    L = _lib()
    im, keep = _ipl(rnd_image(1, 60, 80, 3))
    p = C.pointer(im)
    assert L.imp_Crop(C.byref(p), b"400px,200", None) == 50 and (im.width, im.height) == (80, 60)
    assert L.imp_Crop(C.byref(p), b"40px,30px,c,c", None) == 0 and (im.width, im.height) == (40, 30)
    cfg, k = api.Config(max_w=100, max_h=100).to_c()
    assert L.imp_Resize(C.byref(p), b"3000,0,up", C.byref(cfg), 0) == 54 and (im.width, im.height) == (40, 30)
    assert L.imp_Resize(C.byref(p), b"20", C.byref(cfg), 0) == 0 and (im.width, im.height) == (20, 15)
    assert L.imp_Filter(C.byref(p), b"vignette=0.8", 0) == 52
    assert L.imp_Filter(C.byref(p), b"modulate=181,1,1", 0) == 50
    assert L.imp_Filter(C.byref(p), b"rotate=90", 0) == 0 and (im.width, im.height) == (15, 20)
    assert L.imp_ops_pending(p) == 3
    L.imp_Discard(p)
    assert L.imp_ops_pending(p) == 0 and (im.width, im.height, im.nChannels) == (80, 60, 3)
    g, keep2 = _ipl(rnd_image(2, 10, 12, 1))
    pg = C.pointer(g)
    assert L.imp_Filter(C.byref(pg), b"gamma=1.2", 0) == 0 and g.nChannels == 3     # gray -> BGR (bridge.c:613-618)
    L.imp_Discard(pg)


@pytest.mark.gpu
def test_recorded_chain_runs_fused_and_matches_the_oracle(gpu, orc):
    L = _lib()
    wm = rnd_image(3, 12, 20, 4)
    kw = dict(allow_experiments=True, max_filters=8, watermark=wm, wm_gravity_x="r", wm_gravity_y="b", wm_offset_x=3, wm_offset_y=2, wm_opacity=60)
    cfg, k = api.Config(**kw).to_c()
    for c in (3, 4, 1):
        img = smooth_image(10 + c, 90, 120, c)
        im, keep = _ipl(img)
        p = C.pointer(im)
        assert L.imp_Crop(C.byref(p), b"100px,80px,c,c", None) == 0
        assert L.imp_Resize(C.byref(p), b"45,37", C.byref(cfg), 0) == 0
        for f in (b"modulate=0,0,100", b"colorize=704214,0.6", b"rotate=90", b"scanline=0.4,0.3,2,1"):
            assert L.imp_Filter(C.byref(p), f, 1) == 0
        assert L.imp_Watermark(p, C.byref(cfg)) == 0
        if c == 4:
            assert L.imp_BlendWithPaper(p) == 0
        before = gpu.launch_count()
        assert L.imp_Flush(C.byref(p)) == 0
        assert gpu.launch_count() - before == 1                      # the whole chain is ONE kernel
        out = _to_np(p)
        code, step, ref = orc.run_chain(img, "100px,80px,c,c", None, "45,37", ["modulate=0,0,100", "colorize=704214,0.6", "rotate=90", "scanline=0.4,0.3,2,1"],
                                        orc.OracleConfig(**kw), False, flatten=(c == 4))
        assert code == 0 and out.shape == ref.shape and np.array_equal(out, ref)
        assert L.imp_ops_pending(p) == 0


@pytest.mark.gpu
def test_flush_all_frames_of_an_album(gpu, orc):
    L = _lib()
    frames = [smooth_image(40 + i, 54, 96, 4) for i in range(6)]
    ipls = [_ipl(f) for f in frames]
    arr = (C.POINTER(IplImage) * len(frames))(*[C.pointer(i[0]) for i in ipls])
    cfg, k = api.Config(max_w=0, max_h=0).to_c()
    for i in range(len(frames)):
        pp = C.cast(C.byref(arr, i * C.sizeof(C.c_void_p)), C.POINTER(C.POINTER(IplImage)))
        assert L.imp_Resize(pp, b"192,108,up", C.byref(cfg), 0) == 0
        assert L.imp_Filter(pp, b"gamma=1.3", 0) == 0
    assert L.imp_FlushAll(arr, len(frames)) == 0
    for i, f in enumerate(frames):
        code, step, ref = orc.run_chain(f, None, None, "192,108,up", ["gamma=1.3"], orc.OracleConfig(max_w=0, max_h=0))
        assert np.array_equal(_to_np(arr[i]), ref)


@pytest.mark.gpu
def test_gray_frame_with_no_operator_still_leaves_as_bgr(gpu, orc):
    """bridge.c:613-618 converts every 1-channel frame at the filter step even when no operator was requested; the flush
    does the same (found by running the reference's RunJob over the GPU path, test_reference_runjob_with_the_gpu_path_dropped_in)."""
    L = _lib()
    img = rnd_image(9, 21, 34, 1)
    im, keep = _ipl(img)
    p = C.pointer(im)
    assert L.imp_ops_pending(p) == 0
    assert L.imp_Flush(C.byref(p)) == 0
    out = _to_np(p)
    assert out.shape == (21, 34, 3) and np.array_equal(out, np.repeat(img, 3, axis=2))
    img3 = rnd_image(10, 21, 34, 3)
    im3, keep3 = _ipl(img3)
    p3 = C.pointer(im3)
    before = gpu.launch_count()
    assert L.imp_Flush(C.byref(p3)) == 0 and gpu.launch_count() == before          # nothing recorded, nothing launched
    assert np.array_equal(_to_np(p3), img3)
