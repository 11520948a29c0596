"""
Generates tests/golden/golden_v1.npz from the REFERENCE ITSELF, in the build container:
  * filters / compositing / crop+resize argument maths: the reference's own filters.c, helpers.c,
    bridge.c compiled unmodified (oracle/_ref/libimp_ref.so), driven through RunJob (bridge.c:302-724)
    with a RAW codec, so stage order and error codes are the reference's too;
  * the OpenCV calls inside it (cvResize, cvSmooth, cvFlip, cvTranspose): cv2 4.13.0, IPP off, 1 thread
    (OpenCV 2.4.9 itself is not available offline).
Run:  python tests/golden/make_golden.py        (needs /root/reference and cv2; both exist only here)
The .npz travels with the repo; nothing on the GPU box reads /root/reference.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

KAT = np.array([[[0, 0, 0], [255, 255, 255], [10, 128, 250], [200, 30, 180]],
                [[37, 201, 99], [128, 128, 128], [255, 0, 170], [3, 2, 1]]], np.uint8)   # SURVEY App. B input


def image(seed, h, w, c):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, c), np.float64)
    for k in range(c):
        img[:, :, k] = 127 + 100 * np.sin(xx / (3.0 + k)) * np.cos(yy / (4.0 + k)) + rng.normal(0, 25, (h, w))
    if c == 4:
        img[:, :, 3] = rng.choice([0, 64, 128, 200, 255], size=(h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


WM = image(900, 10, 14, 4)
CFG = {
    "default": dict(),
    "exp": dict(allow_experiments=True, max_filters=8),
    "wm": dict(allow_experiments=True, max_filters=8, watermark="WM", wm_gravity_x="r", wm_gravity_y="b", wm_offset_x=3, wm_offset_y=2, wm_opacity=60),
    "wm_c": dict(watermark="WM", wm_gravity_x="c", wm_gravity_y="c", wm_offset_x=-30, wm_offset_y=4, wm_opacity=100),
}

# (name, image spec, query, config key, extension)
CASES = []
for f in ["flip=10", "flip=01", "flip=11", "rotate=90", "rotate=180", "rotate=270", "modulate=0,0,100", "modulate=60,70,80",
          "modulate=100,500,500", "modulate=180,-50,100", "colorize=704214,0.6", "colorize=ff0000", "gamma=1.3", "gamma=0.5",
          "contrast=1.5", "contrast=0.5", "gradmap=306090,eecc00", "gradmap=000000,ff0000,ffffff", "vignette=0.8", "vignette=4,3",
          "gotham=1", "lomo=1", "kelvin=1", "rainbow=full", "rainbow=mid", "rainbow=pale", "scanline=0", "scanline=0.5,0.25,1,1",
          "scanline=0.3,0.9,2,3", "blur=0.5", "blur=1", "blur=2.3"]:
    CASES.append(("kat/" + f, "KAT", "filter-" + f, "exp", "png"))
    CASES.append(("rgba/" + f, (11, 21, 27, 4), "filter-" + f, "exp", "png"))
CASES += [
    ("cfg1/area3x3", (1, 54, 96, 3), "resize=32,18", "default", "png"),
    ("cfg1/area2x2", (2, 36, 64, 3), "resize=32,18", "default", "png"),
    ("cfg2/crop-area4.5-wm", (3, 108, 192, 4), "crop=180px,99px,c,c&resize=40,22", "wm", "png"),
    ("cfg2/area-frac-3ch", (4, 77, 131, 3), "resize=50,31", "default", "png"),
    ("cfg3/cubic2x-sepia", (5, 27, 48, 4), "resize=96,54,up&filter-modulate=0,0,100&filter-colorize=704214,0.6", "default", "png"),
    ("cfg3/cubic-odd", (6, 23, 31, 3), "resize=47,40,up", "default", "png"),
    ("cfg3/cubic-gray", (7, 19, 26, 1), "resize=41,33,up", "default", "png"),
    ("cfg4/blur-vignette-rot90", (8, 40, 30, 3), "filter-blur=2.3&filter-vignette=0.8&filter-rotate=90", "exp", "png"),
    ("cfg5/thumb-wm", (9, 90, 120, 3), "resize=32,32", "wm", "png"),
    ("crop/ratio", (10, 50, 70, 3), "crop=1,1", "default", "png"),
    ("crop/ratio-gravity", (10, 50, 70, 3), "crop=16,9&gravity=r,b", "default", "png"),
    ("crop/abs", (10, 50, 70, 4), "crop=40px,20px,6px,3px", "default", "png"),
    ("resize/width-only", (12, 50, 70, 3), "resize=35", "default", "png"),
    ("resize/height-only", (12, 50, 70, 3), "resize=0,20", "default", "png"),
    ("flatten/jpg", (13, 20, 30, 4), "resize=15&format=jpg", "default", "png"),
    ("flatten/wm-jpg", (13, 40, 60, 4), "filter-gamma=0.8&format=jpg", "wm_c", "png"),
    ("gray/filters", (14, 20, 30, 1), "filter-colorize=704214,0.6", "default", "png"),
    ("chain/mixed", (15, 40, 52, 4), "crop=3,2&resize=30&filter-rotate=90&filter-vignette=0.7&filter-flip=01&filter-scanline=0.4,0.3,2,1&filter-blur=1.1&filter-kelvin=1", "wm", "png"),
    ("err/bad-crop", (16, 20, 30, 3), "crop=400px,200", "default", "png"),
    ("err/too-big", (16, 20, 30, 3), "resize=3000,0,up", "default", "png"),
    ("err/no-filter", (16, 20, 30, 3), "filter-vignette=0.8", "default", "png"),
    ("err/bad-args", (16, 20, 30, 3), "filter-modulate=181,1,1", "default", "png"),
    ("err/too-many", (16, 20, 30, 3), "filter-gamma=1&filter-gamma=1&filter-gamma=1&filter-gamma=1&filter-gamma=1&filter-gamma=1", "default", "png"),
]


def main():
    assert O.Ref.available(), "needs /root/reference"
    assert O.Ref.use_cv2(True), "needs cv2"
    arrays, meta = {"WM": WM}, []
    for i, (name, spec, query, cfgk, ext) in enumerate(CASES):
        img = KAT if spec == "KAT" else image(*spec)
        kw = {k: (WM if v == "WM" else v) for k, v in CFG[cfgk].items()}
        code, step, out = O.Ref.run_job(query, img, O.OracleConfig(**kw), exten=ext)
        arrays[f"in{i}"] = img
        if out is not None:
            arrays[f"out{i}"] = out
        meta.append(dict(name=name, query=query, cfg=cfgk, code=code, step=step))
    arrays["meta"] = np.frombuffer(json.dumps(dict(cases=meta, cfg=CFG)).encode(), np.uint8)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **arrays)
    print(path, os.path.getsize(path), "bytes,", len(meta), "cases")


if __name__ == "__main__":
    main()
