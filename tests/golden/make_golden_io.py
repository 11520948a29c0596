"""
Generates tests/golden/golden_io_v1.npz from the REFERENCE's advancedio.c (compiled unmodified into
oracle/_ref/libimp_ref.so over the stand-in FreeImage of oracle/fake_freeimage.c), in the build container:
  * gif/<k>:   LoadGIF (advancedio.c:104-262) on a multi-page container: palette-index pages with FrameLeft/Top,
               DisposalMethod and transparent index in, the BGRA canvases of FiLoadFrames out (isdestructive 0 and 1);
  * pack/<k>:  IplToFI32 / IplToFI24 (advancedio.c:65-101) through FiSaveFrames/SaveSingle: frame in, FIBITMAP bits out;
  * job/<k>:   whole RunJob (bridge.c:302-724) on a GIF container with page=N: FiLoadFrames -> steps 3-7 -> encode.
Frames avoid the one pixel whose value the reference leaves to the heap (row[w] of the top scanline when w % 4 == 0).
Run:  python tests/golden/make_golden_io.py      (needs /root/reference and cv2)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def gif_frames(seed, cw, ch, n, keys=(-1, 0, 3, 255)):
    rng = np.random.default_rng(seed)
    frames = []
    for f in range(n):
        if f == 0:
            w, h, left, top = cw, ch, 0, 0
        else:
            while True:
                w = int(rng.integers(1, cw + 1)); h = int(rng.integers(1, ch + 1))
                left = int(rng.integers(0, cw - w + 1)); top = int(rng.integers(0, ch - h + 1))
                if w % 4 or left + w >= cw:
                    break
        dispose = int(rng.integers(0, 4))
        if f == 0 and dispose == 2:
            dispose = 1                                   # frame 0 + BACKGROUND reads the uninitialised master
        frames.append(dict(indices=rng.integers(0, 256 if f % 2 else 6, (h, w), dtype=np.uint8),
                           palette=rng.integers(0, 256, (256, 4), dtype=np.uint8),
                           left=left, top=top, dispose=dispose, key=int(rng.choice(keys)), time=10 * f))
    return frames


GIFS = [(1, 48, 27, 5), (2, 131, 77, 7), (3, 1, 1, 2), (4, 9, 40, 4), (5, 64, 64, 6)]
PACKS = [(1, 5, 7, 3), (2, 5, 7, 4), (3, 33, 18, 3), (4, 32, 20, 4), (5, 1, 1, 3)]
JOBS = [(1, "page=3&resize=60,35&format=png"), (4, "page=2&crop=1,1&filter-gamma=1.3&format=png"),
        (0, "page=4&filter-rotate=90&format=bmp"), (1, "page=1&resize=40&format=jpg"), (0, "page=0&format=tga"),
        (1, "page=2&resize=50,30&format=ppm")]      # (index into GIFS, query)


def main():
    assert O.Ref.available() and O.Ref.use_cv2(True)
    rng = np.random.default_rng(99)
    arrays, meta = {}, dict(gifs=[], packs=[], jobs=[])
    for gi, (seed, cw, ch, n) in enumerate(GIFS):
        frames = gif_frames(seed, cw, ch, n)
        blob = O.Ref.gif_container(frames)
        m = dict(cw=cw, ch=ch, frames=[{k: v for k, v in f.items() if k not in ("indices", "palette")} for f in frames])
        for fi, f in enumerate(frames):
            arrays[f"g{gi}_idx{fi}"] = f["indices"]; arrays[f"g{gi}_pal{fi}"] = f["palette"]
        for d in (0, 1):
            err, out = O.Ref.fi_load(blob, O.Ref.FIF_GIF, bool(d))
            assert err == 0 and len(out) == n
            for fi, o in enumerate(out):
                arrays[f"g{gi}_d{d}_out{fi}"] = o["image"]
                assert (o["time"], o["dispose"], o["key"]) == (frames[fi]["time"], frames[fi]["dispose"], frames[fi]["key"])
        meta["gifs"].append(m)
    for pi, (seed, h, w, c) in enumerate(PACKS):
        img = np.random.default_rng(seed).integers(0, 256, (h, w, c), dtype=np.uint8)
        arrays[f"p{pi}_in"] = img
        for fmt, bits in ((O.Ref.FIF_BMP, 32), (O.Ref.FIF_JPEG, 24)):
            err, bpp, out = O.Ref.fi_save(img, fmt)
            assert err == 0 and bpp == bits
            arrays[f"p{pi}_fi{bits}"] = out
        meta["packs"].append(dict(h=h, w=w, c=c))
    for ji, (gi, query) in enumerate(JOBS):
        seed, cw, ch, n = GIFS[gi]
        blob = O.Ref.gif_container(gif_frames(seed, cw, ch, n))
        code, step, out = O.Ref.run_job_blob(query, blob)
        assert code == 0 and out is not None, (query, code, step)
        arrays[f"j{ji}_out"] = out
        meta["jobs"].append(dict(gif=gi, query=query, code=code, step=step))
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_io_v1.npz")
    np.savez_compressed(path, **arrays)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
