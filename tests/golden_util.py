"""Loads tests/golden/golden_v1.npz (made by tests/golden/make_golden.py from the compiled reference + cv2)."""
import json
import os
from urllib.parse import unquote

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def load():
    z = np.load(PATH)
    meta = json.loads(bytes(z["meta"]).decode())
    cases = []
    for i, m in enumerate(meta["cases"]):
        kw = {k: (z["WM"] if v == "WM" else v) for k, v in meta["cfg"][m["cfg"]].items()}
        cases.append(dict(m, img=z[f"in{i}"], out=(z[f"out{i}"] if f"out{i}" in z.files else None), cfgkw=kw))
    return cases


def split_query(query):
    """RunJob's query scan (bridge.c:346-372) for the keys the hot path uses. Returns request kwargs + format."""
    req = dict(crop=None, gravity=None, resize=None, filters=[])
    fmt = None
    for tok in [t for t in query.split("&") if t]:
        if tok.startswith("crop"): req["crop"] = tok.split("=", 1)[1]
        elif tok.startswith("gravity"): req["gravity"] = tok.split("=", 1)[1]
        elif tok.startswith("resize"): req["resize"] = tok.split("=", 1)[1]
        elif tok.startswith("format"): fmt = tok.split("=", 1)[1]
        elif tok.startswith("filter"): req["filters"].append(tok.split("-", 1)[1])
    req["flatten"] = (fmt == "jpg")
    return req


# gradmap LUT tails the reference leaves uninitialised: no golden case uses 4/6/7/8 colours.
VIGNETTE_TOL = 1   # <= 1 LSB where libm's and CUDA's double cos may differ


IO_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_io_v1.npz")


def load_io():
    """tests/golden/golden_io_v1.npz (made by tests/golden/make_golden_io.py from the reference's advancedio.c).
    Returns dict(gifs=[dict(cw, ch, frames=[dict(indices top-down, palette, left, top, dispose, key)], out={0: [...], 1: [...]})],
                 packs=[dict(img, fi24, fi32)], jobs=[dict(gif, query, out)])."""
    z = np.load(IO_PATH)
    meta = json.loads(bytes(z["meta"]).decode())
    gifs = []
    for gi, m in enumerate(meta["gifs"]):
        frames = [dict(f, indices=z[f"g{gi}_idx{fi}"], palette=z[f"g{gi}_pal{fi}"]) for fi, f in enumerate(m["frames"])]
        out = {d: [z[f"g{gi}_d{d}_out{fi}"] for fi in range(len(frames))] for d in (0, 1)}
        gifs.append(dict(cw=m["cw"], ch=m["ch"], frames=frames, out=out))
    packs = [dict(img=z[f"p{pi}_in"], fi24=z[f"p{pi}_fi24"], fi32=z[f"p{pi}_fi32"]) for pi in range(len(meta["packs"]))]
    jobs = [dict(j, out=z[f"j{ji}_out"]) for ji, j in enumerate(meta["jobs"])]
    return dict(gifs=gifs, packs=packs, jobs=jobs)


def fi_page(frame):
    """A golden GIF frame as FreeImage holds the locked page (what LoadGIF reads and the C ABI takes): bottom-up scanlines
    padded to 4 bytes with zeros, plus the true width."""
    idx = np.ascontiguousarray(frame["indices"], np.uint8)
    h, w = idx.shape
    bits = np.zeros((h, (w + 3) & ~3), np.uint8)
    bits[:, :w] = idx[::-1]
    return dict(indices=bits, width=w, palette=frame["palette"], left=frame["left"], top=frame["top"], dispose=frame["dispose"], key=frame["key"])
