import numpy as np, sys
sys.path.insert(0, '.')
import ngx_http_imgproc_b200 as M
from oracle import oracle as O
L = M.library(); L.init(0)
def rnd(seed,h,w,c): return np.random.default_rng(seed).integers(0,256,(h,w,c),dtype=np.uint8)
for shape in [(256,256,3),(256,256,4),(131,67,3),(64,256,3),(256,64,3),(128,128,3),(512,512,3)]:
    for f in ["rotate=90","rotate=270","flip=10"]:
        img = rnd(1,*shape)
        kw = dict(allow_experiments=True)
        for rep in range(3):
            out = L.run(img, M.Config(**kw), filters=[f])
            code, step, ref = O.run_chain(img, None, None, None, [f], O.OracleConfig(**kw), False, False)
            d = (out != ref)
            if d.any():
                ys, xs, cs = np.nonzero(d)
                print(shape, f, rep, "bad bytes", d.sum(), "rows", ys.min(), ys.max(), "cols", xs.min(), xs.max(), "uniq cols", len(np.unique(xs)), "uniq rows", len(np.unique(ys)))
                # is the wrong data from another location?
                y, x = ys[0], xs[0]
                print("  first bad at", y, x, "got", out[y,x], "want", ref[y,x], " cols hist", np.bincount(xs//8)[:40])
            else:
                print(shape, f, rep, "ok")
