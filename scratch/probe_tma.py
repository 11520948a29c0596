import subprocess, sys, json
cases = [(131,77,3,50,31),(131,77,4,50,31),(128,77,3,50,31),(100,77,4,38,31),(192,108,3,73,41),(400,300,3,150,113),(131,77,1,50,31),(131,77,3,50,77),(131,77,3,131,31),(262,154,3,50,31)]
code = '''
import sys, numpy as np
sys.path.insert(0, ".")
import ngx_http_imgproc_b200 as M
from oracle import oracle as O
w,h,c,dw,dh = map(int, sys.argv[1:6])
L = M.library(); L.init(0)
img = np.random.default_rng(1).integers(0,256,(h,w,c),dtype=np.uint8)
out = L.run(img, M.Config(max_w=0,max_h=0), resize=f"{dw},{dh}")
ref = O.run_chain(img, resize=f"{dw},{dh}", cfg=O.OracleConfig(max_w=0,max_h=0))[2]
print("OK" if np.array_equal(out, ref) else "MISMATCH %d" % np.abs(out.astype(int)-ref.astype(int)).max())
'''
for cs in cases:
    r = subprocess.run([sys.executable, "-c", code] + [str(x) for x in cs], capture_output=True, text=True)
    last = (r.stdout.strip().splitlines() or [""])[-1]
    err = [l for l in r.stderr.splitlines() if "ImpError" in l or "rror" in l][-1:] 
    print(cs, last, err[:1])
