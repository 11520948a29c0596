import sys, time, ctypes as C
sys.path.insert(0, ".")
import torch, numpy as np
import ngx_http_imgproc_b200 as M
L = M.library(); L.init(0)
lib = L.lib
lib.imp_gpu_upload_2d.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
lib.imp_gpu_download_2d.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
lib.imp_gpu_sync.argtypes = [C.c_void_p]
for (rows, wb) in [(888, 3552), (888, 3552 * 2), (2025, 14400), (256, 768), (1080, 5760)]:
    h = torch.randint(0, 256, (rows, wb), dtype=torch.uint8).pin_memory()
    for dp in (wb, (wb + 15) & ~15 if wb % 16 else wb + 16):
        d = torch.empty(rows * dp + 64, dtype=torch.uint8, device="cuda")
        for name, fn, a in (("H2D", lib.imp_gpu_upload_2d, (d.data_ptr(), dp, h.data_ptr(), wb)), ("D2H", lib.imp_gpu_download_2d, (h.data_ptr(), wb, d.data_ptr(), dp))):
            for _ in range(5): fn(*a, wb, rows, None)
            lib.imp_gpu_sync(None)
            t0 = time.perf_counter()
            n = 200
            for _ in range(n): fn(*a, wb, rows, None)
            lib.imp_gpu_sync(None)
            dt = (time.perf_counter() - t0) / n
            print(f"{name} rows={rows} width={wb} device pitch={dp}: {dt*1e6:7.1f} us  {rows*wb/dt/1e9:6.1f} GB/s")
