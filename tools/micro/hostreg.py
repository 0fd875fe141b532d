"""Micro-benchmark: cost of cudaHostRegister / cudaHostUnregister on pageable numpy arrays vs a plain memcpy."""
import ctypes, time
import numpy as np
import torch
rt = ctypes.CDLL("libcudart.so.12") if True else None
torch.cuda.init(); torch.zeros(1, device="cuda")
for mb in (8, 40, 80):
    a = np.random.rand(mb * 1024 * 1024 // 8)
    b = np.empty_like(a)
    p = ctypes.c_void_p(a.ctypes.data)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); rc = rt.cudaHostRegister(p, ctypes.c_size_t(a.nbytes), 0); t1 = time.perf_counter()
        rc2 = rt.cudaHostUnregister(p); t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1))
    t0 = time.perf_counter(); b[:] = a; tm = time.perf_counter() - t0
    d = torch.empty(a.size, dtype=torch.float64, device="cuda")
    t0 = time.perf_counter(); d.copy_(torch.from_numpy(a)); torch.cuda.synchronize(); th = time.perf_counter() - t0
    print("%3d MB: register %.2f ms  unregister %.2f ms (rc %d %d) | numpy copy %.2f ms | pageable H2D %.2f ms"
          % (mb, 1e3 * min(t[0] for t in ts), 1e3 * min(t[1] for t in ts), rc, rc2, 1e3 * tm, 1e3 * th))
