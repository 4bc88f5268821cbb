#!/usr/bin/env python
"""Host<->device copy rates the e2e design depends on: pinned H2D, 2-D (page de-framing) H2D,
and the cost of cudaHostRegister on a large pageable buffer."""
import ctypes as C
import time

import numpy as np
import torch

rt = C.CDLL("libcudart.so.12")
torch.cuda.init()
n = 4 << 30
host = np.empty(n, np.uint8)
host[::4096] = 1
t0 = time.time()
rc = rt.cudaHostRegister(C.c_void_p(host.ctypes.data), C.c_size_t(n), C.c_uint(0))
t1 = time.time()
print(f"cudaHostRegister rc={rc}: {n / 2**30:.0f} GiB in {t1 - t0:.2f}s = {n / 1e9 / (t1 - t0):.1f} GB/s")
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
for name, fn in (("1-D", lambda: rt.cudaMemcpyAsync(dev.data_ptr(), host.ctypes.data, n, 1, s)),
                 ("2-D 8168/8192", lambda: rt.cudaMemcpy2DAsync(dev.data_ptr(), 8168, host.ctypes.data + 24, 8192, 8168, n // 8192 - 1, 1, s)),
                 ("2-D x 3400 calls of 150 pages", None)):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        if fn is not None:
            fn()
        else:
            pages = 150
            for i in range(3400):
                rt.cudaMemcpy2DAsync(dev.data_ptr() + i * pages * 8168, 8168, host.ctypes.data + i * pages * 8192 + 24, 8192, 8168, pages, 1, s)
        t_issue = time.time() - t0
        torch.cuda.synchronize()
        dt = time.time() - t0
        moved = n if fn is not None else 3400 * 150 * 8168
    print(f"{name}: {moved / 1e9 / dt:.1f} GB/s (issue {t_issue * 1e3:.1f} ms, total {dt * 1e3:.1f} ms)")
