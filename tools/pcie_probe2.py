#!/usr/bin/env python
"""H2D copy rate as a function of copy size and source kind (cudaHostRegister'ed malloc memory vs
cudaHostAlloc), the question behind the e2e staging design."""
import ctypes as C
import time

import numpy as np
import torch

rt = C.CDLL("libcudart.so.12")
torch.cuda.init()
n = 2 << 30
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
s2 = torch.cuda.Stream().cuda_stream

host = np.empty(n, np.uint8)
host[::4096] = 1
assert rt.cudaHostRegister(C.c_void_p(host.ctypes.data), C.c_size_t(n), C.c_uint(0)) == 0
p = C.c_void_p()
assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(0)) == 0
C.memset(p, 1, n)

for name, base in (("registered", host.ctypes.data), ("cudaHostAlloc", p.value)):
    for sz in (1 << 31, 64 << 20, 3686400, 1228800, 262144):
        cnt = n // sz
        for two in (False, True):
            best = 0
            for rep in range(3):
                torch.cuda.synchronize()
                t0 = time.time()
                for i in range(cnt):
                    rt.cudaMemcpyAsync(dev.data_ptr() + i * sz, base + i * sz, sz, 1, s2 if (two and i & 1) else s)
                ti = time.time() - t0
                torch.cuda.synchronize()
                dt = time.time() - t0
                best = max(best, cnt * sz / 1e9 / dt)
            print(f"{name:14s} size {sz:>11d} x {cnt:>5d} {'2 streams' if two else '1 stream '}: {best:5.1f} GB/s (issue {ti * 1e3:.1f} ms)")
