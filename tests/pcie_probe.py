#!/usr/bin/env python
"""Developer probe (test infrastructure): host<->device copy rates of default-pinned vs write-combined pinned memory, one way and
both ways at once -- what bounds the `e2e` number of bench.py.
    python tests/pcie_probe.py
"""
import time

import torch
from cuda.bindings import runtime as rt

N = 1 << 30
torch.cuda.init()
dev_in = torch.empty(N, dtype=torch.uint8, device="cuda")
dev_out = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost


def rate(fn, nbytes, reps=4):
    best = 0.0
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        best = max(best, reps * nbytes / (time.perf_counter() - t0) / 1e9)
    return best


for name, flags in (("default pinned", rt.cudaHostAllocDefault), ("write-combined pinned", rt.cudaHostAllocWriteCombined)):
    err, src = rt.cudaHostAlloc(N, flags)
    err2, dst = rt.cudaHostAlloc(N, rt.cudaHostAllocDefault)
    assert err == rt.cudaError_t.cudaSuccess and err2 == rt.cudaError_t.cudaSuccess, (err, err2)
    h2d = rate(lambda: rt.cudaMemcpyAsync(dev_in.data_ptr(), src, N, H2D, s1.cuda_stream), N)
    d2h = rate(lambda: rt.cudaMemcpyAsync(dst, dev_out.data_ptr(), N, D2H, s2.cuda_stream), N)

    def both():
        rt.cudaMemcpyAsync(dev_in.data_ptr(), src, N, H2D, s1.cuda_stream)
        rt.cudaMemcpyAsync(dst, dev_out.data_ptr(), (N * 6) // 10, D2H, s2.cuda_stream)

    duplex = rate(both, N)
    print(f"{name:24s} H2D alone {h2d:5.1f} GB/s   D2H alone {d2h:5.1f} GB/s   H2D with a 0.6x D2H at the same time {duplex:5.1f} GB/s")
    rt.cudaFreeHost(src)
    rt.cudaFreeHost(dst)
