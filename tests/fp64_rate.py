#!/usr/bin/env python
"""Developer micro-benchmark: FP64 pipe rate with uniform operands vs register operands (see cs2_dfma_rate_regs)."""
import os, sys
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gt4py-dwarf-p-cloudsc2-tl-ad_b200")]
import torch
from cloudsc2_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
for warps_per_sched in (1, 2, 4, 8, 16):
    blocks = 148 * warps_per_sched // 2 + 1 if warps_per_sched < 2 else 148 * (warps_per_sched * 4 // 8) + 1
    blocks = max(blocks, 2)
    iters = 20000
    scratch = torch.rand(256 + blocks * 256, dtype=torch.float64, device=dev) + 0.5
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, fn, flops in (
        ("uniform", lambda it: lib.cs2_dfma_rate(scratch.data_ptr(), blocks - 1, it, st), 2),
        ("regs", lambda it: lib.cs2_dfma_rate_regs(scratch.data_ptr(), blocks, it, 1, st), 2),
        ("mul+fma", lambda it: lib.cs2_dfma_rate_regs(scratch.data_ptr(), blocks, it, 2, st), 3),
    ):
        _lib.check(fn(200), name)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _lib.check(fn(iters), name); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        ninstr = (blocks - 1) * 8 * 8 * iters * (1 if flops == 2 else 2)  # warp instructions
        out[name] = (round((blocks - 1) * 256 * 8 * iters * flops / ms / 1e9, 2), round(ninstr / ms / 1e6 / 592 / 1.965e3 * 1e3, 3))
    print(f"blocks {blocks - 1} x 256 thr (8 chains/thread): TFLOP/s, FP64 warp-instr per scheduler-cycle @1965MHz:", out, flush=True)
