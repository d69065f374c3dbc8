"""`timing` context manager and `Timer` (stand-ins for `ifs_physics_common.timing`, used at
drivers/run_nonlinear.py:116-119 and tangent_linear/validation.py:151,167,178).  A label's time
is measured on the device with CUDA events on the current stream when a GPU is present
(wall-clock otherwise) and accumulated across uses of the same label."""
from __future__ import annotations

import time
from contextlib import contextmanager
from typing import Dict, List

import torch


class Timer:
    _pending: Dict[str, List] = {}
    _host: Dict[str, float] = {}

    @classmethod
    def reset(cls) -> None:
        cls._pending = {}
        cls._host = {}

    @classmethod
    def get_time(cls, label: str, units: str = "ms") -> float:
        total_ms = cls._host.get(label, 0.0) * 1e3
        if cls._pending.get(label):
            torch.cuda.synchronize()
            total_ms += sum(a.elapsed_time(b) for a, b in cls._pending[label])
        return {"ms": total_ms, "s": total_ms * 1e-3, "us": total_ms * 1e3}[units]


@contextmanager
def timing(label: str):
    if torch.cuda.is_available():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield Timer
        b.record()
        Timer._pending.setdefault(label, []).append((a, b))
    else:
        t0 = time.perf_counter()
        yield Timer
        Timer._host[label] = Timer._host.get(label, 0.0) + time.perf_counter() - t0
