#!/usr/bin/env python
"""Benchmark of the CLOUDSC2 hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--columns C] [--precision double|single]
    python bench.py --impl reference ...      # CPU arm: the NumPy oracle on all host cores

One "step" = the reference driver's timed region (drivers/run_nonlinear.py:114-119): the saturation
stencil followed by the CLOUDSC2-NL stencil over one batch of columns, through the component API.
Workload at N=1 = BASELINE.json configs[1]: 65 536 columns x 137 levels, fp64, synthetic inputs
tiled from the seeded 100-column block (data/input.h5 is not shipped).  With N>1 every rank owns its
own contiguous block of 65 536 columns (weak scaling; columns are independent, no data-path
collective); the time is the MAX over ranks of the CUDA-event time of the K steps.

The JSON line also carries: `roofline` of the dominant kernel (cloudsc2_nl, algorithmic bytes per
column from SURVEY.md section 8d / DESIGN.md), kernel-only numbers for TL and AD (`variants`), `e2e`
(host buffers in, host buffers out, through the same component calls), `cpu_baseline`, `clocks`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from datetime import timedelta

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "gt4py-dwarf-p-cloudsc2-tl-ad_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

NLEV = 137
# algorithmic (compulsory) elements per column: every field touched once
ELEMS = {
    "saturation": 3 * NLEV,                                   # ap, t -> qsat
    "nl": 15 * NLEV + (NLEV + 1) + 6 * NLEV + 4 * (NLEV + 1),  # 3567
    "tl": 2 * (15 * NLEV + (NLEV + 1) + 6 * NLEV + 4 * (NLEV + 1)),  # 7134
    "ad": 2 * (15 * NLEV + (NLEV + 1) + 6 * NLEV + 4 * (NLEV + 1)) + (6 * NLEV + 4 * (NLEV + 1)),  # 8508 incl. seed zeroing
}
METRIC = "columns/s for NL/TL/AD fp64 at 137 levels, % of HBM roofline, 1/2/4/8 B200"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(kernel: str):
    """Per-launch DRAM bytes of a kernel from the committed ncu summary, or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(kernel)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int, enabled: bool = True):
        self.index, self.proc, self.path, self.enabled = device_index, None, None, enabled

    def count(self) -> int:
        """Samples written so far (nvidia-smi needs 1-2 s to start on an 8-GPU box)."""
        if self.proc is None or not self.path:
            return 1 << 30  # nothing to wait for
        try:
            self.fh.flush()
            with open(self.path) as fh:
                return sum(1 for _ in fh)
        except OSError:
            return 1 << 30

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.fh.close()
        return False

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        with open(self.path) as fh:
            for line in fh:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------
# CPU arms
# ------------------------------------------------------------------------------------------
def _oracle_chunk(args):
    ncol, seed, precision = args
    import helpers as H  # noqa: WPS433

    dtype = np.float64 if precision == "double" else np.float32
    P = H.externals()
    st = H.make_state("base", dtype, ncol, seed=0)
    st["f_eta"] = H.onp.eta_levels(st["f_ap"], st["f_aph"])
    t0 = time.perf_counter()
    st["f_qsat"] = H.onp.saturation(st["f_ap"], st["f_t"], P)
    H.onp.cloudsc2_nl(st, H.DT, P)
    return time.perf_counter() - t0


def cpu_oracle_rate(cols_per_worker: int, workers: int, steps: int, warmup: int, precision: str):
    """sat + NL of the NumPy oracle (execution model of the reference's default `numpy` backend),
    one process per core, each over its own chunk of columns.  Returns (columns/s, ms per step)."""
    from concurrent.futures import ProcessPoolExecutor

    with ProcessPoolExecutor(max_workers=workers) as pool:
        jobs = [(cols_per_worker, w, precision) for w in range(workers)]
        for _ in range(max(warmup, 1)):
            list(pool.map(_oracle_chunk, jobs))
        t0 = time.perf_counter()
        for _ in range(steps):
            list(pool.map(_oracle_chunk, jobs))
        elapsed = time.perf_counter() - t0
    return cols_per_worker * workers * steps / elapsed, elapsed / steps * 1e3


def cpu_twin_rate(ncol: int, runs: int, precision: str):
    """sat + NL of the C++/OpenMP host twin of the column code on all cores (columns/s)."""
    import helpers as H

    dtype = np.float64 if precision == "double" else np.float32
    P = H.externals()
    st = H.make_state("base", dtype, ncol)
    st["f_eta"] = H.onp.eta_levels(st["f_ap"], st["f_aph"])
    st["f_qsat"] = H.twin_saturation(st["f_ap"], st["f_t"], P)
    H.twin_nl(st, H.DT, P)
    # time the compute calls only (HostFields packing excluded): call the twin on prepared buffers
    import ctypes as C

    from cloudsc2_b200 import _lib

    h = H.HostFields(ncol, NLEV, dtype)
    f = H._nl_struct(h, st)
    tab = H.level_tables(P, st["f_eta"], NLEV, dtype)
    params, dims = _lib.make_params(P), h.dims()
    lib = H.twin()
    best = float("inf")
    for _ in range(runs):
        t0 = time.perf_counter()
        lib.twin_saturation(C.byref(dims), C.byref(params), C.c_void_p(f.in_ap), C.c_void_p(f.in_t), C.c_void_p(f.in_qsat))
        lib.twin_nl(C.byref(dims), C.byref(params), C.c_double(H.DT), C.c_void_p(tab.ctypes.data), C.byref(f))
        best = min(best, time.perf_counter() - t0)
    return ncol / best


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    cols = 1024
    rate, ms = cpu_oracle_rate(cols, workers, args.steps, args.warmup, args.precision)
    sample = (f"saturation + cloudsc2_nl of the NumPy oracle on {workers} processes x {cols} synthetic columns x {NLEV} "
              f"levels per step ({args.precision})")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "columns/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if args.precision == "double" else "f32", "data": "synthetic",
        "config": {"workload": f"CLOUDSC2-NL {args.precision}: saturation + cloudsc2_nl, bounded CPU sample "
                               f"{workers * cols} columns x {NLEV} levels per step"},
        "cpu_baseline": {"value": rate, "unit": "columns/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference (GT4Py) cannot be installed in this image; this is its NumPy-backend execution model restated "
                "(oracle/cloudsc2_numpy.py)",
    }
    try:
        line["cpu_twin_openmp"] = {"value": cpu_twin_rate(8192, 3, args.precision), "unit": "columns/s", "cores": workers,
                                   "kind": "port", "sample": "C++/OpenMP host twin, sat + NL, 8192 columns, best of 3"}
    except Exception as exc:  # pragma: no cover
        line["cpu_twin_openmp"] = {"error": str(exc)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from cloudsc2_b200 import _lib, distributed, iox, setup
    from cloudsc2_b200.framework.config import DataTypes, GridConfig, GT4PyConfig
    from cloudsc2_b200.framework.grid import ComputationalGrid
    from cloudsc2_b200.framework.storage import Field
    from cloudsc2_b200.physics.adjoint.validation import SymmetryTest
    from cloudsc2_b200.physics.common.diagnostics import EtaLevels
    from cloudsc2_b200.physics.common.saturation import Saturation
    from cloudsc2_b200.physics.nonlinear.microphysics import Cloudsc2NL

    if not torch.cuda.is_available():
        raise _lib.CUDAExtensionError("bench.py (GPU arm) needs a CUDA device; there is no CPU fallback")
    rank, world, local_rank = distributed.init_from_env()
    dev = torch.device("cuda", torch.cuda.current_device())
    ncol = args.columns
    np_float = np.float64 if args.precision == "double" else np.float32
    esize = np.dtype(np_float).itemsize
    cfg = GT4PyConfig(dtypes=DataTypes(bool=bool, float=np_float, int=np.int64))
    grid = ComputationalGrid(GridConfig(nx=ncol, ny=1, nz=NLEV))
    col0, _ = distributed.shard_columns(ncol * world, rank, world)
    state = setup.get_synthetic_state(grid, gt4py_config=cfg, column_offset=col0)
    state.update(EtaLevels(grid, gt4py_config=cfg)(state))
    distributed.broadcast_eta(state["f_eta"], src=0)
    p = iox.ifs_defaults()
    dt = timedelta(seconds=3600.0)
    sat = Saturation(grid, 1, True, p["yoethf"], p["yomcst"], gt4py_config=cfg)
    nl = Cloudsc2NL(grid, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrphnc"], gt4py_config=cfg)
    diags = sat(state)
    state.update(diags)
    tends, diags_nl = nl(state, dt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        sat(state, out=diags)
        nl(state, dt, out_tendencies=tends, out_diagnostics=diags_nl)

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(torch.cuda.current_device(), enabled=(rank == 0)) as clocks:
        # the sampler (20 ms period) spans warm-up + timed region: the timed region alone lasts only
        # steps x ~0.5 ms, so the warm-up is repeated until >= 0.25 s of load precede it
        t_load = time.perf_counter()
        nwarm = 0
        while (nwarm < args.warmup or time.perf_counter() - t_load < 0.25
               or (clocks.count() < 3 and time.perf_counter() - t_load < 20.0)):  # the sampler must be running under load
            step()
            nwarm += 1
            if nwarm % 16 == 0:
                torch.cuda.synchronize()
        barrier()
        t_start.record()
        for a, b in ev:
            sat(state, out=diags)
            a.record()
            nl(state, dt, out_tendencies=tends, out_diagnostics=diags_nl)
            b.record()
        t_stop.record()
        barrier()
    elapsed_ms = torch.tensor([t_start.elapsed_time(t_stop)], dtype=torch.float64, device=dev)
    nl_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.steps], dtype=torch.float64, device=dev)
    distributed.allreduce_max_(elapsed_ms)
    distributed.allreduce_max_(nl_ms)
    elapsed_ms, nl_ms = float(elapsed_ms.item()), float(nl_ms.item())
    launches = 2 * args.steps

    # ---- FP64-pipe peak of this device (DFMA micro-benchmark, CUDA events)
    fp64_peak = None
    try:
        lib = _lib.load()
        blocks, iters = 148 * 16, 20000
        scratch = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.cs2_dfma_rate(scratch.data_ptr(), blocks, 200, stream), "cs2_dfma_rate")
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.cs2_dfma_rate(scratch.data_ptr(), blocks, iters, stream), "cs2_dfma_rate")
        b.record()
        torch.cuda.synchronize()
        fp64_peak = blocks * 256 * 8 * iters * 2 / (a.elapsed_time(b) * 1e-3) / 1e12
    except Exception:  # pragma: no cover
        fp64_peak = None

    # ---- kernel-only variants (TL, AD) on the same columns, rank-local, outside the headline region
    variants = {}
    if not args.no_variants:
        stest = SymmetryTest(grid, 0.01, 1, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"],
                             p["yrphnc"], gt4py_config=cfg)
        stest(state, dt, enable_validation=True, verbose=False)
        variants["symmetry_norm3_max_eps"] = stest.norm3_max

        def time_call(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        tl_ms = time_call(lambda: stest.cloudsc2_tl(state, dt, out_tendencies=stest.tends_tl, out_diagnostics=stest.diags_tl))
        ad_ms = time_call(lambda: stest.cloudsc2_ad(state, dt, out_tendencies=stest.tends_ad, out_diagnostics=stest.diags_ad))
        sat_ms = time_call(lambda: sat(state, out=diags))
        peak, _ = measured_peaks()
        for name, ms in (("saturation", sat_ms), ("tl", tl_ms), ("ad", ad_ms)):
            gbs = ELEMS[name] * esize * ncol / (ms * 1e-3) / 1e9
            variants[name] = {"ms": ms, "columns_per_s": ncol / (ms * 1e-3), "achieved_GBs": gbs, "frac_hbm": gbs / peak}
        del stest
        torch.cuda.empty_cache()
        if args.precision == "double":
            # BASELINE.json config 2 names fp32 beside fp64: the same step (saturation + cloudsc2_nl) in single precision
            cfg32 = GT4PyConfig(dtypes=DataTypes(bool=bool, float=np.float32, int=np.int64))
            state32 = setup.get_synthetic_state(grid, gt4py_config=cfg32, column_offset=col0)
            state32.update(EtaLevels(grid, gt4py_config=cfg32)(state32))
            distributed.broadcast_eta(state32["f_eta"], src=0)
            sat32 = Saturation(grid, 1, True, p["yoethf"], p["yomcst"], gt4py_config=cfg32)
            nl32 = Cloudsc2NL(grid, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrphnc"],
                              gt4py_config=cfg32)
            d32 = sat32(state32)
            state32.update(d32)
            t32, g32 = nl32(state32, dt)
            nl32_ms = time_call(lambda: nl32(state32, dt, out_tendencies=t32, out_diagnostics=g32), reps=20)
            step32_ms = time_call(lambda: (sat32(state32, out=d32), nl32(state32, dt, out_tendencies=t32, out_diagnostics=g32)),
                                  reps=20)
            gbs = ELEMS["nl"] * 4 * ncol / (nl32_ms * 1e-3) / 1e9
            variants["nl_fp32"] = {"ms": nl32_ms, "columns_per_s": ncol / (nl32_ms * 1e-3), "achieved_GBs": gbs,
                                   "frac_hbm": gbs / peak, "step_ms": step32_ms, "step_columns_per_s": ncol / (step32_ms * 1e-3)}
            del state32, t32, g32, d32
            torch.cuda.empty_cache()
        if world == 1:
            # BASELINE.json configs 3 and 4: wall time of one Taylor run (10 factors) and one symmetry run on these columns,
            # with the reference orchestration and with the opt-in fused sweeps (host clock around a synchronised run)
            from cloudsc2_b200.physics.tangent_linear.validation import TaylorTest

            def wall(fn, reps=5):
                for _ in range(2):  # the first call allocates the output fields and the workspaces
                    fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(reps):
                    fn()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / reps * 1e3

            f2s = tuple(float(10 ** -(i + 1)) for i in range(10))
            runs = {}
            for label, fused in (("taylor_run_ms", False), ("taylor_run_fused_sums_ms", "sums")):
                pt = iox.ifs_defaults()
                tt = TaylorTest(grid, 0.01, f2s, 1, True, False, pt["yoethf"], pt["yomcst"], pt["yrecldp"], pt["yrephli"],
                                pt["yrncl"], pt["yrphnc"], gt4py_config=cfg, fused=fused)
                runs[label] = wall(lambda: tt.run(state, dt))
                runs[label.replace("_ms", "_penalty")] = tt.validate(tt.run(state, dt), verbose=False)[1]
                del tt
                torch.cuda.empty_cache()
            for label, fused in (("symmetry_run_ms", False), ("symmetry_run_fused_ms", True)):
                ps = iox.ifs_defaults()
                stt = SymmetryTest(grid, 0.01, 1, True, False, ps["yoethf"], ps["yomcst"], ps["yrecldp"], ps["yrephli"],
                                   ps["yrncl"], ps["yrphnc"], gt4py_config=cfg, fused=fused)
                runs[label] = wall(lambda: stt(state, dt, enable_validation=True, verbose=False))
                del stt
                torch.cuda.empty_cache()
            variants["validation_runs"] = runs

    # ---- end to end with HOST buffers through the public host pipeline (cloudsc2_b200.pipeline): the batch lives in
    #      pinned host memory as NPROMA-style column blocks; per block one H2D copy of the 15 packed inputs, the
    #      Saturation + Cloudsc2NL component calls, one D2H copy of the 10 packed outputs; 3 streams, 3 device slots
    from cloudsc2_b200.pipeline import IN_NAMES, NonlinearHostPipeline

    block_cols = min(args.e2e_block, ncol)
    nblocks = -(-ncol // block_cols)
    pipe = NonlinearHostPipeline(block_cols, NLEV, p, gt4py_config=cfg, timestep=dt, eta=state["f_eta"].numpy())
    host_blocks = []
    for b in range(nblocks):
        blk = pipe.alloc_host_block()
        lo, hi = b * block_cols, min((b + 1) * block_cols, ncol)
        for n, name in enumerate(IN_NAMES):
            blk["in"][n, :, : hi - lo].copy_(state[name].buffer[:, lo:hi])
        host_blocks.append(blk)
    h2d = pipe.h2d_bytes_per_block * nblocks
    d2h = pipe.d2h_bytes_per_block * nblocks
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        pipe.run(host_blocks)
    barrier()
    pipe.launches = 0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.run(host_blocks)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    distributed.allreduce_max_(e2e_s)
    e2e_rate = ncol * world * e2e_steps / float(e2e_s.item())
    # the pipelined result must equal the resident-state result (same kernels, same inputs)
    from cloudsc2_b200.pipeline import OUT_NAMES

    ref_t = tends["f_t"].buffer[:, : min(block_cols, ncol)]

    got_t = host_blocks[0]["out"][OUT_NAMES.index("f_t")][:, : min(block_cols, ncol)].to(dev)
    if not torch.equal(got_t, ref_t):
        raise RuntimeError("host pipeline result differs from the resident-state result")

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    nl_bytes = ELEMS["nl"] * esize * ncol
    achieved = nl_bytes / (nl_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": ncol * world * args.steps / (elapsed_ms * 1e-3), "unit": "columns/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if args.precision == "double" else "f32", "data": "synthetic",
        "config": {
            "workload": f"CLOUDSC2-NL {args.precision}: saturation + cloudsc2_nl per step, {ncol} columns x {NLEV} levels per GPU "
                        f"(BASELINE.json configs[1]), synthetic block tiled from seed 0",
            "columns_per_gpu": ncol, "levels": NLEV, "parallelism": f"columns sharded over {world} GPU(s), no collective",
            "l2": "per-step working set (27 fields x 72 MB) exceeds the 126 MB L2; no flush needed",
        },
        "roofline": {
            "kernel": "cloudsc2_nl", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": ncu_traffic("nl_kernel"),
            "algorithmic_bytes_per_column": ELEMS["nl"] * esize, "kernel_ms": nl_ms,
            "kernel_columns_per_s": ncol / (nl_ms * 1e-3), "peak_source": peak_src,
            # second axis of the roofline: algorithmic flops (334 arithmetic ops + 15 transcendental calls per
            # point, all branches) against the DFMA rate measured on this device just before
            "fp64": {"peak_tflops_measured": fp64_peak, "algorithmic_flops_per_column": 349 * NLEV,
                     "achieved_tflops": 349 * NLEV * ncol / (nl_ms * 1e-3) / 1e12,
                     "frac": (349 * NLEV * ncol / (nl_ms * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
                     "note": "bytes are the slower axis: the HBM fraction is the roofline fraction"},
        },
        "variants": variants,
        "e2e": {"value": e2e_rate, "unit": "columns/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "block_columns": block_cols, "gpu_launches": pipe.launches,
                "what": "NonlinearHostPipeline: pinned host column blocks -> H2D (15 inputs) -> Saturation + Cloudsc2NL "
                        "components -> D2H (10 outputs), 3 streams / 3 device slots, copies overlapped with kernels"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        workers = 1
        cols = 4096
        rate, ms = cpu_oracle_rate(cols, workers, 3, 1, args.precision)
        line["cpu_baseline"] = {
            "value": rate, "unit": "columns/s", "cores": workers, "kind": "port",
            "sample": f"NumPy oracle (reference numpy-backend execution model), saturation + cloudsc2_nl, {cols} columns x {NLEV} "
                      f"levels, 3 runs after 1 warm-up, single process",
        }
        try:
            line["cpu_twin_openmp"] = {"value": cpu_twin_rate(16384, 3, args.precision), "unit": "columns/s",
                                       "cores": os.cpu_count(), "kind": "port",
                                       "sample": "C++/OpenMP host twin of the column code, sat + NL, 16384 columns, best of 3"}
        except Exception as exc:  # pragma: no cover
            line["cpu_twin_openmp"] = {"error": str(exc)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--columns", type=int, default=65536, help="columns per GPU")
    ap.add_argument("--precision", choices=("double", "single"), default="double")
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--e2e-block", type=int, default=4096, help="columns per host block of the e2e pipeline")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
