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


def ncu_traffic(kernel: str, columns: int, dtype: str):
    """DRAM bytes (read + write) of ONE launch of `kernel` on `columns` columns in `dtype` from the committed ncu captures
    (profiles/ncu_traffic.json, keys "kernel|columns|dtype"), or None when no capture matches."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(f"{kernel}|{columns}|{dtype}")
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
                "(oracle/cloudsc2_numpy.py; bit-identical in fp64 to the reference's own stencil sources executed under that "
                "model, tests/test_ref_exec.py).  This arm loads no product code: no CUDA library, no kernels.",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def _events():
    import torch

    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


class _Ranks:
    """barrier + max-over-ranks helpers (no-ops for one process)."""

    def __init__(self, world, dev):
        self.world, self.dev = world, dev

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, value: float) -> float:
        import torch

        from cloudsc2_b200 import distributed

        t = torch.tensor([value], dtype=torch.float64, device=self.dev)
        distributed.allreduce_max_(t)
        return float(t.item())

    def timed(self, fn, reps: int, warm: int = 2) -> float:
        """ms per call: CUDA events on the current stream around `reps` calls, barrier + synchronize on both sides,
        MAX over ranks."""
        import torch

        for _ in range(warm):
            fn()
        self.barrier()
        a, b = _events()
        try:  # ~0.2 ms of device-side spinning queued ahead of the start event: the host gets the first calls enqueued while the
            torch.cuda._sleep(400_000)  # device is still busy, so the timed region holds GPU work only (no launch gap at its head)
        except Exception:  # pragma: no cover
            pass
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return self.max(a.elapsed_time(b) / reps)


def _free():
    import gc

    import torch

    gc.collect()
    torch.cuda.empty_cache()


def _field_bytes(ncol: int, esize: int) -> int:
    return (NLEV + 1) * (-(-ncol // 32) * 32) * esize


def _validation_runs(R, grid, state, dt, cfg, iox, fused_too=True, reps=3):
    """Wall of one Taylor run (10 factors) and one symmetry run on `state`, collectives included (CUDA events, max over
    ranks): reference orchestration and the opt-in fused sweeps.  `state` is not modified (the harnesses get a copy of the
    dict)."""
    from cloudsc2_b200.physics.adjoint.validation import SymmetryTest
    from cloudsc2_b200.physics.tangent_linear.validation import TaylorTest

    f2s = tuple(float(10 ** -(i + 1)) for i in range(10))
    runs = {}
    for label, fused in (("taylor_run_ms", False), ("taylor_run_fused_sums_ms", "sums")):
        if fused and not fused_too:
            continue
        pt = iox.ifs_defaults()  # TaylorTest switches LREGCL off in the models it is given (tangent_linear/validation.py:84-85)
        s = dict(state)
        tt = TaylorTest(grid, 0.01, f2s, 1, True, False, pt["yoethf"], pt["yomcst"], pt["yrecldp"], pt["yrephli"], pt["yrncl"],
                        pt["yrphnc"], gt4py_config=cfg, fused=fused)
        norms = tt.run(s, dt)
        runs[label.replace("_ms", "_penalty")] = tt.validate(norms, verbose=False)[1]
        runs[label] = R.timed(lambda: tt.run(s, dt), reps, warm=1)
        del tt, s
        _free()
    for label, fused in (("symmetry_run_ms", False), ("symmetry_run_fused_ms", True)):
        if fused and not fused_too:
            continue
        ps = iox.ifs_defaults()
        s = dict(state)
        # the bench block crosses RTT inside levels: the exact-adjoint predicates ("tl"); timing is the same for both modes
        stt = SymmetryTest(grid, 0.01, 1, True, False, ps["yoethf"], ps["yomcst"], ps["yrecldp"], ps["yrephli"], ps["yrncl"],
                           ps["yrphnc"], gt4py_config=cfg, fused=fused, ad_predicates="tl")
        stt(s, dt, enable_validation=True, verbose=False)
        runs[label.replace("_run", "_norm3_max_eps").replace("_ms", "")] = stt.norm3_max
        runs[label] = R.timed(lambda: stt(s, dt, enable_validation=True, verbose=False), reps, warm=1)
        del stt, s
        _free()
    return runs


def run_config5(args, R, rank, world, cfg, p, dt):
    """BASELINE.json configs[4]: NL + TL + AD on 1 048 576 columns sharded over the ranks (1 M / N each), the Taylor test
    (ONE all-reduce SUM of 200 doubles) and the symmetry test (ONE all-reduce MAX of 1 double) with their collectives inside
    the timed region, and a control: the all-reduced results of a sharded 4 000-column problem equal the un-sharded ones."""
    import torch

    from cloudsc2_b200 import distributed, iox, setup
    from cloudsc2_b200.framework.config import GridConfig
    from cloudsc2_b200.framework.grid import ComputationalGrid
    from cloudsc2_b200.physics.adjoint.validation import SymmetryTest
    from cloudsc2_b200.physics.common.diagnostics import EtaLevels
    from cloudsc2_b200.physics.common.saturation import Saturation
    from cloudsc2_b200.physics.nonlinear.microphysics import Cloudsc2NL
    from cloudsc2_b200.physics.tangent_linear.validation import TaylorTest

    total = args.config5_columns
    esize = np.dtype(cfg.dtypes.float).itemsize
    out = {"total_columns": total, "n_gpus": world}

    def problem(ntotal, sharded=True):
        c0, c1 = distributed.shard_columns(ntotal, rank, world) if sharded else (0, ntotal)
        grid = ComputationalGrid(GridConfig(nx=c1 - c0, ny=1, nz=NLEV))
        state = setup.get_synthetic_state(grid, gt4py_config=cfg, column_offset=c0)
        if c1 - c0 > 0 and c0 == 0:
            state.update(EtaLevels(grid, gt4py_config=cfg)(state))
        else:  # eta comes from GLOBAL column 0 (common/diagnostics.py:45), which only rank 0 owns
            g1 = ComputationalGrid(GridConfig(nx=1, ny=1, nz=NLEV))
            state.update(EtaLevels(g1, gt4py_config=cfg)(setup.get_synthetic_state(g1, gt4py_config=cfg)))
        if sharded:
            distributed.broadcast_eta(state["f_eta"], src=0)
        return grid, state

    # ---- control: sharded + all-reduced == un-sharded, 4 000 columns
    f2s = tuple(float(10 ** -(i + 1)) for i in range(10))

    def taylor_and_symmetry(grid, state):
        pt = iox.ifs_defaults()
        tt = TaylorTest(grid, 0.01, f2s, 1, True, False, pt["yoethf"], pt["yomcst"], pt["yrecldp"], pt["yrephli"], pt["yrncl"],
                        pt["yrphnc"], gt4py_config=cfg)
        norms = tt.run(dict(state), dt)
        ps = iox.ifs_defaults()
        st = SymmetryTest(grid, 0.01, 1, True, False, ps["yoethf"], ps["yomcst"], ps["yrecldp"], ps["yrephli"], ps["yrncl"],
                          ps["yrphnc"], gt4py_config=cfg, ad_predicates="tl")
        st(dict(state), dt, enable_validation=True, verbose=False)
        return norms, st.norm3_max

    norms_sh, n3_sh = taylor_and_symmetry(*problem(4000, sharded=True))
    with distributed.local_only():
        norms_lo, n3_lo = taylor_and_symmetry(*problem(4000, sharded=False))
    ok = bool(np.allclose(norms_sh, norms_lo, rtol=1e-9, atol=0.0)) and abs(n3_sh - n3_lo) <= 1e-12 * max(abs(n3_lo), 1.0)
    out["control_4000_columns"] = {
        "sharded_equals_unsharded": ok, "taylor_norms_max_rel_diff": float(np.max(np.abs(norms_sh / norms_lo - 1.0))),
        "symmetry_norm3_max_eps": n3_sh, "symmetry_norm3_max_eps_unsharded": n3_lo,
    }
    if not ok:
        raise RuntimeError(f"config 5 control failed: sharded {norms_sh}, {n3_sh} vs un-sharded {norms_lo}, {n3_lo}")
    _free()

    # ---- the 1 M-column problem
    grid, state = problem(total)
    ncol = grid.nx
    out["columns_per_gpu"] = ncol
    sat = Saturation(grid, 1, True, p["yoethf"], p["yomcst"], gt4py_config=cfg)
    state.update(sat(state))
    nl = Cloudsc2NL(grid, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrphnc"], gt4py_config=cfg)
    tn, dg = nl(state, dt)
    peak, _ = measured_peaks()
    kern = {}
    kern["saturation"] = R.timed(lambda: sat(state, out={"f_qsat": state["f_qsat"]}), 5)
    kern["nl"] = R.timed(lambda: nl(state, dt, out_tendencies=tn, out_diagnostics=dg), 5)
    del nl, tn, dg
    _free()
    s = dict(state)
    st = SymmetryTest(grid, 0.01, 1, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"],
                      p["yrphnc"], gt4py_config=cfg, ad_predicates="tl")
    st(s, dt, enable_validation=True, verbose=False)
    out["symmetry_norm3_max_eps"] = st.norm3_max
    kern["tl"] = R.timed(lambda: st.cloudsc2_tl(s, dt, out_tendencies=st.tends_tl, out_diagnostics=st.diags_tl), 5)
    kern["ad"] = R.timed(lambda: st.cloudsc2_ad(s, dt, out_tendencies=st.tends_ad, out_diagnostics=st.diags_ad), 5)
    out["symmetry_run_ms"] = R.timed(lambda: st(s, dt, enable_validation=True, verbose=False), 3, warm=1)
    del st, s
    _free()
    out["kernels"] = {
        k: {"ms": ms, "columns_per_s": total / (ms * 1e-3),
            "frac_hbm": ELEMS[k] * esize * ncol / (ms * 1e-3) / 1e9 / peak} for k, ms in kern.items()
    }
    # Taylor test: the reference orchestration needs ~90 resident fields; fall back to the fused sweeps if they do not fit
    free_b, _ = torch.cuda.mem_get_info()
    need = 75 * _field_bytes(ncol, esize)
    modes = (("taylor_run_ms", False),) if free_b > 1.15 * need else ()
    modes += (("taylor_run_fused_sums_ms", "sums"),)
    for label, fused in modes:
        pt = iox.ifs_defaults()
        s = dict(state)
        tt = TaylorTest(grid, 0.01, f2s, 1, True, False, pt["yoethf"], pt["yomcst"], pt["yrecldp"], pt["yrephli"], pt["yrncl"],
                        pt["yrphnc"], gt4py_config=cfg, fused=fused)
        norms = tt.run(s, dt)
        out[label.replace("_ms", "_penalty")] = tt.validate(norms, verbose=False)[1]
        out[label] = R.timed(lambda: tt.run(s, dt), 2, warm=0)
        del tt, s
        _free()
    # the collectives on their own: ONE all-reduce (SUM) of the Taylor sums, ONE all-reduce (MAX) of the symmetry residual
    buf = torch.zeros(200, dtype=torch.float64, device=R.dev)
    one = torch.zeros(1, dtype=torch.float64, device=R.dev)
    out["allreduce_sum_200_doubles_us"] = 1e3 * R.timed(lambda: distributed.allreduce_sum_(buf), 50, warm=5)
    out["allreduce_max_1_double_us"] = 1e3 * R.timed(lambda: distributed.allreduce_max_(one), 50, warm=5)
    out["collective"] = ("NCCL all-reduce over NVLink: SUM of 200 doubles once per Taylor test, MAX of 1 double once per symmetry "
                         "test; broadcast of 137 eta values at set-up" if world > 1 else "none (one rank)")
    del state
    _free()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from cloudsc2_b200 import _lib, distributed, iox, setup
    from cloudsc2_b200.framework.config import DataTypes, GridConfig, GT4PyConfig
    from cloudsc2_b200.framework.grid import ComputationalGrid
    from cloudsc2_b200.physics.adjoint.validation import SymmetryTest
    from cloudsc2_b200.physics.common.diagnostics import EtaLevels
    from cloudsc2_b200.physics.common.saturation import Saturation
    from cloudsc2_b200.physics.nonlinear.microphysics import Cloudsc2NL

    if not torch.cuda.is_available():
        raise _lib.CUDAExtensionError("bench.py (GPU arm) needs a CUDA device; there is no CPU fallback")
    rank, world, local_rank = distributed.init_from_env()
    dev = torch.device("cuda", torch.cuda.current_device())
    R = _Ranks(world, dev)
    ncol = args.columns
    np_float = np.float64 if args.precision == "double" else np.float32
    esize = np.dtype(np_float).itemsize
    dname = "f64" if args.precision == "double" else "f32"
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

    def step():
        sat(state, out=diags)
        nl(state, dt, out_tendencies=tends, out_diagnostics=diags_nl)

    ev = [_events() for _ in range(args.steps)]
    t_start, t_stop = _events()
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
        R.barrier()
        t_start.record()
        for a, b in ev:
            sat(state, out=diags)
            a.record()
            nl(state, dt, out_tendencies=tends, out_diagnostics=diags_nl)
            b.record()
        t_stop.record()
        R.barrier()
    elapsed_ms = R.max(t_start.elapsed_time(t_stop))
    nl_ms = R.max(sum(a.elapsed_time(b) for a, b in ev) / args.steps)
    launches = 2 * args.steps

    # ---- FP64-pipe peak of this device (DFMA micro-benchmark, CUDA events)
    fp64_peak = None
    try:
        lib = _lib.load()
        blocks, iters = 148 * 16, 20000
        scratch = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.cs2_dfma_rate(scratch.data_ptr(), blocks, 200, stream), "cs2_dfma_rate")
        a, b = _events()
        a.record()
        _lib.check(lib.cs2_dfma_rate(scratch.data_ptr(), blocks, iters, stream), "cs2_dfma_rate")
        b.record()
        torch.cuda.synchronize()
        fp64_peak = blocks * 256 * 8 * iters * 2 / (a.elapsed_time(b) * 1e-3) / 1e12
        del scratch
    except Exception:  # pragma: no cover
        fp64_peak = None

    peak, peak_src = measured_peaks()

    def kernel_entry(name, kernel, ms, n=ncol, es=esize, dn=dname):
        gbs = ELEMS[name] * es * n / (ms * 1e-3) / 1e9
        return {"name": name, "kernel": kernel, "columns": n, "dtype": dn, "ms": ms, "columns_per_s": n / (ms * 1e-3),
                "algorithmic_bytes": ELEMS[name] * es * n, "achieved": gbs, "unit": "GB/s", "frac": gbs / peak,
                "traffic": ncu_traffic(kernel, n, dn)}

    # ---- every kernel of the path on the same columns (kernel-only, CUDA events, max over ranks), outside the headline region
    kernels = [kernel_entry("nl", "nl_kernel", nl_ms)]
    variants = {}
    if not args.no_variants:
        s = dict(state)
        stest = SymmetryTest(grid, 0.01, 1, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"],
                             p["yrphnc"], gt4py_config=cfg, ad_predicates="tl")
        stest(s, dt, enable_validation=True, verbose=False)
        variants["symmetry_norm3_max_eps"] = stest.norm3_max
        tl_ms = R.timed(lambda: stest.cloudsc2_tl(s, dt, out_tendencies=stest.tends_tl, out_diagnostics=stest.diags_tl), 30, 3)
        ad_ms = R.timed(lambda: stest.cloudsc2_ad(s, dt, out_tendencies=stest.tends_ad, out_diagnostics=stest.diags_ad), 30, 3)
        sat_ms = R.timed(lambda: sat(state, out=diags), 50, 3)
        kernels += [kernel_entry("tl", "tl_kernel", tl_ms), kernel_entry("ad", "nl_kernel<LIN>+ad_bwd_kernel", ad_ms),
                    kernel_entry("saturation", "saturation_kernel", sat_ms)]
        # the streaming helpers and the reductions of the two harnesses (unfused orchestration), same columns
        from cloudsc2_b200.physics.common.increment import PerturbedState
        from cloudsc2_b200.reductions import TaylorSums

        ELEMS.update(state_increment=32 * (NLEV + 1), perturbed_state=48 * (NLEV + 1), taylor_sums=30 * (NLEV + 1),
                     symmetry_norm1=10 * (NLEV + 1), symmetry_norm2=32 * (NLEV + 1))
        pert = PerturbedState(grid, 1e-3, gt4py_config=cfg)
        state_p = pert(s)
        tsum, tbuf = TaylorSums(), torch.zeros(20, dtype=torch.float64, device=dev)
        tl_f = [stest.tends_tl[n] for n in ("f_t", "f_q", "f_ql", "f_qi")] + [stest.diags_tl[n] for n in
                                                                           ("f_clc", "f_fhpsl", "f_fhpsn", "f_fplsl", "f_fplsn", "f_covptot")]
        tl_i = [stest.tends_tl[n + "_i"] for n in ("f_t", "f_q", "f_ql", "f_qi")] + [stest.diags_tl[n + "_i"] for n in
                                                                                  ("f_clc", "f_fhpsl", "f_fhpsn", "f_fplsl", "f_fplsn", "f_covptot")]
        nl_f = [tends[n] for n in ("f_t", "f_q", "f_ql", "f_qi")] + [diags_nl[n] for n in
                                                                  ("f_clc", "f_fhpsl", "f_fhpsn", "f_fplsl", "f_fplsn", "f_covptot")]
        for name, kernel, fn in (
            ("state_increment", "state_increment_kernel", lambda: stest.state_increment(s, out=stest.state_i)),
            ("perturbed_state", "perturbed_state_kernel", lambda: pert(s, out=state_p)),
            ("taylor_sums", "taylor_partial_kernel", lambda: tsum(tl_f, nl_f, tl_i, tbuf)),  # 30 distinct fields, as in a Taylor run
            ("symmetry_norm1", "symmetry_norm_kernel[norm1]", lambda: stest.get_norm1(stest.tends_tl, stest.diags_tl)),
            ("symmetry_norm2", "symmetry_norm_kernel[norm2]", lambda: stest.get_norm2(stest.state_i, stest.tends_ad, stest.diags_ad)),
        ):
            kernels.append(kernel_entry(name, kernel, R.timed(fn, 10, 3)))
        del pert, state_p
        for k in kernels[1:4]:  # kept under the round-1 key names as well
            variants[k["name"]] = {"ms": k["ms"], "columns_per_s": k["columns_per_s"], "achieved_GBs": k["achieved"], "frac_hbm": k["frac"]}
        del stest, s
        _free()
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
            nl32_ms = R.timed(lambda: nl32(state32, dt, out_tendencies=t32, out_diagnostics=g32), 50, 3)
            step32_ms = R.timed(lambda: (sat32(state32, out=d32), nl32(state32, dt, out_tendencies=t32, out_diagnostics=g32)), 50, 3)
            k32 = kernel_entry("nl", "nl_kernel", nl32_ms, es=4, dn="f32")
            kernels.append(k32)
            variants["nl_fp32"] = {"ms": nl32_ms, "columns_per_s": k32["columns_per_s"], "achieved_GBs": k32["achieved"],
                                   "frac_hbm": k32["frac"], "step_ms": step32_ms, "step_columns_per_s": ncol / (step32_ms * 1e-3)}
            del state32, t32, g32, d32, nl32, sat32
            _free()
        # BASELINE.json configs 3 and 4 at the bench size: one Taylor run (10 factors) and one symmetry run, collectives inside
        variants["validation_runs"] = _validation_runs(R, grid, state, dt, cfg, iox)

    # ---- BASELINE.json configs[4]: 1 M columns sharded over the ranks, NL + TL + AD + the two tests with their all-reduces
    config5 = None
    if not args.no_config5:
        del tends, diags_nl
        _free()
        config5 = run_config5(args, R, rank, world, cfg, p, dt)
        tends, diags_nl = nl(state, dt)

    # ---- end to end with HOST buffers through the public host pipeline (cloudsc2_b200.pipeline): the batch lives in
    #      pinned host memory as NPROMA-style column blocks; per block one H2D copy of the 15 packed inputs, the
    #      Saturation + Cloudsc2NL component calls, one D2H copy of the packed outputs; 3 streams, 3 device slots
    from cloudsc2_b200.pipeline import IN_NAMES, OUT_NAMES, NonlinearHostPipeline

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
    R.barrier()
    pipe.launches = 0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.run(host_blocks, sync=False)
    torch.cuda.synchronize()
    e2e_s = R.max(time.perf_counter() - t0)
    e2e_rate = ncol * world * e2e_steps / e2e_s
    # the pipelined result must equal the resident-state result (same kernels, same inputs)
    ref_t = tends["f_t"].buffer[:, : min(block_cols, ncol)]
    got_t = host_blocks[0]["out"][OUT_NAMES.index("f_t")][:, : min(block_cols, ncol)].to(dev)
    if not torch.equal(got_t, ref_t):
        raise RuntimeError("host pipeline result differs from the resident-state result")
    # the ceiling of this path on this box: the same pinned blocks copied in and out (both directions at once), no kernels
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    slot = pipe.slots[0]
    R.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for blk in host_blocks:
            with torch.cuda.stream(s_in):
                slot.inp.copy_(blk["in"], non_blocking=True)
            with torch.cuda.stream(s_out):
                blk["out"][: pipe.nout_copied].copy_(slot.out[: pipe.nout_copied], non_blocking=True)
    torch.cuda.synchronize()
    copy_s = R.max(time.perf_counter() - t0)
    pipe.run(host_blocks)  # leave the blocks' outputs consistent again

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    nl_bytes = ELEMS["nl"] * esize * ncol
    achieved = nl_bytes / (nl_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": ncol * world * args.steps / (elapsed_ms * 1e-3), "unit": "columns/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dname, "data": "synthetic",
        "config": {
            "workload": f"CLOUDSC2-NL {args.precision}: saturation + cloudsc2_nl per step, {ncol} columns x {NLEV} levels per GPU "
                        f"(BASELINE.json configs[1]), synthetic block tiled from seed 0",
            "columns_per_gpu": ncol, "levels": NLEV,
            "parallelism": f"columns sharded over {world} GPU(s); the headline step has no collective (columns are independent); "
                           f"the `config5` block runs 1 M columns sharded over the same ranks with the Taylor / symmetry all-reduces",
            "l2": "per-step working set (27 fields x 72 MB) exceeds the 126 MB L2; no flush needed",
        },
        "roofline": {
            "kernel": "cloudsc2_nl", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": ncu_traffic("nl_kernel", ncol, dname),
            "algorithmic_bytes_per_column": ELEMS["nl"] * esize, "kernel_ms": nl_ms,
            "kernel_columns_per_s": ncol / (nl_ms * 1e-3), "peak_source": peak_src,
            # every kernel of the path at the bench size: time (CUDA events), algorithmic bytes, fraction of the measured HBM
            # peak, DRAM bytes per launch from the committed ncu capture of the same kernel / columns / dtype (null: no capture)
            "kernels": kernels,
            "quantisation": "65 536 columns = 3.46 warps per scheduler: the kernels cost what 75 776 columns (4 warps) cost; "
                            "per column they are 12-14 % better at multiples of 18 944 columns per GPU (profiles/r2b_nl_pipeline.md)",
            # second axis of the roofline: algorithmic flops (334 arithmetic ops + 15 transcendental calls per
            # point, all branches) against the DFMA rate measured on this device just before
            "fp64": {"peak_tflops_measured": fp64_peak, "algorithmic_flops_per_column": 349 * NLEV,
                     "achieved_tflops": 349 * NLEV * ncol / (nl_ms * 1e-3) / 1e12,
                     "frac": (349 * NLEV * ncol / (nl_ms * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
                     "note": "bytes are the slower axis: the HBM fraction is the roofline fraction"},
        },
        "variants": variants,
        "e2e": {"value": e2e_rate, "unit": "columns/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "steps": e2e_steps, "block_columns": block_cols, "gpu_launches": pipe.launches,
                "per_rank": {"h2d_GBs": h2d * e2e_steps / e2e_s / 1e9, "d2h_GBs": d2h * e2e_steps / e2e_s / 1e9},
                "copy_only_ceiling": {"columns_per_s": ncol * world * e2e_steps / copy_s,
                                      "per_rank_h2d_GBs": h2d * e2e_steps / copy_s / 1e9, "per_rank_d2h_GBs": d2h * e2e_steps / copy_s / 1e9,
                                      "what": "the same pinned blocks copied H2D and D2H at once on two streams, no kernels: what "
                                              "PCIe and the host memory system of this box allow for these bytes (max over ranks)"},
                "what": "NonlinearHostPipeline: pinned host column blocks -> H2D (15 inputs) -> Saturation + Cloudsc2NL "
                        "components -> D2H (9 outputs; f_covptot is identically 0 with the default flags and stays a zero plane "
                        "on the host), 3 streams / 3 device slots, copies overlapped with kernels"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    if config5 is not None:
        line["config5"] = config5
    if world == 1 and not args.no_cpu_baseline:
        workers = 1
        cols = 4096
        rate, ms = cpu_oracle_rate(cols, workers, 3, 1, args.precision)
        line["cpu_baseline"] = {
            "value": rate, "unit": "columns/s", "cores": workers, "kind": "port",
            "sample": f"NumPy oracle (reference numpy-backend execution model; bit-identical in fp64 to the reference's stencil "
                      f"sources run under the same model, tests/test_ref_exec.py), saturation + cloudsc2_nl, {cols} columns x {NLEV} "
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
    ap.add_argument("--no-config5", action="store_true", help="skip the 1 M-column NL + TL + AD + Taylor + symmetry block")
    ap.add_argument("--config5-columns", type=int, default=1 << 20, help="TOTAL columns of the config-5 block (sharded over the ranks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries exactly ONE line, the JSON record: whatever libraries write to file descriptor 1 while the run is set
    # up (e.g. "NCCL version ..." at communicator creation) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
