"""Multi-GPU worker (one process per GPU, NCCL): sharded Taylor and symmetry tests.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_gpu_worker.py [--columns C]

Every rank owns a contiguous block of the C global columns (tiled synthetic block), eta comes from global
column 0 (broadcast), the Taylor sums are all-reduced once (SUM) and the symmetry residual once (MAX).  Rank 0
checks the result against a single-GPU run of the same global problem on its own device."""
import argparse
import os
import sys
from datetime import timedelta

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.join(HERE, ".."), os.path.join(HERE, "..", "gt4py-dwarf-p-cloudsc2-tl-ad_b200")]

import gpu_harness as G  # noqa: E402
from cloudsc2_b200 import distributed, iox, setup  # noqa: E402
from cloudsc2_b200.framework.config import GridConfig  # noqa: E402
from cloudsc2_b200.framework.grid import ComputationalGrid  # noqa: E402
from cloudsc2_b200.physics.adjoint.validation import SymmetryTest  # noqa: E402
from cloudsc2_b200.physics.common.diagnostics import EtaLevels  # noqa: E402
from cloudsc2_b200.physics.tangent_linear.validation import TaylorTest  # noqa: E402

F2S = tuple(float(10 ** -(i + 1)) for i in range(10))


def run(ncol_local, col0, cfg, sharded):
    grid = ComputationalGrid(GridConfig(nx=ncol_local, ny=1, nz=137))
    state = setup.get_synthetic_state(grid, gt4py_config=cfg, column_offset=col0)
    state.update(EtaLevels(grid, gt4py_config=cfg)(state))
    if sharded:
        distributed.broadcast_eta(state["f_eta"], src=0)
    p = iox.ifs_defaults()
    dt = timedelta(seconds=3600)
    tt = TaylorTest(grid, 0.01, F2S, 1, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"],
                    p["yrphnc"], gt4py_config=cfg)
    import contextlib

    with (contextlib.nullcontext() if sharded else distributed.local_only()):  # control run: no collectives
        norms = tt.run(state, dt)
        p = iox.ifs_defaults()
        state2 = setup.get_synthetic_state(grid, gt4py_config=cfg, column_offset=col0)
        state2["f_eta"] = state["f_eta"]
        st = SymmetryTest(grid, 0.01, 1, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"],
                          p["yrphnc"], gt4py_config=cfg, ad_predicates="tl")  # the base block crosses RTT inside levels
        passed = st(state2, dt, verbose=False)
    return tt, norms, st, passed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=4000)
    args = ap.parse_args()
    rank, world, _ = distributed.init_from_env()
    cfg = G.config_for(np.float64)
    lo, hi = distributed.shard_columns(args.columns, rank, world)
    tt, norms, st, passed = run(hi - lo, lo, cfg, sharded=True)
    ok, code = tt.validate(norms.copy(), verbose=False)
    assert ok and code <= 5, (norms, code)
    assert passed and st.norm3_max < 1e4, st.norm3_max
    if rank == 0:
        _, norms1, st1, passed1 = run(args.columns, 0, cfg, sharded=False)
        np.testing.assert_allclose(norms, norms1, rtol=1e-9)
        assert abs(st.norm3_max - st1.norm3_max) <= 1e-9 * max(1.0, st1.norm3_max), (st.norm3_max, st1.norm3_max)
        print(f"DIST_GPU_OK world={world} columns={args.columns} taylor_penalty={code} "
              f"symmetry_max_eps={st.norm3_max:.1f} norms[3:6]={norms[3:6]}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
