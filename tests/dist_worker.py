"""world_size-2 gloo worker for tests/test_host_logic.py: exercises the sharded Taylor / symmetry
reduction plumbing of cloudsc2_b200.distributed on CPU, with the oracle standing in for the kernels."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.join(HERE, ".."), os.path.join(HERE, "..", "gt4py-dwarf-p-cloudsc2-tl-ad_b200")]

import helpers as H  # noqa: E402
from cloudsc2_b200 import distributed  # noqa: E402
from cloudsc2_b200.framework.config import GridConfig, GT4PyConfig  # noqa: E402
from cloudsc2_b200.framework.grid import ComputationalGrid, K  # noqa: E402
from cloudsc2_b200.framework.storage import zeros  # noqa: E402


def main():
    rank, world, _ = distributed.init_from_env(backend="gloo")
    assert world == 2 and distributed.is_distributed()
    nx = 64
    lo, hi = distributed.shard_columns(nx, rank, world)
    P = H.externals()
    full = H.with_diagnostics(H.make_state("base", ncol=nx), P)
    # eta comes from GLOBAL column 0: rank 1 must receive rank 0's
    cfg = GT4PyConfig(device="cpu")
    grid = ComputationalGrid(GridConfig(nx=hi - lo, ny=1, nz=137))
    eta = zeros(grid, (K,), gt4py_config=cfg)
    local = {k: (v if v.ndim == 1 else np.ascontiguousarray(v[:, lo:hi])) for k, v in full.items()}
    eta.assign(H.onp.eta_levels(local["f_ap"], local["f_aph"]))
    distributed.broadcast_eta(eta, src=0)
    assert np.array_equal(eta.numpy(), full["f_eta"])
    local["f_eta"] = eta.numpy()
    # local Taylor-style sums, then ONE all-reduce
    tn, dg = H.onp.cloudsc2_nl(local, H.DT, P)
    sums = torch.tensor([float(np.sum(tn["f_t"])), float(np.sum(dg["f_clc"]))], dtype=torch.float64)
    distributed.allreduce_sum_(sums)
    tn_full, dg_full = H.onp.cloudsc2_nl(full, H.DT, P)
    np.testing.assert_allclose(sums.numpy(), [np.sum(tn_full["f_t"]), np.sum(dg_full["f_clc"])], rtol=1e-12)
    # sharded results equal the global run bit for bit (columns are independent)
    assert np.array_equal(tn["f_t"], tn_full["f_t"][:, lo:hi])
    mx = torch.tensor([float(rank + 1)], dtype=torch.float64)
    distributed.allreduce_max_(mx)
    assert mx.item() == 2.0
    dist.barrier()
    print(f"DIST_OK rank={rank}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
