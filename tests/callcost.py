import sys, os, time
ROOT = os.getcwd()
sys.path[:0] = [ROOT, os.path.join(ROOT, "gt4py-dwarf-p-cloudsc2-tl-ad_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from datetime import timedelta
from cloudsc2_b200 import iox, setup
from cloudsc2_b200.framework.config import DataTypes, GridConfig, GT4PyConfig
from cloudsc2_b200.framework.grid import ComputationalGrid
from cloudsc2_b200.physics.common.saturation import Saturation
from cloudsc2_b200.physics.common.diagnostics import EtaLevels
from cloudsc2_b200.physics.nonlinear.microphysics import Cloudsc2NL
for dt_np in (np.float64, np.float32):
    cfg = GT4PyConfig(dtypes=DataTypes(bool=bool, float=dt_np, int=np.int64))
    grid = ComputationalGrid(GridConfig(nx=65536, ny=1, nz=137))
    state = setup.get_synthetic_state(grid, gt4py_config=cfg)
    state.update(EtaLevels(grid, gt4py_config=cfg)(state))
    p = iox.ifs_defaults(); dt = timedelta(seconds=3600.0)
    sat = Saturation(grid, 1, True, p["yoethf"], p["yomcst"], gt4py_config=cfg)
    nl = Cloudsc2NL(grid, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrphnc"], gt4py_config=cfg)
    d = sat(state); state.update(d)
    t, g = nl(state, dt)
    for _ in range(5): nl(state, dt, out_tendencies=t, out_diagnostics=g)
    torch.cuda.synchronize()
    for reps in (20, 200):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for _ in range(reps): nl(state, dt, out_tendencies=t, out_diagnostics=g)
        b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(dt_np.__name__, reps, "host enqueue ms/call %.4f" % ((t1-t0)*1e3/reps), "gpu ms/call %.4f" % (a.elapsed_time(b)/reps), "wall %.4f" % ((t2-t0)*1e3/reps))
if "--profile" in sys.argv:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(300): nl(state, dt, out_tendencies=t, out_diagnostics=g)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
