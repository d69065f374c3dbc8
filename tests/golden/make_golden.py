"""Regenerates the fixtures in this directory.  Run from the repo root in the authoring
container (needs /root/reference for part 1; part 2 only needs the oracle):

    python tests/golden/make_golden.py

1. `reference_double.npz` / `reference_single.npz`: the datasets of the reference's golden
   NL outputs `data/reference_{double,single}.h5` (read with cloudsc2_b200.h5lite, stored
   loss-free with np.savez_compressed).  These are DATA published by the reference for
   validation (drivers/run_nonlinear.py:139-147), not source code.
2. `ref_*.npz` (helpers.REF_FIXTURES): outputs of the REFERENCE'S OWN stencil sources
   (`/root/reference/src/cloudsc2_gt4py/physics/**/_stencils/*.py`, unmodified, executed by
   oracle/gtscript_exec.py through oracle/ref_run.py) for saturation, NL, state_increment,
   perturbed_state, TL and AD on the seeded synthetic blocks of cloudsc2_b200.synthetic,
   fp64 and fp32, default and non-default flags.  These pin the oracle AND the CUDA kernels
   to outputs of the reference itself: tests/test_ref_exec.py (CPU), tests/test_gpu_parity.py.
3. `oracle_<block>_<precision>.npz`: outputs of the NumPy oracle (oracle/cloudsc2_numpy.py)
   for NL, TL and AD on the same blocks (first 16 columns), incl. the AD with the TL's
   predicates (not a reference behaviour), so that a drift of the oracle itself is caught
   by tests/test_oracle.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gt4py-dwarf-p-cloudsc2-tl-ad_b200"))

REF_DATA = "/root/reference/data"


def golden_from_reference():
    from cloudsc2_b200.h5lite import File

    for precision in ("double", "single"):
        src = os.path.join(REF_DATA, f"reference_{precision}.h5")
        if not os.path.exists(src):
            print(f"skip {src} (not present)")
            continue
        f = File(src)
        np.savez_compressed(os.path.join(HERE, f"reference_{precision}.npz"), **{k: f[k] for k in f.keys()})
        print("wrote", f"reference_{precision}.npz", sorted(f.keys()))


def reference_fixtures():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H  # noqa: E402
    from oracle import ref_run  # noqa: E402

    if not ref_run.available():
        print("skip ref_*.npz: /root/reference/src not present")
        return
    for name, (block, dtype, ncol, flags) in H.REF_FIXTURES.items():
        out = H.pipeline_run_all(ref_run, block, dtype, ncol, **flags)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name + ".npz", len(out), "arrays")


def oracle_fixtures(ncol=16):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_run_all  # noqa: E402

    for block in ("base", "cold"):
        for precision, dtype in (("double", np.float64), ("single", np.float32)):
            out = oracle_run_all(block, dtype, ncol)
            np.savez_compressed(os.path.join(HERE, f"oracle_{block}_{precision}.npz"), **out)
            print("wrote", f"oracle_{block}_{precision}.npz", len(out), "arrays")


if __name__ == "__main__":
    golden_from_reference()
    reference_fixtures()
    if "--no-oracle" not in sys.argv:
        oracle_fixtures()
