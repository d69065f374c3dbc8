"""Regenerates the fixtures in this directory.  Run from the repo root in the authoring
container (needs /root/reference for part 1; part 2 only needs the oracle):

    python tests/golden/make_golden.py

1. `reference_double.npz` / `reference_single.npz`: the datasets of the reference's golden
   NL outputs `data/reference_{double,single}.h5` (read with cloudsc2_b200.h5lite, stored
   loss-free with np.savez_compressed).  These are DATA published by the reference for
   validation (drivers/run_nonlinear.py:139-147), not source code.
2. `oracle_<block>_<precision>.npz`: outputs of the NumPy oracle (oracle/cloudsc2_numpy.py)
   for NL, TL and AD on the seeded synthetic blocks of cloudsc2_b200.synthetic (first 16
   columns), so that a drift of the oracle itself is caught by tests/test_oracle.py.
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
    if "--no-oracle" not in sys.argv:
        oracle_fixtures()
