"""CPU oracle for the CLOUDSC2 NL/TL/AD hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`gt4py-dwarf-p-cloudsc2-tl-ad_b200/`) may import this package; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` do, and there only as the checker / the CPU arm.

Parity status: **pinned to outputs of the reference's own source run here.**  The GT4Py runtime
(`gt4py`, `ifs_physics_common`, `h5py`) cannot be imported in this image, but the reference's arithmetic lives in plain
Python gtscript files; `oracle/gtscript_exec.py` executes those files UNMODIFIED where they lie under /root/reference
(an interpreter with GT4Py cartesian semantics under NumPy), `oracle/ref_run.py` takes the externals and argument maps
from the reference's component classes, and
  * `tests/test_ref_exec.py` compares this restatement with them on both synthetic blocks, fp64 and fp32, saturation /
    NL / state_increment / perturbed_state / TL / AD and the flag variants: bit-identical in fp64, <= 4e-6 in fp32;
  * `tests/golden/ref_*.npz` (written by `tests/golden/make_golden.py` from the reference source) carry the pin to boxes
    without /root/reference: the oracle, the host twin and the CUDA path are compared with them directly.
What stays unavailable is `data/input.h5` (not shipped), so the golden NL outputs `data/reference_*.h5` pin the oracle
through their invariants only (`tests/test_golden.py`; the point-wise comparison is wired and runs the moment a matching
`input.h5` is supplied).
"""
