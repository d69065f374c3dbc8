"""CPU oracle for the CLOUDSC2 NL/TL/AD hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`gt4py-dwarf-p-cloudsc2-tl-ad_b200/`) may import this package; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` do, and there only as the checker / the CPU arm.

Parity status ("Oracle"): the reference (GT4Py + ifs_physics_common
+ h5py) cannot be imported in this image and `data/input.h5` is not shipped, so
this restatement is **parity unpinned** against a live run of the reference.  It
is pinned by (a) invariants of the reference's golden outputs
(`tests/golden/reference_*.npz`), (b) the reference's own Taylor test (TL vs NL)
and (c) the reference's own symmetry test (AD vs TL), and it becomes pinned
point-wise the moment a matching `input.h5` is supplied (`tests/test_golden.py`).
"""
