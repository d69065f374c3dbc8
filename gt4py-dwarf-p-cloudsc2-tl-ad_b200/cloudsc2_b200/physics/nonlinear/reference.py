"""Golden-output loaders (reference: physics/nonlinear/reference.py:28-55).

Note the reference maps `TENDENCY_LOC_Q` to the key `f_qv` although `Cloudsc2NL` emits `f_q`
(so its driver never compares that field); here both keys are provided."""
from __future__ import annotations

from typing import Any, Dict

from ...setup import REFERENCE_TIME, HDF5GridOperator


def get_reference_tendencies(hdf5_grid_operator: HDF5GridOperator) -> Dict[str, Any]:
    g = hdf5_grid_operator
    tends = {
        "f_qi": g.get_field("TENDENCY_LOC_CLD", 1, False, "g g^-1 s^-1", "f_qi"),
        "f_ql": g.get_field("TENDENCY_LOC_CLD", 0, False, "g g^-1 s^-1", "f_ql"),
        "f_qv": g.get_field("TENDENCY_LOC_Q", None, False, "g g^-1 s^-1", "f_qv"),
        "f_t": g.get_field("TENDENCY_LOC_T", None, False, "K s^-1", "f_t"),
    }
    tends["f_q"] = tends["f_qv"]
    tends["time"] = REFERENCE_TIME
    return tends


def get_reference_diagnostics(hdf5_grid_operator: HDF5GridOperator) -> Dict[str, Any]:
    g = hdf5_grid_operator
    diags = {
        "f_clc": g.get_field("PCLC", None, False, "1", "f_clc"),
        "f_covptot": g.get_field("PCOVPTOT", None, False, "1", "f_covptot"),
        "f_fhpsl": g.get_field("PFHPSL", None, True, "J m^-2 s^-1", "f_fhpsl"),
        "f_fhpsn": g.get_field("PFHPSN", None, True, "J m^-2 s^-1", "f_fhpsn"),
        "f_fplsl": g.get_field("PFPLSL", None, True, "kg m^-2 s^-1", "f_fplsl"),
        "f_fplsn": g.get_field("PFPLSN", None, True, "kg m^-2 s^-1", "f_fplsn"),
    }
    diags["time"] = REFERENCE_TIME
    return diags
