"""Field-name tables shared by the components: (state key, units) per stencil argument.
The names are the drop-in contract of the reference (`*_grid_properties` of its components)."""
from __future__ import annotations

from ..framework.grid import I, J, K

FULL = (I, J, K)
HALF = (I, J, K - 1 / 2)

# NL inputs: stencil argument suffix -> (grid dims, units)   (nonlinear/microphysics.py:81-101)
NL_INPUTS = {
    "ap": (FULL, "Pa"), "aph": (HALF, "Pa"), "lu": (FULL, "g g^-1"), "lude": (FULL, "kg m^-3 s^-1"),
    "mfd": (FULL, "kg m^-2 s^-1"), "mfu": (FULL, "kg m^-2 s^-1"), "q": (FULL, "g g^-1"), "qi": (FULL, "g g^-1"),
    "ql": (FULL, "g g^-1"), "qsat": (FULL, "g g^-1"), "supsat": (FULL, "g g^-1"), "t": (FULL, "K"),
    "tnd_cml_q": (FULL, "g g^-1 s^-1"), "tnd_cml_qi": (FULL, "g g^-1 s^-1"), "tnd_cml_ql": (FULL, "g g^-1 s^-1"),
    "tnd_cml_t": (FULL, "K s^-1"),
}
# NL tendencies: output name -> units   (nonlinear/microphysics.py:103-110)
NL_TENDENCIES = {"q": "g g^-1 s^-1", "qi": "g g^-1 s^-1", "ql": "g g^-1 s^-1", "t": "K s^-1"}
# NL diagnostics: output name -> (grid dims, units)   (nonlinear/microphysics.py:112-121)
NL_DIAGNOSTICS = {
    "clc": (FULL, ""), "covptot": (FULL, ""), "fhpsl": (HALF, "J m^-2 s^-1"), "fhpsn": (HALF, "J m^-2 s^-1"),
    "fplsl": (HALF, "Kg m^-2 s^-1"), "fplsn": (HALF, "Kg m^-2 s^-1"),
}
# the 16 fields handled by StateIncrement / PerturbedState   (common/increment.py:52-69)
STATE_FIELDS = {
    "aph": (HALF, "Pa"), "ap": (FULL, "Pa"), "q": (FULL, "g g^-1"), "qsat": (FULL, "g g^-1"), "t": (FULL, "K"),
    "ql": (FULL, "g g^-1"), "qi": (FULL, "g g^-1"), "lude": (FULL, "kg m^-3 s^-1"), "lu": (FULL, "g g^-1"),
    "mfu": (FULL, "kg m^-2 s^-1"), "mfd": (FULL, "kg m^-2 s^-1"), "tnd_cml_t": (FULL, "K s^-1"),
    "tnd_cml_q": (FULL, "K s^-1"), "tnd_cml_ql": (FULL, "K s^-1"), "tnd_cml_qi": (FULL, "K s^-1"),
    "supsat": (FULL, "g g^-1"),
}


def props(dims, units):
    return {"grid_dims": dims, "units": units}
