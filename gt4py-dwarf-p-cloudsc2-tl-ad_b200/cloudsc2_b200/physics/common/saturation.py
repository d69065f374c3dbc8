"""`Saturation` diagnostic component (reference: physics/common/saturation.py:33-76)."""
from __future__ import annotations

from functools import cached_property

from ...framework.components import DiagnosticComponent
from ...framework.grid import I, J, K


class Saturation(DiagnosticComponent):
    """qsat(ap, t): in f_ap, f_t -> out f_qsat, on the full levels only."""

    def __init__(self, computational_grid, kflag, lphylin, yoethf_params, yomcst_params, *, enable_checks=True, gt4py_config):
        super().__init__(computational_grid, enable_checks=enable_checks, gt4py_config=gt4py_config)
        externals = {"KFLAG": kflag, "LPHYLIN": lphylin, "QMAX": 0.5}
        externals.update(yoethf_params.dict())
        externals.update(yomcst_params.dict())
        self.saturation = self.compile_stencil("saturation", externals)

    @cached_property
    def input_grid_properties(self):
        return {"f_ap": {"grid_dims": (I, J, K), "units": "Pa"}, "f_t": {"grid_dims": (I, J, K), "units": "K"}}

    @cached_property
    def diagnostic_grid_properties(self):
        return {"f_qsat": {"grid_dims": (I, J, K), "units": "g g^-1"}}

    def array_call(self, state, out):
        self.saturation(
            in_ap=state["f_ap"], in_t=state["f_t"], out_qsat=out["f_qsat"], origin=(0, 0, 0),
            domain=self.computational_grid.grids[I, J, K].shape,
            validate_args=self.gt4py_config.validate_args, exec_info=self.gt4py_config.exec_info,
        )
