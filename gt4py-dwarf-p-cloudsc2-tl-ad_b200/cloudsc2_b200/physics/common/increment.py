"""`StateIncrement` and `PerturbedState` (reference: physics/common/increment.py:32-261)."""
from __future__ import annotations

from functools import cached_property

from ...framework.components import DiagnosticComponent
from ...framework.grid import I, J, K
from .._names import STATE_FIELDS, props


class StateIncrement(DiagnosticComponent):
    """x_i = f * x for the 16 state fields; `ignore_supsat` zeroes f_supsat_i."""

    def __init__(self, computational_grid, factor, ignore_supsat=False, *, enable_checks=True, gt4py_config):
        super().__init__(computational_grid, enable_checks=enable_checks, gt4py_config=gt4py_config)
        self.f = gt4py_config.dtypes.float(factor)
        self.increment = self.compile_stencil("state_increment", externals={"IGNORE_SUPSAT": ignore_supsat})

    @cached_property
    def input_grid_properties(self):
        return {f"f_{n}": props(d, u) for n, (d, u) in STATE_FIELDS.items()}

    @cached_property
    def diagnostic_grid_properties(self):
        return {f"f_{n}_i": props(d, u) for n, (d, u) in STATE_FIELDS.items()}

    def array_call(self, state, out):
        kwargs = {f"in_{n}": state[f"f_{n}"] for n in STATE_FIELDS}
        kwargs.update({f"out_{n}_i": out[f"f_{n}_i"] for n in STATE_FIELDS})
        self.increment(
            **kwargs, f=self.f, origin=(0, 0, 0), domain=self.computational_grid.grids[I, J, K - 1 / 2].shape,
            validate_args=self.gt4py_config.validate_args, exec_info=self.gt4py_config.exec_info,
        )


class PerturbedState(DiagnosticComponent):
    """x_p = x + f * x_i for the 16 state fields (qsat is perturbed as an independent input)."""

    def __init__(self, computational_grid, factor, *, enable_checks=True, gt4py_config):
        super().__init__(computational_grid, enable_checks=enable_checks, gt4py_config=gt4py_config)
        self.f = gt4py_config.dtypes.float(factor)
        self.perturbed_state = self.compile_stencil("perturbed_state")

    @cached_property
    def input_grid_properties(self):
        out = {f"f_{n}": props(d, u) for n, (d, u) in STATE_FIELDS.items()}
        out.update({f"f_{n}_i": props(d, u) for n, (d, u) in STATE_FIELDS.items()})
        return out

    @cached_property
    def diagnostic_grid_properties(self):
        return {f"f_{n}": props(d, u) for n, (d, u) in STATE_FIELDS.items()}

    def array_call(self, state, out):
        kwargs = {f"in_{n}": state[f"f_{n}"] for n in STATE_FIELDS}
        kwargs.update({f"in_{n}_i": state[f"f_{n}_i"] for n in STATE_FIELDS})
        kwargs.update({f"out_{n}": out[f"f_{n}"] for n in STATE_FIELDS})
        self.perturbed_state(
            **kwargs, f=self.f, origin=(0, 0, 0), domain=self.computational_grid.grids[I, J, K - 1 / 2].shape,
            validate_args=self.gt4py_config.validate_args, exec_info=self.gt4py_config.exec_info,
        )
