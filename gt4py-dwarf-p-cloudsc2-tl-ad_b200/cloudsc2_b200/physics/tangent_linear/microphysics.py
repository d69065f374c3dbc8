"""`Cloudsc2TL` (reference: physics/tangent_linear/microphysics.py:46-242)."""
from __future__ import annotations

from functools import cached_property
from itertools import repeat

import numpy as np

from ...framework.components import ImplicitTendencyComponent
from ...framework.grid import I, J, K
from ...framework.storage import gt_zeros, managed_temporary_storage
from .._names import FULL, NL_DIAGNOSTICS, NL_INPUTS, NL_TENDENCIES, props
from ..nonlinear.microphysics import physics_externals


class Cloudsc2TL(ImplicitTendencyComponent):
    def __init__(self, computational_grid, lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params,
                 yrephli_params, yrncl_params, yrphnc_params, *, enable_checks=True, gt4py_config):
        super().__init__(computational_grid, enable_checks=enable_checks, gt4py_config=gt4py_config)
        nk = self.computational_grid.grids[I, J, K].shape[2]
        self.klevel = gt_zeros(self.computational_grid, (K,), gt4py_config=self.gt4py_config, dtype_name="int")
        self.klevel[:] = self.klevel.new_tensor(np.arange(0, nk + 1))
        externals = physics_externals(lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params, yrephli_params,
                                      yrncl_params, yrphnc_params, NLEV=nk)
        self.cloudsc2 = self.compile_stencil("cloudsc2_tl", externals)

    @cached_property
    def input_grid_properties(self):
        out = {"f_eta": props((K,), "")}
        for n, (d, u) in NL_INPUTS.items():
            out[f"f_{n}"] = props(d, u)
            out[f"f_{n}_i"] = props(d, u)
        return out

    @cached_property
    def tendency_grid_properties(self):
        out = {}
        for n, u in NL_TENDENCIES.items():
            out[f"f_{n}"] = props(FULL, u)
            out[f"f_{n}_i"] = props(FULL, u)
        return out

    @cached_property
    def diagnostic_grid_properties(self):
        out = {}
        for n, (d, u) in NL_DIAGNOSTICS.items():
            out[f"f_{n}"] = props(d, u)
            out[f"f_{n}_i"] = props(d, u)
        return out

    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies):
        with managed_temporary_storage(
            self.computational_grid, *repeat(((I, J), "float"), 9), gt4py_config=self.gt4py_config
        ) as (aph_s, aph_s_i, rfl, rfl_i, sfl, sfl_i, covptot, covptot_i, trpaus):
            kwargs = {}
            for sfx in ("", "_i"):
                kwargs.update({f"in_{n}{sfx}": state[f"f_{n}{sfx}"] for n in NL_INPUTS})
                kwargs.update({f"out_{n}{sfx}": out_diagnostics[f"f_{n}{sfx}"] for n in NL_DIAGNOSTICS})
                kwargs.update({f"out_tnd_{n}{sfx}": out_tendencies[f"f_{n}{sfx}"] for n in NL_TENDENCIES})
            self.cloudsc2(
                **kwargs, in_eta=state["f_eta"],
                tmp_aph_s=aph_s, tmp_aph_s_i=aph_s_i, tmp_covptot=covptot, tmp_covptot_i=covptot_i,
                tmp_klevel=self.klevel, tmp_rfl=rfl, tmp_rfl_i=rfl_i, tmp_sfl=sfl, tmp_sfl_i=sfl_i, tmp_trpaus=trpaus,
                dt=self.gt4py_config.dtypes.float(timestep.total_seconds()), origin=(0, 0, 0),
                domain=self.computational_grid.grids[I, J, K - 1 / 2].shape,
                validate_args=self.gt4py_config.validate_args, exec_info=self.gt4py_config.exec_info,
            )


class IncrementedCloudsc2TL(Cloudsc2TL):
    """`StateIncrement(factor, ignore_supsat)` followed by `Cloudsc2TL`, fused: the perturbation of every input is
    factor * input, formed inside the sweep, so the state needs no `f_*_i` fields.  Not in the reference; opt-in fast path
    of both validation harnesses (tangent_linear/validation.py:158-162, adjoint/validation.py:136-140); equal to the two
    separate components up to FMA contraction (~1e-14 field-scaled, within the 1e-12 parity tolerance)."""

    def __init__(self, computational_grid, factor, ignore_supsat, lphylin, ldrain1d, yoethf_params, yomcst_params,
                 yrecldp_params, yrephli_params, yrncl_params, yrphnc_params, *, enable_checks=True, gt4py_config):
        super().__init__(computational_grid, lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params,
                         yrephli_params, yrncl_params, yrphnc_params, enable_checks=enable_checks, gt4py_config=gt4py_config)
        self.f = gt4py_config.dtypes.float(factor)
        externals = dict(self.cloudsc2.externals, IGNORE_SUPSAT=bool(ignore_supsat))
        self.cloudsc2_increment = self.compile_stencil("cloudsc2_tl_increment", externals)
        # optional fp64 device tensor [nx]: if set, every call also leaves SUM_k SUM_fields (TL output)^2 per column in
        # it -- the first inner product of the symmetry test (adjoint/validation.py:167-181), from the sweep itself
        self.norm1 = None

    @cached_property
    def input_grid_properties(self):
        out = {"f_eta": props((K,), "")}
        for n, (d, u) in NL_INPUTS.items():
            out[f"f_{n}"] = props(d, u)
        return out

    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies):
        kwargs = {f"in_{n}": state[f"f_{n}"] for n in NL_INPUTS}
        for sfx in ("", "_i"):
            kwargs.update({f"out_{n}{sfx}": out_diagnostics[f"f_{n}{sfx}"] for n in NL_DIAGNOSTICS})
            kwargs.update({f"out_tnd_{n}{sfx}": out_tendencies[f"f_{n}{sfx}"] for n in NL_TENDENCIES})
        self.cloudsc2_increment(
            **kwargs, in_eta=state["f_eta"], f=self.f, norm1=self.norm1,
            dt=self.gt4py_config.dtypes.float(timestep.total_seconds()),
            origin=(0, 0, 0), domain=self.computational_grid.grids[I, J, K - 1 / 2].shape,
            validate_args=self.gt4py_config.validate_args, exec_info=self.gt4py_config.exec_info,
        )
