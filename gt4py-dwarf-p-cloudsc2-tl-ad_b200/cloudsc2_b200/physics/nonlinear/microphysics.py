"""`Cloudsc2NL` (reference: physics/nonlinear/microphysics.py:43-172)."""
from __future__ import annotations

from functools import cached_property
from itertools import repeat

from ...framework.components import ImplicitTendencyComponent
from ...framework.grid import I, J, K
from ...framework.storage import managed_temporary_storage
from .._names import FULL, NL_DIAGNOSTICS, NL_INPUTS, NL_TENDENCIES, props


def physics_externals(lphylin, ldrain1d, *param_sets, **extra):
    """The externals dict the reference components assemble (nonlinear/microphysics.py:61-78)."""
    externals = {}
    for params in param_sets:
        externals.update(params.dict())
    externals.update({"ICALL": 0, "LPHYLIN": lphylin, "LDRAIN1D": ldrain1d, "ZEPS1": 1e-12, "ZEPS2": 1e-10,
                      "ZQMAX": 0.5, "ZSCAL": 0.9})
    externals.update(extra)
    return externals


class Cloudsc2NL(ImplicitTendencyComponent):
    def __init__(self, computational_grid, lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params,
                 yrephli_params, yrphnc_params, *, enable_checks=True, gt4py_config):
        super().__init__(computational_grid, enable_checks=enable_checks, gt4py_config=gt4py_config)
        externals = physics_externals(lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params, yrephli_params,
                                      yrphnc_params)
        self.cloudsc2 = self.compile_stencil("cloudsc2_nl", externals)

    @cached_property
    def input_grid_properties(self):
        out = {f"f_{n}": props(d, u) for n, (d, u) in NL_INPUTS.items()}
        out["f_eta"] = props((K,), "")
        return out

    @cached_property
    def tendency_grid_properties(self):
        return {f"f_{n}": props(FULL, u) for n, u in NL_TENDENCIES.items()}

    @cached_property
    def diagnostic_grid_properties(self):
        return {f"f_{n}": props(d, u) for n, (d, u) in NL_DIAGNOSTICS.items()}

    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies):
        with managed_temporary_storage(
            self.computational_grid, *repeat(((I, J), "float"), 5), gt4py_config=self.gt4py_config
        ) as (aph_s, rfl, sfl, covptot, trpaus):
            kwargs = {f"in_{n}": state[f"f_{n}"] for n in NL_INPUTS}
            kwargs.update({f"out_{n}": out_diagnostics[f"f_{n}"] for n in NL_DIAGNOSTICS})
            kwargs.update({f"out_tnd_{n}": out_tendencies[f"f_{n}"] for n in NL_TENDENCIES})
            self.cloudsc2(
                **kwargs, in_eta=state["f_eta"],
                tmp_aph_s=aph_s, tmp_covptot=covptot, tmp_rfl=rfl, tmp_sfl=sfl, tmp_trpaus=trpaus,
                dt=self.gt4py_config.dtypes.float(timestep.total_seconds()), origin=(0, 0, 0),
                domain=self.computational_grid.grids[I, J, K - 1 / 2].shape,
                validate_args=self.gt4py_config.validate_args, exec_info=self.gt4py_config.exec_info,
            )


class PerturbedCloudsc2NL(Cloudsc2NL):
    """`PerturbedState(factor)` followed by `Cloudsc2NL`, fused: the NL of the state x + factor * x_i, reading x
    and x_i directly (32 input fields, no 16-field intermediate state).  Not in the reference; it is the opt-in
    fast path of the Taylor test's inner loop (tangent_linear/validation.py:167-176) and returns bit-identical
    results to the two separate components."""

    def __init__(self, computational_grid, factor, lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params,
                 yrephli_params, yrphnc_params, *, enable_checks=True, gt4py_config):
        super().__init__(computational_grid, lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params,
                         yrephli_params, yrphnc_params, enable_checks=enable_checks, gt4py_config=gt4py_config)
        self.f = gt4py_config.dtypes.float(factor)
        self.cloudsc2_perturbed = self.compile_stencil("cloudsc2_nl_perturbed", self.cloudsc2.externals)

    @cached_property
    def input_grid_properties(self):
        out = {"f_eta": props((K,), "")}
        for n, (d, u) in NL_INPUTS.items():
            out[f"f_{n}"] = props(d, u)
            out[f"f_{n}_i"] = props(d, u)
        return out

    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies):
        kwargs = {f"in_{n}": state[f"f_{n}"] for n in NL_INPUTS}
        kwargs.update({f"in_{n}_i": state[f"f_{n}_i"] for n in NL_INPUTS})
        kwargs.update({f"out_{n}": out_diagnostics[f"f_{n}"] for n in NL_DIAGNOSTICS})
        kwargs.update({f"out_tnd_{n}": out_tendencies[f"f_{n}"] for n in NL_TENDENCIES})
        self.cloudsc2_perturbed(
            **kwargs, in_eta=state["f_eta"], f=self.f, dt=self.gt4py_config.dtypes.float(timestep.total_seconds()),
            origin=(0, 0, 0), domain=self.computational_grid.grids[I, J, K - 1 / 2].shape,
            validate_args=self.gt4py_config.validate_args, exec_info=self.gt4py_config.exec_info,
        )


class TaylorCloudsc2NL(Cloudsc2NL):
    """`StateIncrement(factor1)` -> `PerturbedState(factor2)` -> `Cloudsc2NL` -> field sums of the Taylor test, as ONE
    sweep: the perturbed NL outputs are compared with the unperturbed ones level by level and only the ten sums
    SUM(F_p - F_nl) leave the kernel (tangent_linear/validation.py:158-176,252-261).  Not in the reference; the opt-in
    fast path `TaylorTest(fused="sums")`.  Call: `comp(state, timestep, tends_nl, diags_nl, sums)` with the unperturbed
    NL outputs and an fp64 device tensor [10][2]; adds to sums[:, 0] in the field order of TaylorTest.get_norm."""

    def __init__(self, computational_grid, factor1, factor2, lphylin, ldrain1d, yoethf_params, yomcst_params,
                 yrecldp_params, yrephli_params, yrphnc_params, *, ignore_supsat=False, enable_checks=True, gt4py_config):
        super().__init__(computational_grid, lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params,
                         yrephli_params, yrphnc_params, enable_checks=enable_checks, gt4py_config=gt4py_config)
        self.f1 = gt4py_config.dtypes.float(factor1)
        self.f2 = gt4py_config.dtypes.float(factor2)
        externals = dict(self.cloudsc2.externals, IGNORE_SUPSAT=bool(ignore_supsat))
        self.cloudsc2_taylor = self.compile_stencil("cloudsc2_nl_taylor_sums", externals)

    def __call__(self, state, timestep, tends_nl, diags_nl, sums):  # noqa: D102 - not a tendency component call
        raw = lambda fld: getattr(fld, "data", fld)  # noqa: E731 - Field -> logical (nx, 1, nz+1) view, like the base class
        kwargs = {f"in_{n}": raw(state[f"f_{n}"]) for n in NL_INPUTS}
        kwargs.update({f"out_{n}": raw(diags_nl[f"f_{n}"]) for n in NL_DIAGNOSTICS})
        kwargs.update({f"out_tnd_{n}": raw(tends_nl[f"f_{n}"]) for n in NL_TENDENCIES})
        self.cloudsc2_taylor(
            **kwargs, in_eta=raw(state["f_eta"]), f1=self.f1, f2=self.f2, sums=sums,
            dt=self.gt4py_config.dtypes.float(timestep.total_seconds()), origin=(0, 0, 0),
            domain=self.computational_grid.grids[I, J, K - 1 / 2].shape,
            validate_args=self.gt4py_config.validate_args, exec_info=self.gt4py_config.exec_info,
        )
        return sums
