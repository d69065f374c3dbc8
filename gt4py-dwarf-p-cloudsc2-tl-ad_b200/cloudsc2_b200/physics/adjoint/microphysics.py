"""`Cloudsc2AD` (reference: physics/adjoint/microphysics.py:46-238).

Side effects kept from the reference stencil: the adjoint seeds found in `state`
(`f_tnd_{t,q,ql,qi}_i`, `f_clc_i`, `f_covptot_i`, `f_f{h,p}ps{l,n}_i`) are consumed, i.e. zeroed
in place (adjoint/_stencils/cloudsc2.py:482-484,506-542,650,714,920,972-984).

`ad_predicates`: "reference" (default: what a drop-in user of the reference gets) restates the
reference AD stencil literally -- second freezing test on the pre-adjustment temperature
(adjoint/_stencils/cloudsc2.py:427,577), backward first-freezing test on the post-adjustment
temperature (:729); parity-tested against the reference's own stencil source
(tests/golden/ref_*.npz).  "tl" (opt-in, also `CS2_AD_PREDICATES=tl`) evaluates every branch
predicate exactly like the TL sweep on the same trajectory, which makes the component the exact
adjoint of `Cloudsc2TL` on ANY input.  The two are identical on all-cold inputs such as the
reference's `input.h5` and differ only where a level crosses RTT during the saturation adjustment;
there the literal mode -- in the reference and here alike -- fails the reference's own symmetry
test ("AD predicates")."""
from __future__ import annotations

import os
from functools import cached_property
from itertools import repeat

import numpy as np

from ...framework.components import ImplicitTendencyComponent
from ...framework.grid import I, J, K
from ...framework.storage import gt_zeros, managed_temporary_storage
from .._names import FULL, HALF, NL_DIAGNOSTICS, NL_INPUTS, NL_TENDENCIES, STATE_FIELDS, props
from ..nonlinear.microphysics import physics_externals

# adjoint seeds: state key -> (stencil argument, dims, units)   (adjoint/microphysics.py:106-120)
AD_SEEDS = {
    "f_tnd_t_i": ("in_tnd_t_i", FULL, "K s^-1"), "f_tnd_q_i": ("in_tnd_q_i", FULL, "K s^-1"),
    "f_tnd_ql_i": ("in_tnd_ql_i", FULL, "K s^-1"), "f_tnd_qi_i": ("in_tnd_qi_i", FULL, "K s^-1"),
    "f_clc_i": ("in_clc_i", FULL, ""), "f_covptot_i": ("in_covptot_i", FULL, ""),
    "f_fhpsl_i": ("in_fhpsl_i", HALF, "J m^-2 s^-1"), "f_fhpsn_i": ("in_fhpsn_i", HALF, "J m^-2 s^-1"),
    "f_fplsl_i": ("in_fplsl_i", HALF, "kg m^-2 s^-1"), "f_fplsn_i": ("in_fplsn_i", HALF, "kg m^-2 s^-1"),
}
# adjoint outputs in the diagnostics dict   (adjoint/microphysics.py:137-150)
AD_DIAG_ADJOINTS = ("aph", "ap", "q", "qsat", "t", "ql", "qi", "lude", "lu", "mfu", "mfd", "supsat")


class Cloudsc2AD(ImplicitTendencyComponent):
    def __init__(self, computational_grid, lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params,
                 yrephli_params, yrncl_params, yrphnc_params, *, enable_checks=True, gt4py_config,
                 ad_predicates=None, ad_trajectory=None, symmetry_increment=None):
        """`symmetry_increment=(factor, ignore_supsat)` + the attribute `norm2` (fp64 device tensor [nx]): every call also
        leaves SUM_k SUM_fields (factor * input) * (adjoint output) per column in `norm2` -- the second inner product of
        the symmetry test (adjoint/validation.py:183-215), from the backward sweep itself (cs2_ad_norm2)."""
        super().__init__(computational_grid, enable_checks=enable_checks, gt4py_config=gt4py_config)
        self.symmetry_increment = symmetry_increment
        self.norm2 = None
        nk = self.computational_grid.grids[I, J, K].shape[2]
        self.klevel = gt_zeros(self.computational_grid, (K,), gt4py_config=self.gt4py_config, dtype_name="int")
        self.klevel[:] = self.klevel.new_tensor(np.arange(0, nk + 1))
        ad_predicates = ad_predicates or os.environ.get("CS2_AD_PREDICATES", "reference")
        if ad_predicates not in ("tl", "reference"):
            raise ValueError("ad_predicates must be 'tl' or 'reference'")
        self.ad_predicates = ad_predicates
        # "checkpoint": the forward sweep stores the 9 transcendental results per point to an HBM workspace (72 B/point in
        # fp64) and the backward sweep replays them; "recompute": the backward sweep recomputes each level's trajectory
        # from the inputs (no workspace); "auto" (default): whichever measured faster on B200 for the call's size -- with
        # the lockstep exponentials and the in-kernel seed reset that is recompute everywhere (1 % at 65 536 columns,
        # 10 % and 10 GB at 1 M).
        self.ad_trajectory = ad_trajectory or os.environ.get("CS2_AD_TRAJECTORY", "auto")
        externals = physics_externals(lphylin, ldrain1d, yoethf_params, yomcst_params, yrecldp_params, yrephli_params,
                                      yrncl_params, yrphnc_params, NLEV=nk,
                                      AD_TL_PREDICATES=(ad_predicates == "tl"), AD_TRAJECTORY=self.ad_trajectory,
                                      IGNORE_SUPSAT=bool(symmetry_increment[1]) if symmetry_increment else False)
        self.cloudsc2 = self.compile_stencil("cloudsc2_ad", externals)

    @cached_property
    def input_grid_properties(self):
        out = {"f_eta": props((K,), "")}
        out.update({f"f_{n}": props(d, u) for n, (d, u) in NL_INPUTS.items()})
        out.update({key: props(d, u) for key, (_, d, u) in AD_SEEDS.items()})
        return out

    @cached_property
    def tendency_grid_properties(self):
        out = {}
        for n, u in NL_TENDENCIES.items():
            out[f"f_{n}"] = props(FULL, u)
            out[f"f_cml_{n}_i"] = props(FULL, u)
        return out

    @cached_property
    def diagnostic_grid_properties(self):
        out = {f"f_{n}_i": props(*STATE_FIELDS[n]) for n in AD_DIAG_ADJOINTS}
        out.update({f"f_{n}": props(d, u) for n, (d, u) in NL_DIAGNOSTICS.items()})
        return out

    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies):
        with managed_temporary_storage(
            self.computational_grid, *repeat(((I, J), "float"), 8), gt4py_config=self.gt4py_config
        ) as (aph_s, aph_s_i, covptotp, rfln, rfln_i, sfln, sfln_i, trpaus):
            kwargs = {f"in_{n}": state[f"f_{n}"] for n in NL_INPUTS}
            kwargs.update({arg: state[key] for key, (arg, _, _) in AD_SEEDS.items()})
            kwargs.update({f"out_{n}": out_diagnostics[f"f_{n}"] for n in NL_DIAGNOSTICS})
            kwargs.update({f"out_tnd_{n}": out_tendencies[f"f_{n}"] for n in NL_TENDENCIES})
            kwargs.update({f"out_{n}_i": out_diagnostics[f"f_{n}_i"] for n in AD_DIAG_ADJOINTS})
            kwargs.update({f"out_tnd_cml_{n}_i": out_tendencies[f"f_cml_{n}_i"] for n in NL_TENDENCIES})
            if self.norm2 is not None and self.symmetry_increment is not None:
                kwargs.update(norm2=self.norm2, increment_factor=self.gt4py_config.dtypes.float(self.symmetry_increment[0]))
            self.cloudsc2(
                **kwargs, in_eta=state["f_eta"],
                tmp_aph_s=aph_s, tmp_aph_s_i=aph_s_i, tmp_covptotp=covptotp, tmp_klevel=self.klevel,
                tmp_rfln=rfln, tmp_rfln_i=rfln_i, tmp_sfln=sfln, tmp_sfln_i=sfln_i, tmp_trpaus=trpaus,
                dt=self.gt4py_config.dtypes.float(timestep.total_seconds()), origin=(0, 0, 0),
                domain=self.computational_grid.grids[I, J, K - 1 / 2].shape,
                validate_args=self.gt4py_config.validate_args, exec_info=self.gt4py_config.exec_info,
            )
