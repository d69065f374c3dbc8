"""`SymmetryTest` (reference: physics/adjoint/validation.py:43-231).

Same orchestration as the reference; the per-column inner products <TL x, TL x> and
<x, AD TL x> are computed on the device in fp64 (`reductions.symmetry_norms`), only the scalar
max_i norm3 is all-reduced (MAX) when the columns are sharded."""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch

from ... import distributed
from ...framework.grid import I, J, K
from ...reductions import SymmetryResidual, symmetry_norms
from ..common.increment import StateIncrement
from ..common.saturation import Saturation
from ..tangent_linear.microphysics import Cloudsc2TL, IncrementedCloudsc2TL
from .microphysics import Cloudsc2AD

TL_TENDS = ("f_t_i", "f_q_i", "f_ql_i", "f_qi_i")
TL_DIAGS = ("f_clc_i", "f_fhpsl_i", "f_fhpsn_i", "f_fplsl_i", "f_fplsn_i", "f_covptot_i")
AD_TENDS = ("f_cml_t_i", "f_cml_q_i", "f_cml_ql_i", "f_cml_qi_i")
AD_DIAGS = ("f_ap_i", "f_aph_i", "f_t_i", "f_q_i", "f_qsat_i", "f_ql_i", "f_qi_i", "f_lu_i", "f_lude_i", "f_mfd_i",
            "f_mfu_i", "f_supsat_i")


class SymmetryTest:
    def __init__(self, computational_grid, factor, kflag, lphylin, ldrain1d, yoethf_params, yomcst_params,
                 yrecldp_params, yrephli_params, yrncl_params, yrphnc_params, *, enable_checks=True, gt4py_config,
                 ad_predicates=None, ad_trajectory=None, fused=False):
        """`fused=True`: the TL sweep forms its perturbation factor * state itself (IncrementedCloudsc2TL) and the two
        inner products come out of the TL and AD sweeps (per-column fp64 sums accumulated level by level), so `state_i`
        is not materialised and no separate reduction runs: 3 kernels + the seed reset instead of 6.  Same residuals up
        to round-off.  With LEVAPLS2 / LDRAIN1D only the TL part is fused."""
        self.f = factor
        self.fused = fused
        self.fused_norms = bool(fused) and not (ldrain1d or getattr(yrphnc_params, "LEVAPLS2", False))
        kw = dict(enable_checks=enable_checks, gt4py_config=gt4py_config)
        self.saturation = Saturation(computational_grid, kflag, lphylin, yoethf_params, yomcst_params, **kw)
        self.cloudsc2_tl = Cloudsc2TL(computational_grid, lphylin, ldrain1d, yoethf_params, yomcst_params,
                                      yrecldp_params, yrephli_params, yrncl_params, yrphnc_params, **kw)
        self.cloudsc2_ad = Cloudsc2AD(computational_grid, lphylin, ldrain1d, yoethf_params, yomcst_params,
                                      yrecldp_params, yrephli_params, yrncl_params, yrphnc_params,
                                      ad_predicates=ad_predicates, ad_trajectory=ad_trajectory,
                                      symmetry_increment=(factor, True) if self.fused_norms else None, **kw)
        self.cloudsc2_tl_inc = IncrementedCloudsc2TL(computational_grid, factor, True, lphylin, ldrain1d, yoethf_params,
                                                     yomcst_params, yrecldp_params, yrephli_params, yrncl_params,
                                                     yrphnc_params, **kw) if fused else None
        self.state_increment = StateIncrement(computational_grid, factor, ignore_supsat=True, **kw)
        self.diags_sat: Dict[str, Any] = {}
        self.state_i: Dict[str, Any] = {}
        self.tends_tl: Dict[str, Any] = {}
        self.diags_tl: Dict[str, Any] = {}
        self.tends_ad: Dict[str, Any] = {}
        self.diags_ad: Dict[str, Any] = {}
        self.norm3_max: Optional[float] = None
        self.norm3: Optional[torch.Tensor] = None
        self._norm1: Optional[torch.Tensor] = None  # per-column inner products written by the fused sweeps
        self._norm2: Optional[torch.Tensor] = None
        self._residual = SymmetryResidual()

    def __call__(self, state, timestep, enable_validation: bool = True, verbose: bool = True) -> Optional[bool]:
        self.diags_sat = self.saturation(state, out=self.diags_sat)
        state.update(self.diags_sat)
        fused_norms = self.fused_norms and enable_validation
        if fused_norms:
            nx = self.saturation.computational_grid.grids[I, J, K].shape[0]
            dev = state["f_ap"].buffer.device
            if self._norm1 is None or self._norm1.numel() != nx or self._norm1.device != dev:
                self._norm1 = torch.zeros(nx, dtype=torch.float64, device=dev)
                self._norm2 = torch.zeros(nx, dtype=torch.float64, device=dev)
        self.cloudsc2_ad.norm2 = self._norm2 if fused_norms else None
        if self.cloudsc2_tl_inc is not None:
            self.cloudsc2_tl_inc.norm1 = self._norm1 if fused_norms else None
        # the increment is only read by the unfused TL sweep and by the unfused second inner product
        if not self.fused or (enable_validation and not self.fused_norms):
            self.state_i = self.state_increment(state, out=self.state_i)
            state.update(self.state_i)
        tl = self.cloudsc2_tl_inc if self.fused else self.cloudsc2_tl
        self.tends_tl, self.diags_tl = tl(state, timestep, out_tendencies=self.tends_tl, out_diagnostics=self.diags_tl)
        norm1 = None
        if enable_validation:
            norm1 = self._norm1 if fused_norms else self.get_norm1(self.tends_tl, self.diags_tl)

        self.add_tendencies_to_state(state, self.tends_tl)
        state.update(self.diags_tl)
        self.tends_ad, self.diags_ad = self.cloudsc2_ad(
            state, timestep, out_tendencies=self.tends_ad, out_diagnostics=self.diags_ad
        )
        if not enable_validation:
            return None

        norm2 = self._norm2 if fused_norms else self.get_norm2(self.state_i, self.tends_ad, self.diags_ad)
        self.norm1, self.norm2 = norm1, norm2  # per-column <TL x, TL x> and <x, AD TL x>
        eps = float(np.finfo(self.saturation.gt4py_config.dtypes.float).eps)
        # norm3 and its maximum in one device epilogue (cs2_symmetry_residual); an empty shard contributes -inf
        self.norm3, nmax = self._residual(norm1, norm2, eps)
        distributed.allreduce_max_(nmax)
        self.norm3_max = float(nmax.item())
        passed = self.norm3_max < 1e4
        if verbose:
            print("The symmetry test passed. HOORAY!" if passed else "The symmetry test failed.")
            print(f"The maximum error is {self.norm3_max:.10e} times the machine epsilon.")
        return passed

    @staticmethod
    def get_norm1(tends_tl, diags_tl) -> torch.Tensor:
        """<TL x, TL x> per column (:167-181)."""
        flds = [tends_tl[n] for n in TL_TENDS] + [diags_tl[n] for n in TL_DIAGS]
        return symmetry_norms(flds, flds)

    @staticmethod
    def get_norm2(state_i, tends_ad, diags_ad) -> torch.Tensor:
        """<x, AD TL x> per column (:183-215)."""
        a = [state_i["f_tnd_" + n[2:]] for n in AD_TENDS] + [state_i[n] for n in AD_DIAGS]
        b = [tends_ad[n] for n in AD_TENDS] + [diags_ad[n] for n in AD_DIAGS]
        return symmetry_norms(a, b)

    @staticmethod
    def add_tendencies_to_state(state, tends_tl) -> None:
        """(:222-231)"""
        for x in ("t", "q", "ql", "qi"):
            state[f"f_tnd_{x}"] = tends_tl[f"f_{x}"]
            state[f"f_tnd_{x}_i"] = tends_tl[f"f_{x}_i"]
