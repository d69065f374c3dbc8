"""`EtaLevels` (reference: physics/common/diagnostics.py:28-45): eta[k] = ap[0,0,k] / aph[0,0,nz]
from column 0 ONLY -- a K-field shared by all columns.  Under column sharding every rank must use
the eta of GLOBAL column 0 (see cloudsc2_b200.distributed.broadcast_eta)."""
from __future__ import annotations

from functools import cached_property

from ...framework.components import DiagnosticComponent
from ...framework.grid import I, J, K


class EtaLevels(DiagnosticComponent):
    @cached_property
    def input_grid_properties(self):
        return {"f_ap": {"grid_dims": (I, J, K), "units": "Pa"}, "f_aph": {"grid_dims": (I, J, K - 1 / 2), "units": "Pa"}}

    @cached_property
    def diagnostic_grid_properties(self):
        return {"f_eta": {"grid_dims": (K,), "units": ""}}

    def array_call(self, state, out):
        nz = self.computational_grid.grids[I, J, K].shape[2]
        if self.computational_grid.nx < 1:
            raise ValueError("EtaLevels needs column 0 of the (global) domain; this grid has no columns")
        # one device->host transfer of column 0 instead of the reference's nz scalar reads
        ap0 = state["f_ap"][0, 0, :nz].detach().cpu()
        aph_s = state["f_aph"][0, 0, nz].detach().cpu()
        out["f_eta"][:nz] = ap0 / aph_s
