"""Host-resident states through the GPU: blocked, stream-pipelined saturation + CLOUDSC2-NL.

The IFS (and the CLOUDSC dwarfs) keep their fields on the host in NPROMA blocks: `nblocks` independent
groups of `block_cols` columns.  `NonlinearHostPipeline` takes such blocks from pinned host memory, and per
block does ONE host->device copy of the packed inputs, the `Saturation` + `Cloudsc2NL` component calls, and
ONE device->host copy of the packed outputs, on three CUDA streams with a ring of device slots, so that the
H2D copy of block b+1, the kernels of block b and the D2H copy of block b-1 overlap (PCIe is full duplex;
the kernels are ~2 % of a block's transfer time, so throughput is bound by the slower PCIe direction).

This is the path `bench.py` reports as `e2e`: host buffers in, host buffers out, through the component API.
"""
from __future__ import annotations

from datetime import timedelta
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import iox
from .framework.config import GridConfig, GT4PyConfig
from .framework.grid import ComputationalGrid, I, J, K
from .framework.storage import Field, column_stride, torch_dtype
from .physics._names import NL_DIAGNOSTICS, NL_INPUTS, NL_TENDENCIES
from .physics.common.saturation import Saturation
from .physics.nonlinear.microphysics import Cloudsc2NL

# packed order of a block's input / output planes (f_qsat is produced on the device, not copied in)
IN_NAMES = tuple(f"f_{n}" for n in NL_INPUTS if n != "qsat")
# f_covptot last: without the precipitation-evaporation branch (LEVAPLS2 / LDRAIN1D off, the default) the stencil writes
# an identical 0 to it (nonlinear/_stencils/cloudsc2.py:138), so its plane is not copied back (10 % of the D2H traffic):
# the host plane is zero from its allocation on
OUT_NAMES = (tuple(f"f_{n}" for n in NL_TENDENCIES) + tuple(f"f_{n}" for n in NL_DIAGNOSTICS if n != "covptot")
             + ("f_covptot",))


class _Slot:
    """Device buffers of one in-flight block: packed inputs, qsat, packed outputs, as Fields."""

    def __init__(self, grid: ComputationalGrid, cfg: GT4PyConfig, device: torch.device) -> None:
        dt = torch_dtype(cfg.dtypes.float)
        rows, stride = grid.nz + 1, column_stride(grid.nx)
        self.inp = torch.zeros((len(IN_NAMES), rows, stride), dtype=dt, device=device)
        self.out = torch.zeros((len(OUT_NAMES), rows, stride), dtype=dt, device=device)
        self.qsat = torch.zeros((rows, stride), dtype=dt, device=device)
        dims = lambda name: (I, J, K - 1 / 2) if name in ("f_aph", "f_fhpsl", "f_fhpsn", "f_fplsl", "f_fplsn") else (I, J, K)  # noqa: E731
        self.state: Dict[str, Field] = {n: Field(self.inp[i], grid.nx, dims(n), name=n) for i, n in enumerate(IN_NAMES)}
        self.state["f_qsat"] = Field(self.qsat, grid.nx, (I, J, K), name="f_qsat")
        fields = {n: Field(self.out[i], grid.nx, dims(n), name=n) for i, n in enumerate(OUT_NAMES)}
        self.tends = {f"f_{n}": fields[f"f_{n}"] for n in NL_TENDENCIES}
        self.diags = {f"f_{n}": fields[f"f_{n}"] for n in NL_DIAGNOSTICS}
        self.sat_out = {"f_qsat": self.state["f_qsat"]}
        self.ev_in = torch.cuda.Event()
        self.ev_done = torch.cuda.Event()
        self.ev_free = torch.cuda.Event()


class NonlinearHostPipeline:
    def __init__(self, block_cols: int, nz: int, params: Dict[str, object] | None = None, *, gt4py_config: GT4PyConfig,
                 timestep: timedelta = iox.DEFAULT_TIMESTEP, nslots: int = 3, eta: np.ndarray | None = None) -> None:
        if not torch.cuda.is_available():
            from ._lib import CUDAExtensionError

            raise CUDAExtensionError("NonlinearHostPipeline needs a CUDA device; there is no CPU fallback")
        p = params or iox.ifs_defaults()
        self.cfg, self.dt = gt4py_config, timestep
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.grid = ComputationalGrid(GridConfig(nx=block_cols, ny=1, nz=nz))
        self.stride = column_stride(block_cols)
        self.saturation = Saturation(self.grid, 1, True, p["yoethf"], p["yomcst"], gt4py_config=gt4py_config)
        self.cloudsc2_nl = Cloudsc2NL(self.grid, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"],
                                      p["yrphnc"], gt4py_config=gt4py_config)
        self.covptot_is_zero = not bool(getattr(p["yrphnc"], "LEVAPLS2", False))  # ldrain1d is False here
        self.nout_copied = len(OUT_NAMES) - 1 if self.covptot_is_zero else len(OUT_NAMES)
        self.slots = [_Slot(self.grid, gt4py_config, self.device) for _ in range(nslots)]
        self.s_in, self.s_run, self.s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        self.eta = None if eta is None else torch.as_tensor(np.asarray(eta), dtype=torch_dtype(gt4py_config.dtypes.float))
        self.launches = 0

    # ---- host-side block buffers ---------------------------------------------------------
    def alloc_host_block(self) -> Dict[str, torch.Tensor]:
        dt = torch_dtype(self.cfg.dtypes.float)
        rows = self.grid.nz + 1
        return {"in": torch.zeros((len(IN_NAMES), rows, self.stride), dtype=dt).pin_memory(),
                "out": torch.zeros((len(OUT_NAMES), rows, self.stride), dtype=dt).pin_memory()}

    def pack_inputs(self, block: Dict[str, torch.Tensor], arrays: Dict[str, np.ndarray]) -> None:
        """Fill a host block from `(K, IJ)` arrays (full-level arrays may omit the padding level)."""
        for i, name in enumerate(IN_NAMES):
            a = torch.as_tensor(np.ascontiguousarray(arrays[name]), dtype=block["in"].dtype)
            block["in"][i, : a.shape[0], : a.shape[1]].copy_(a)

    @staticmethod
    def unpack_outputs(block: Dict[str, torch.Tensor], ncol: int) -> Dict[str, np.ndarray]:
        return {name: block["out"][i, :, :ncol].numpy().copy() for i, name in enumerate(OUT_NAMES)}

    @property
    def h2d_bytes_per_block(self) -> int:
        s = self.slots[0]
        return s.inp.numel() * s.inp.element_size()

    @property
    def d2h_bytes_per_block(self) -> int:
        s = self.slots[0]
        return s.out[: self.nout_copied].numel() * s.out.element_size()

    # ---- the pipeline ------------------------------------------------------------------------
    def run(self, blocks: Sequence[Dict[str, torch.Tensor]], eta: torch.Tensor | None = None, sync: bool = True) -> None:
        """Process all host blocks.  With `sync=True` (default) the host waits for the last device-to-host copy, so on
        return every block's "out" buffer holds its 10 NL outputs and may be read.  With `sync=False` the call only
        orders the CURRENT stream after the pipeline's streams (stream-ordered, asynchronous): the caller must
        synchronise (e.g. `torch.cuda.current_stream().synchronize()`) before touching the pinned buffers."""
        eta = eta if eta is not None else self.eta
        if eta is None:
            raise ValueError("eta (the K-field of EtaLevels, from global column 0) is required")
        eta_field = Field(eta, None, (K,), name="f_eta")
        cur = torch.cuda.current_stream()
        for st in (self.s_in, self.s_run, self.s_out):
            st.wait_stream(cur)
        for b, blk in enumerate(blocks):
            slot = self.slots[b % len(self.slots)]
            with torch.cuda.stream(self.s_in):
                if b >= len(self.slots):
                    self.s_in.wait_event(slot.ev_free)  # the slot's previous outputs have left the device
                slot.inp.copy_(blk["in"], non_blocking=True)
                slot.ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(slot.ev_in)
                slot.state["f_eta"] = eta_field
                self.saturation(slot.state, out=slot.sat_out)
                self.cloudsc2_nl(slot.state, self.dt, out_tendencies=slot.tends, out_diagnostics=slot.diags)
                slot.ev_done.record(self.s_run)
                self.launches += 2
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(slot.ev_done)
                blk["out"][: self.nout_copied].copy_(slot.out[: self.nout_copied], non_blocking=True)
                slot.ev_free.record(self.s_out)
        for st in (self.s_in, self.s_run, self.s_out):
            cur.wait_stream(st)
        if sync:
            self.s_out.synchronize()
