"""Seeded synthetic CLOUDSC2 inputs (the reference's `data/input.h5` is not shipped).

Produces the 16 state fields of `cloudsc2_gt4py/setup.py:47-70` for a base block of
`KLON` columns x `nz` levels in the HDF5 `(K, IJ)` layout (level-major, column-fastest),
and tiles it to any number of columns the way `--num-cols` does in the reference drivers
(column i <- column i mod KLON), so every replicated block must yield bit-identical output.

The block is built to exercise both sides of every data-dependent predicate of the stencils
(SURVEY.md section 9.4): warm and cold surfaces (RTT, RTT+2 crossings, melting snow), an
inversion above a per-column tropopause in 0.1 < eta < 0.4, clear / partly cloudy / overcast
levels, convective detrainment (lude >= RLMIN with lu >= ZEPS2), both signs of the
subsidence term, supersaturated points (q > qsat) and very dry points.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

KLON = 100
KLEV = 137

FIELDS_FULL = (
    "f_ap", "f_lu", "f_lude", "f_mfd", "f_mfu", "f_q", "f_qi", "f_ql", "f_supsat", "f_t",
    "f_tnd_cml_q", "f_tnd_cml_qi", "f_tnd_cml_ql", "f_tnd_cml_t",
)
FIELDS_HALF = ("f_aph",)


def _qsat_guess(p, t):
    """Rough Tetens saturation specific humidity used only to place the humidity profile."""
    es_l = 611.21 * np.exp(17.502 * (t - 273.16) / (t - 32.19))
    es_i = 611.21 * np.exp(22.587 * (t - 273.16) / (t + 0.7))
    w = np.clip((t - 250.16) / 23.0, 0.0, 1.0) ** 2
    es = w * es_l + (1 - w) * es_i
    qs = np.minimum(0.622 * es / np.maximum(p, 1.0), 0.5)
    return qs / (1.0 - 0.608 * qs)


def base_block(nz: int = KLEV, ncol: int = KLON, seed: int = 0, dtype=np.float64) -> Dict[str, np.ndarray]:
    """One block of `ncol` synthetic columns; arrays are `[nz+1, ncol]` (full-level fields
    carry a zero padding level at index nz, exactly like the reference storages)."""
    rng = np.random.default_rng(seed)
    k = np.arange(nz + 1, dtype=np.float64)
    s = k / nz
    # hybrid-like half levels: aph = A + B * ps, aph[0] = 0, aph[nz] = ps
    etah = 0.55 * s**3.2 + 0.45 * s**1.35
    etah[0] = 0.0
    etah[nz] = 1.0
    b = etah**1.6
    a = (etah - b) * 1.0e5
    ps = rng.uniform(0.95e5, 1.05e5, size=ncol)
    ps[0] = 1.0e5
    aph = a[:, None] + b[:, None] * ps[None, :]
    ap = np.zeros_like(aph)
    ap[:nz] = 0.5 * (aph[:-1] + aph[1:])
    eta = ap[:nz] / aph[nz][None, :]

    # temperature: troposphere with lapse rate, inversion above a per-column tropopause
    ts = rng.uniform(248.0, 306.0, size=ncol)
    ts[: ncol // 10] = rng.uniform(272.5, 277.5, size=ncol // 10)  # hug RTT / RTT+2
    eta_tp = rng.uniform(0.12, 0.36, size=ncol)
    eta_tp[-max(ncol // 12, 1):] = 0.05  # no inversion inside 0.1 < eta < 0.4: the tropopause rule must fall back to 0.1
    kappa = rng.uniform(0.17, 0.21, size=ncol)
    t_trop = ts[None, :] * np.maximum(eta, 1e-6) ** kappa[None, :]
    t_tp = ts * eta_tp**kappa
    t_strat = t_tp[None, :] * (1.0 + 0.06 * np.log(eta_tp[None, :] / np.maximum(eta, 1e-6)))
    t = np.where(eta >= eta_tp[None, :], t_trop, np.minimum(t_strat, 290.0))
    t += rng.normal(0.0, 0.15, size=t.shape)
    # a few boundary-layer inversions
    inv = rng.random(ncol) < 0.2
    t[nz - 6 : nz - 2, inv] += 2.5

    # humidity: smooth RH with saturated slabs
    rh = rng.uniform(0.1, 0.75, size=ncol)[None, :] + 0.25 * np.sin(
        6.0 * eta + rng.uniform(0, 6.28, size=ncol)[None, :]
    )
    n_slab = 3
    for _ in range(n_slab):
        c = rng.uniform(0.3, 0.98, size=ncol)
        w = rng.uniform(0.02, 0.12, size=ncol)
        amp = rng.uniform(0.0, 0.75, size=ncol)
        rh += amp[None, :] * np.exp(-(((eta - c[None, :]) / w[None, :]) ** 2))
    rh = np.clip(rh, 0.02, 1.12)
    rh[eta < eta_tp[None, :] * 0.8] *= 0.15
    qs = _qsat_guess(ap[:nz], t)
    q = np.minimum(rh * qs, 0.025)
    q = np.where(eta < eta_tp[None, :] * 0.8, np.minimum(q, 6.0e-6), q)

    cloudy = rh > 0.8
    ql = np.where(cloudy & (t > 250.0), rng.uniform(0.0, 3e-4, size=t.shape), 0.0)
    qi = np.where(cloudy & (t < 273.16), rng.uniform(0.0, 3e-4, size=t.shape), 0.0)
    sparse = rng.random(t.shape) < 0.03
    ql = np.where(sparse & (eta > 0.3), rng.uniform(0, 1e-4, size=t.shape), ql)

    # convection
    conv = rng.random(ncol) < 0.45
    ktop = rng.uniform(0.25, 0.6, size=ncol)
    in_conv = conv[None, :] & (eta > ktop[None, :]) & (eta < 0.93)
    lude = np.where(in_conv & (rng.random(t.shape) < 0.5), 10.0 ** rng.uniform(-7.0, -4.0, size=t.shape), 0.0)
    lu = np.where(in_conv, rng.uniform(0.0, 1e-3, size=t.shape), 0.0)
    lu = np.where(rng.random(t.shape) < 0.1, 0.0, lu)
    mfu = np.where(in_conv, rng.uniform(0.0, 0.3, size=t.shape), 0.0)
    mfd = np.where(in_conv & (eta > 0.6), -rng.uniform(0.0, 0.1, size=t.shape), 0.0)
    supsat = np.where(rng.random(t.shape) < 0.05, rng.uniform(0.0, 1e-5, size=t.shape), 0.0)

    tnd_t = rng.normal(0.0, 1.0e-4, size=t.shape)
    tnd_q = rng.normal(0.0, 1.0e-8, size=t.shape) * (q > 1e-6)
    tnd_ql = np.where(ql > 0, rng.normal(0.0, 1.0e-9, size=t.shape), 0.0)
    tnd_qi = np.where(qi > 0, rng.normal(0.0, 1.0e-9, size=t.shape), 0.0)

    def full(x):
        out = np.zeros((nz + 1, ncol), dtype=dtype)
        out[:nz] = x
        return out

    return {
        "f_aph": aph.astype(dtype),
        "f_ap": full(ap[:nz]),
        "f_t": full(t),
        "f_q": full(q),
        "f_ql": full(ql),
        "f_qi": full(qi),
        "f_lude": full(lude),
        "f_lu": full(lu),
        "f_mfu": full(mfu),
        "f_mfd": full(mfd),
        "f_supsat": full(supsat),
        "f_tnd_cml_t": full(tnd_t),
        "f_tnd_cml_q": full(tnd_q),
        "f_tnd_cml_ql": full(tnd_ql),
        "f_tnd_cml_qi": full(tnd_qi),
    }


def cold_block(nz: int = KLEV, ncol: int = KLON, seed: int = 1, dtype=np.float64) -> Dict[str, np.ndarray]:
    """Like the shipped (all-cold) input of the reference: no column reaches RTT."""
    blk = base_block(nz, ncol, seed, np.float64)
    shift = np.maximum(blk["f_t"][:nz].max(axis=0) - 268.0, 0.0)
    blk["f_t"][:nz] -= shift[None, :]
    blk["f_q"][:nz] *= _qsat_guess(blk["f_ap"][:nz], blk["f_t"][:nz]) / _qsat_guess(
        blk["f_ap"][:nz], blk["f_t"][:nz] + shift[None, :]
    )
    blk["f_ql"][:nz] = np.where(blk["f_t"][:nz] > 250.0, blk["f_ql"][:nz], 0.0)
    return {k: v.astype(dtype) for k, v in blk.items()}


def tile(block: Dict[str, np.ndarray], ncol: int) -> Dict[str, np.ndarray]:
    """Replicate a block to `ncol` columns: column i <- column i mod block_ncol."""
    out = {}
    for name, arr in block.items():
        reps = -(-ncol // arr.shape[1])
        out[name] = np.ascontiguousarray(np.tile(arr, (1, reps))[:, :ncol])
    return out


# ------------------------------------------------------------------------------------------
# `input.h5` / `reference_*.h5` look-alikes (the reference's own data/input.h5 is not shipped)
# ------------------------------------------------------------------------------------------
def input_h5_datasets(block: Dict[str, np.ndarray], params: Dict[str, object], timestep_s: float = 3600.0,
                      dtype=np.float64) -> Dict[str, np.ndarray]:
    """The datasets `setup.get_state` (reference setup.py:48-65) and `iox.HDF5Operator` (iox.py:212-244) read, under the
    reference's names: 2-D fields `(K, IJ)` (full-level fields with KLEV rows, PAPH with KLEV + 1), the 5-species arrays
    `PCLV` / `TENDENCY_CML_CLD` `(5, K, IJ)` (index 0 = liquid, 1 = ice), `KLON`, `KLEV`, `PTSPHY`, and every scalar of the
    parameter models (prefix `YRECLDP_` / `YREPHLI_` where the reference uses one)."""
    nz = block["f_ap"].shape[0] - 1
    klon = block["f_ap"].shape[1]
    full = lambda a: np.ascontiguousarray(a[:nz].astype(dtype))  # noqa: E731
    d: Dict[str, np.ndarray] = {
        "KLON": np.array([klon], dtype=np.int32), "KLEV": np.array([nz], dtype=np.int32),
        "PTSPHY": np.array([timestep_s], dtype=dtype),
        "PA": np.zeros((nz, klon), dtype=dtype), "PAP": full(block["f_ap"]), "PAPH": block["f_aph"].astype(dtype),
        "PLU": full(block["f_lu"]), "PLUDE": full(block["f_lude"]), "PMFD": full(block["f_mfd"]), "PMFU": full(block["f_mfu"]),
        "PQ": full(block["f_q"]), "PSUPSAT": full(block["f_supsat"]), "PT": full(block["f_t"]),
        "TENDENCY_CML_Q": full(block["f_tnd_cml_q"]), "TENDENCY_CML_T": full(block["f_tnd_cml_t"]),
    }
    clv = np.zeros((5, nz, klon), dtype=dtype)
    clv[0], clv[1] = full(block["f_ql"]), full(block["f_qi"])
    cld = np.zeros((5, nz, klon), dtype=dtype)
    cld[0], cld[1] = full(block["f_tnd_cml_ql"]), full(block["f_tnd_cml_qi"])
    d["PCLV"], d["TENDENCY_CML_CLD"] = clv, cld
    prefixes = {"yrecldp": "YRECLDP_", "yrephli": "YREPHLI_"}
    for group, model in params.items():
        for name, value in model.dict().items():
            key = prefixes.get(group, "") + name
            if isinstance(value, bool):
                d[key] = np.array([int(value)], dtype=np.int32)
            elif isinstance(value, int):
                d[key] = np.array([value], dtype=np.int32)
            else:
                d[key] = np.array([value], dtype=np.float64)
    return d


def write_input_h5(filename: str, block: str = "base", params: Dict[str, object] = None, timestep_s: float = 3600.0,
                   seed: int = 0, dtype=np.float64) -> None:
    from . import h5lite, iox

    blk = base_block(seed=seed) if block == "base" else cold_block(seed=seed + 1)
    h5lite.write_file(filename, input_h5_datasets(blk, params or iox.ifs_defaults(), timestep_s, dtype))


def write_reference_h5(filename: str, tendencies: Dict[str, np.ndarray], diagnostics: Dict[str, np.ndarray]) -> None:
    """NL outputs under the dataset names of the reference's golden files data/reference_*.h5 (nonlinear/reference.py:28-55):
    full-level fields with KLEV rows, fluxes with KLEV + 1, `TENDENCY_LOC_CLD` `(5, K, IJ)`."""
    from . import h5lite

    nz = tendencies["f_t"].shape[0] - 1
    klon = tendencies["f_t"].shape[1]
    dt = tendencies["f_t"].dtype
    cld = np.zeros((5, nz, klon), dtype=dt)
    cld[0], cld[1] = tendencies["f_ql"][:nz], tendencies["f_qi"][:nz]
    h5lite.write_file(filename, {
        "KLON": np.array([klon], dtype=np.int32), "KLEV": np.array([nz], dtype=np.int32),
        "PCLC": diagnostics["f_clc"][:nz], "PCOVPTOT": diagnostics["f_covptot"][:nz],
        "PFHPSL": diagnostics["f_fhpsl"], "PFHPSN": diagnostics["f_fhpsn"], "PFPLSL": diagnostics["f_fplsl"],
        "PFPLSN": diagnostics["f_fplsn"], "TENDENCY_LOC_CLD": cld, "TENDENCY_LOC_Q": tendencies["f_q"][:nz],
        "TENDENCY_LOC_T": tendencies["f_t"][:nz],
    })
