"""NumPy restatement of the CLOUDSC2 stencils of `cloudsc2_gt4py` (TEST INFRASTRUCTURE).

Execution model = GT4Py's `numpy` backend (the reference default, `drivers/config.py:45`):
vectorised over columns, Python loop over levels, every branch evaluated under masks.

Layout: every field is a 2-D array `[nz+1, nx]` (level-major, column-fastest, exactly
the `(K, IJ)` layout of the reference's HDF5 files, `setup.py:28-35`); full-level fields
carry one zero padding level at index `nz` like the reference's `(nx, 1, nz+1)` storages.
`eta` is a 1-D array of length >= nz.

All physical constants come in through the externals dict `P` (python floats / bools) with
the names the reference's stencils import from `__externals__`.

Reference files restated here (relative to /root/reference/src/cloudsc2_gt4py/physics):
  common/_stencils/fcttre.py:22-57            -> foealfa, foealfcu, foeewm, foeewmcu
  common/_stencils/saturation.py:23-42        -> saturation
  common/_stencils/state_increment.py:22-80   -> state_increment
  common/_stencils/perturbed_state.py:22-91   -> perturbed_state
  common/diagnostics.py:42-45                 -> eta_levels
  nonlinear/_stencils/cuadjtqs.py:22-68       -> cuadjtqs_nl
  nonlinear/_stencils/cloudsc2.py:24-399      -> cloudsc2_nl
  tangent_linear/_stencils/cuadjtqs.py:22-84  -> cuadjtqs_tl
  tangent_linear/_stencils/cloudsc2.py:23-774 -> cloudsc2_tl
  adjoint/_stencils/cuadjtqs.py:22-158        -> cuadjtqs_ad
  adjoint/_stencils/cloudsc2.py:24-996        -> cloudsc2_ad
  tangent_linear/validation.py:219-261        -> taylor_norm
  adjoint/validation.py:167-215               -> symmetry_norms

Parity: pinned to outputs of the reference's own stencil sources executed here by oracle/gtscript_exec.py (bit-identical in
fp64 on every field of every case, tests/test_ref_exec.py; fixtures tests/golden/ref_*.npz) -- see oracle/__init__.py.
"""
from __future__ import annotations

import sys

import numpy as np

W = np.where

STATE_FIELDS = (
    "f_aph", "f_ap", "f_q", "f_qsat", "f_t", "f_ql", "f_qi", "f_lude", "f_lu", "f_mfu", "f_mfd",
    "f_tnd_cml_t", "f_tnd_cml_q", "f_tnd_cml_ql", "f_tnd_cml_qi", "f_supsat",
)  # common/increment.py:52-69


def _errstate():
    return np.errstate(divide="ignore", invalid="ignore", over="ignore", under="ignore")


# ----------------------------------------------------------------------------------------
# thermodynamic functions -- common/_stencils/fcttre.py
# ----------------------------------------------------------------------------------------
def foealfa(t, P):  # fcttre.py:22-27
    return np.minimum(
        1.0, ((np.maximum(P["RTICE"], np.minimum(P["RTWAT"], t)) - P["RTICE"]) * P["RTWAT_RTICE_R"]) ** 2.0
    )


def foealfcu(t, P):  # fcttre.py:30-35
    return np.minimum(
        1.0,
        ((np.maximum(P["RTICECU"], np.minimum(P["RTWAT"], t)) - P["RTICECU"]) * P["RTWAT_RTICECU_R"]) ** 2.0,
    )


def foeewm(t, P):  # fcttre.py:38-46
    a = foealfa(t, P)
    return P["R2ES"] * (
        a * np.exp(P["R3LES"] * (t - P["RTT"]) / (t - P["R4LES"]))
        + (1.0 - a) * (np.exp(P["R3IES"] * (t - P["RTT"]) / (t - P["R4IES"])))
    )


def foeewmcu(t, P):  # fcttre.py:49-57
    a = foealfcu(t, P)
    return P["R2ES"] * (
        a * np.exp(P["R3LES"] * (t - P["RTT"]) / (t - P["R4LES"]))
        + (1.0 - a) * (np.exp(P["R3IES"] * (t - P["RTT"]) / (t - P["R4IES"])))
    )


# ----------------------------------------------------------------------------------------
# saturation -- common/_stencils/saturation.py:23-42 ; domain = full levels only
# (common/saturation.py:73)
# ----------------------------------------------------------------------------------------
def saturation(ap, t, P, nz=None, out=None):
    """ap, t: [nz+1, nx]; returns qsat [nz+1, nx] (level nz left untouched / zero)."""
    nz = ap.shape[0] - 1 if nz is None else nz
    qsat = np.zeros_like(ap) if out is None else out
    a, tt = ap[:nz], t[:nz]
    QMAX = 0.5  # common/saturation.py:51
    with _errstate():
        if P["LPHYLIN"]:
            alfa = foealfa(tt, P)
            foeewl = P["R2ES"] * np.exp(P["R3LES"] * (tt - P["RTT"]) / (tt - P["R4LES"]))
            foeewi = P["R2ES"] * np.exp(P["R3IES"] * (tt - P["RTT"]) / (tt - P["R4IES"]))
            foeew = alfa * foeewl + (1.0 - alfa) * foeewi
            qs = np.minimum(foeew / a, QMAX)
        else:
            ew = foeewmcu(tt, P) if P.get("KFLAG", 1) == 1 else foeewm(tt, P)
            qs = np.minimum(ew / a, QMAX)
        qsat[:nz] = qs / (1.0 - P["RETV"] * qs)
    return qsat


def eta_levels(ap, aph, nz=None):
    """common/diagnostics.py:42-45 -- eta from column 0 only."""
    nz = ap.shape[0] - 1 if nz is None else nz
    eta = np.zeros(nz + 1, dtype=ap.dtype)
    eta[:nz] = ap[:nz, 0] / aph[nz, 0]
    return eta


def state_increment(state, f, ignore_supsat=False):
    """common/_stencils/state_increment.py:59-80 (all nz+1 levels, increment.py:129)."""
    out = {}
    for name in STATE_FIELDS:
        x = state[name]
        fx = x.dtype.type(f)
        if name == "f_supsat" and ignore_supsat:
            out[name + "_i"] = np.zeros_like(x)
        else:
            out[name + "_i"] = fx * x
    return out


def perturbed_state(state, f):
    """common/_stencils/perturbed_state.py:75-91: x + f * x_i for the 16 state fields."""
    out = {}
    for name in STATE_FIELDS:
        x = state[name]
        fx = x.dtype.type(f)
        out[name] = x + fx * state[name + "_i"]
    return out


# ----------------------------------------------------------------------------------------
# cuadjtqs -- nonlinear/_stencils/cuadjtqs.py
# ----------------------------------------------------------------------------------------
def _cuadjtqs_phase(t, P):
    warm = t > P["RTT"]
    z3es = W(warm, P["R3LES"], P["R3IES"]).astype(t.dtype)
    z4es = W(warm, P["R4LES"], P["R4IES"]).astype(t.dtype)
    z5alcp = W(warm, P["R5ALVCP"], P["R5ALSCP"]).astype(t.dtype)
    zaldcp = W(warm, P["RALVDCP"], P["RALSDCP"]).astype(t.dtype)
    return z3es, z4es, z5alcp, zaldcp


def _cuadjtqs_nl_0(ap, t, q, z3es, z4es, z5alcp, zaldcp, P):  # cuadjtqs.py:22-35
    foeew = P["R2ES"] * np.exp(z3es * (t - P["RTT"]) / (t - z4es))
    qsat = np.minimum(foeew / ap, P["ZQMAX"])
    cor = 1.0 / (1.0 - P["RETV"] * qsat)
    qsat = qsat * cor
    z2s = z5alcp / (t - z4es) ** 2.0
    cond = (q - qsat) / (1.0 + qsat * cor * z2s)
    t = t + zaldcp * cond
    q = q - cond
    return t, q


def cuadjtqs_nl(ap, t, q, P):  # cuadjtqs.py:38-68 (ICALL == 0 only)
    z3es, z4es, z5alcp, zaldcp = _cuadjtqs_phase(t, P)
    t, q = _cuadjtqs_nl_0(ap, t, q, z3es, z4es, z5alcp, zaldcp, P)
    t, q = _cuadjtqs_nl_0(ap, t, q, z3es, z4es, z5alcp, zaldcp, P)
    return t, q


def _crh2(eta_k, trpaus, dtype):
    """nonlinear/_stencils/cloudsc2.py:165-186 (identical in TL :232-253 and AD :202-223)."""
    rh1 = 1.0
    rh2 = 0.35 + 0.14 * ((trpaus - 0.25) / 0.15) ** 2.0 + 0.04 * np.minimum(trpaus - 0.25, 0.0) / 0.15
    rh3 = 1.0
    deta2 = 0.3
    bound1 = trpaus + deta2
    deta1 = 0.09 + 0.16 * (0.4 - trpaus) / 0.3
    bound2 = 1.0 - deta1
    c_a = rh3 + (rh2 - rh3) * (eta_k - trpaus) / deta2
    c_c = rh1 + (rh2 - rh1) * np.sqrt((1.0 - eta_k) / deta1)
    crh2 = W(eta_k < trpaus, rh3, W(eta_k < bound1, c_a, W(eta_k < bound2, rh2, c_c)))
    return crh2.astype(dtype)


def _trpaus(t3d, eta, nz):
    """nonlinear/_stencils/cloudsc2.py:106-111: last level with 0.1<eta<0.4 and t[k]>t[k+1]."""
    dtype = t3d.dtype
    trpaus = np.full(t3d.shape[1], 0.1, dtype=dtype)
    for k in range(nz - 1):
        m = (eta[k] > 0.1) & (eta[k] < 0.4) & (t3d[k] > t3d[k + 1])
        trpaus = W(m, eta[k], trpaus)
    return trpaus


# ----------------------------------------------------------------------------------------
# cloudsc2_nl -- nonlinear/_stencils/cloudsc2.py:24-399
# ----------------------------------------------------------------------------------------
def cloudsc2_nl(s, dt, P):
    """s: dict with f_ap, f_aph, f_eta, f_lu, f_lude, f_mfd, f_mfu, f_q, f_qi, f_ql, f_qsat,
    f_supsat, f_t, f_tnd_cml_{q,qi,ql,t}.  Returns (tendencies, diagnostics) dicts with the
    names of nonlinear/microphysics.py:103-121."""
    ap, aph, eta = s["f_ap"], s["f_aph"], s["f_eta"]
    dtype = ap.dtype
    nz = ap.shape[0] - 1
    nx = ap.shape[1]
    dt = dtype.type(dt)
    RTT, RG, RETV = P["RTT"], P["RG"], P["RETV"]
    evap_on = bool(P["LEVAPLS2"] or P["LDRAIN1D"])

    z = lambda: np.zeros((nz + 1, nx), dtype=dtype)  # noqa: E731
    o_clc, o_covptot, o_fhpsl, o_fhpsn, o_fplsl, o_fplsn = z(), z(), z(), z(), z(), z()
    o_q, o_qi, o_ql, o_t = z(), z(), z(), z()

    with _errstate():
        # :93-100
        rfl = np.zeros(nx, dtype=dtype)
        sfl = np.zeros(nx, dtype=dtype)
        covptot = np.zeros(nx, dtype=dtype)
        aph_s = aph[nz].copy()
        # :102-104
        t3d = s["f_t"][:nz] + dt * s["f_tnd_cml_t"][:nz]
        # :106-111
        trpaus = _trpaus(t3d, eta, nz)

        fplsl = np.zeros((nz, nx), dtype=dtype)
        fplsn = np.zeros((nz, nx), dtype=dtype)
        for k in range(nz):  # :113-388
            t = t3d[k]
            apk, qsk = ap[k], s["f_qsat"][k]
            q = s["f_q"][k] + dt * s["f_tnd_cml_q"][k] + s["f_supsat"][k]
            ql = s["f_ql"][k] + dt * s["f_tnd_cml_ql"][k]
            qi = s["f_qi"][k] + dt * s["f_tnd_cml_qi"][k]

            ckcodtl = 2.0 * P["RKCONV"] * dt
            ckcodti = 5.0 * P["RKCONV"] * dt
            cons2 = 1.0 / (RG * dt)
            cons3 = P["RLVTT"] / P["RCPD"]
            meltp2 = RTT + 2.0

            scalm = dtype.type(P["ZSCAL"] * max(eta[k] - dtype.type(0.2), dtype.type(P["ZEPS1"])) ** dtype.type(0.2))

            dp = aph[k + 1] - aph[k]
            zz = P["RCPD"] + P["RCPD"] * P["RVTMP2"] * q
            lfdcp = P["RLMLT"] / zz
            lsdcp = P["RLSTT"] / zz
            lvdcp = P["RLVTT"] / zz

            if P["LPHYLIN"] or P["LDRAIN1D"]:  # :141-151
                cold = t < RTT
                fwat = W(cold, 0.545 * (np.tanh(0.17 * (t - P["RLPTRC"])) + 1.0), 1.0).astype(dtype)
                z3es = W(cold, P["R3IES"], P["R3LES"]).astype(dtype)
                z4es = W(cold, P["R4IES"], P["R4LES"]).astype(dtype)
                foeew = P["R2ES"] * np.exp(z3es * (t - RTT) / (t - z4es))
                esdp = np.minimum(foeew / apk, P["ZQMAX"])
            else:  # :152-155
                fwat = foealfa(t, P)
                foeew = foeewm(t, P)
                esdp = foeew / apk
            facw = P["R5LES"] / ((t - P["R4LES"]) ** 2.0)
            faci = P["R5IES"] / ((t - P["R4IES"]) ** 2.0)
            fac = fwat * facw + (1.0 - fwat) * faci
            dqsdtemp = fac * qsk / (1.0 - RETV * esdp)
            corqs = 1.0 + cons3 * dqsdtemp

            qlim = np.minimum(q, qsk)  # :163

            crh2 = _crh2(eta[k], trpaus, dtype)  # :165-186

            qsat = W(t < P["RTICE"], qsk * (1.8 - 0.003 * t), qsk)  # :189-192
            qcrit = crh2 * qsat

            qt = q + ql + qi  # :196-207
            b1 = qt < qcrit
            b2 = ~b1 & (qt >= qsat)
            qpd = qsat - qt
            qcd = qsat - qcrit
            clc3 = 1.0 - np.sqrt(qpd / (qcd - scalm * (qt - qcrit)))
            clc = W(b1, 0.0, W(b2, 1.0, clc3)).astype(dtype)
            qc = W(
                b1, 0.0, W(b2, (1.0 - scalm) * (qsat - qcrit), (scalm * qpd + (1.0 - scalm) * qcd) * (clc3**2.0))
            ).astype(dtype)

            gdp = RG / (aph[k + 1] - aph[k])  # :210-215
            lude = dt * s["f_lude"][k] * gdp
            lu1 = s["f_lu"][k + 1]
            lo1 = (lude >= P["RLMIN"]) & (lu1 >= P["ZEPS2"])
            clc = W(lo1, clc + (1.0 - clc) * (1.0 - np.exp(-lude / lu1)), clc)
            qc = W(lo1, qc + lude, qc)

            rho = apk / (P["RD"] * t)  # :218-224
            rodqsdp = -rho * qsk / (apk - RETV * foeew)
            ldcp = fwat * lvdcp + (1.0 - fwat) * lsdcp
            dtdzmo = RG * (1.0 / P["RCPD"] - ldcp * rodqsdp) / (1.0 + ldcp * dqsdtemp)
            dqsdz = dqsdtemp * dtdzmo - RG * rodqsdp
            dqc = np.minimum(dt * dqsdz * (s["f_mfu"][k] + s["f_mfd"][k]) / rho, qc)
            qc = qc - dqc

            qlwc = qc * fwat  # :227-230
            qiwc = qc * (1.0 - fwat)
            condl = (qlwc - ql) / dt
            condi = (qiwc - qi) / dt

            covptot = np.maximum(covptot, clc)  # :234-235
            covpclr = np.maximum(covptot - clc, 0.0)

            melt = sfl != 0.0  # :238-246
            cons = cons2 * dp / lfdcp
            snmlt = np.minimum(sfl, cons * np.maximum(t - meltp2, 0.0))
            rfln = W(melt, rfl + snmlt, rfl)
            sfln = W(melt, sfl - snmlt, sfl)
            t = W(melt, t - snmlt / cons, t)

            cloudy = clc > P["ZEPS2"]  # :249-272
            lcrit = (1.9 if evap_on else 2.0) * P["RCLCRIT"]
            icrit = 0.0001 if evap_on else 2.0 * P["RCLCRIT"]
            cldl = qlwc / clc
            dl = ckcodtl * (1.0 - np.exp(-((cldl / lcrit) ** 2.0)))
            prr = W(cloudy, qlwc - clc * cldl * np.exp(-dl), 0.0).astype(dtype)
            qlwc = W(cloudy, qlwc - prr, qlwc)
            cldi = qiwc / clc
            di = ckcodti * np.exp(0.025 * (t - RTT)) * (1.0 - np.exp(-((cldi / icrit) ** 2.0)))
            prs = W(cloudy, qiwc - clc * cldi * np.exp(-di), 0.0).astype(dtype)
            qiwc = W(cloudy, qiwc - prs, qiwc)

            dr = cons2 * dp * (prr + prs)  # :275-285
            frz = t < RTT
            rfreeze = W(frz, cons2 * dp * prr, 0.0).astype(dtype)
            fwatr = W(frz, 0.0, 1.0).astype(dtype)
            rfln = rfln + fwatr * dr
            sfln = sfln + (1.0 - fwatr) * dr

            prtot = rfln + sfln  # :288-321
            if evap_on:
                ev = (prtot > P["ZEPS2"]) & (covpclr > P["ZEPS2"])
                preclr = prtot * covpclr / covptot
                qe = qsk - (qsk - qlim) * covpclr / ((1.0 - clc) ** 2.0)
                beta = RG * P["RPECONS"] * (np.sqrt(apk / aph_s) / 0.00509 * preclr / covpclr) ** 0.5777
                b = dt * beta * (qsk - qe) / (1.0 + dt * beta * corqs)
                dtgdp = dt * RG / (aph[k + 1] - aph[k])
                dpr = np.minimum(covpclr * b / dtgdp, preclr)
                preclr = preclr - dpr
                covptot = W(ev & (preclr <= 0.0), clc, covptot)
                o_covptot[k] = W(ev, covptot, 0.0)
                evapr = W(ev, dpr * rfln / prtot, 0.0).astype(dtype)
                rfln = rfln - evapr
                evaps = W(ev, dpr * sfln / prtot, 0.0).astype(dtype)
                sfln = sfln - evaps
            else:
                evapr = np.zeros(nx, dtype=dtype)
                evaps = np.zeros(nx, dtype=dtype)

            ludek = s["f_lude"][k]  # :328-339
            dqdt = -(condl + condi) + (ludek + evapr + evaps) * gdp
            dtdt = (
                lvdcp * condl
                + lsdcp * condi
                - (
                    lvdcp * evapr
                    + lsdcp * evaps
                    + ludek * (fwat * lvdcp + (1.0 - fwat) * lsdcp)
                    - (lsdcp - lvdcp) * rfreeze
                )
                * gdp
            )

            t = t + dt * dtdt  # :342-347
            q = q + dt * dqdt
            qold = q
            t, q = cuadjtqs_nl(apk, t, q, P)

            dq = np.maximum(qold - q, 0.0)  # :350-364
            dr2 = cons2 * dp * dq
            frz2 = t < RTT
            rfreeze2 = W(frz2, fwat * dr2, 0.0).astype(dtype)
            fwatr = W(frz2, 0.0, 1.0).astype(dtype)
            rn = fwatr * dr2
            sn = (1.0 - fwatr) * dr2
            condl = condl + fwatr * dq / dt
            condi = condi + (1.0 - fwatr) * dq / dt
            rfln = rfln + rn
            sfln = sfln + sn
            rfreeze = rfreeze + rfreeze2

            o_clc[k] = clc  # :367-380
            o_q[k] = -(condl + condi) + (ludek + evapr + evaps) * gdp
            o_t[k] = (
                lvdcp * condl
                + lsdcp * condi
                - (
                    lvdcp * evapr
                    + lsdcp * evaps
                    + ludek * (fwat * lvdcp + (1.0 - fwat) * lsdcp)
                    - (lsdcp - lvdcp) * rfreeze
                )
                * gdp
            )
            o_ql[k] = (qlwc - ql) / dt
            o_qi[k] = (qiwc - qi) / dt

            fplsl[k] = rfln  # :383-388
            fplsn[k] = sfln
            rfl = rfln
            sfl = sfln

        # :391-399 (level 0 of fplsl/fplsn is never written: stays at its initial zero)
        o_fplsl[1:] = fplsl
        o_fplsn[1:] = fplsn
        o_fhpsl[1:] = -o_fplsl[1:] * P["RLVTT"]
        o_fhpsn[1:] = -o_fplsn[1:] * P["RLSTT"]

    tends = {"f_q": o_q, "f_qi": o_qi, "f_ql": o_ql, "f_t": o_t}
    diags = {
        "f_clc": o_clc, "f_covptot": o_covptot, "f_fhpsl": o_fhpsl, "f_fhpsn": o_fhpsn,
        "f_fplsl": o_fplsl, "f_fplsn": o_fplsn,
    }
    return tends, diags


# ----------------------------------------------------------------------------------------
# cuadjtqs_tl -- tangent_linear/_stencils/cuadjtqs.py:22-84
# ----------------------------------------------------------------------------------------
def _cuadjtqs_tl_0(ap, ap_i, t, t_i, q, q_i, z3es, z4es, z5alcp, zaldcp, P):
    RETV, RTT = P["RETV"], P["RTT"]
    qp = 1.0 / ap
    qp_i = -ap_i / ap**2.0
    foeew = P["R2ES"] * np.exp(z3es * (t - RTT) / (t - z4es))
    foeew_i = foeew * z3es * t_i * (RTT - z4es) / (t - z4es) ** 2
    qsat = qp * foeew
    qsat_i = qp_i * foeew + qp * foeew_i
    clip = qsat > P["ZQMAX"]
    qsat = W(clip, P["ZQMAX"], qsat).astype(t.dtype)
    qsat_i = W(clip, 0.0, qsat_i).astype(t.dtype)
    cor = 1.0 / (1.0 - RETV * qsat)
    cor_i = RETV * qsat_i / (1.0 - RETV * qsat) ** 2.0
    qsat_i = qsat_i * cor + qsat * cor_i
    qsat = qsat * cor
    z2s = z5alcp / (t - z4es) ** 2.0
    z2s_i = -2.0 * z5alcp * t_i / (t - z4es) ** 3.0
    cond = (q - qsat) / (1.0 + qsat * cor * z2s)
    cond_i = (q_i - qsat_i) / (1.0 + qsat * cor * z2s) - (q - qsat) * (
        qsat_i * cor * z2s + qsat * cor_i * z2s + qsat * cor * z2s_i
    ) / (1.0 + qsat * cor * z2s) ** 2.0
    t = t + zaldcp * cond
    t_i = t_i + zaldcp * cond_i
    q = q - cond
    q_i = q_i - cond_i
    return t, t_i, q, q_i


def cuadjtqs_tl(ap, ap_i, t, t_i, q, q_i, P):
    z3es, z4es, z5alcp, zaldcp = _cuadjtqs_phase(t, P)
    t, t_i, q, q_i = _cuadjtqs_tl_0(ap, ap_i, t, t_i, q, q_i, z3es, z4es, z5alcp, zaldcp, P)
    t, t_i, q, q_i = _cuadjtqs_tl_0(ap, ap_i, t, t_i, q, q_i, z3es, z4es, z5alcp, zaldcp, P)
    return t, t_i, q, q_i


# ----------------------------------------------------------------------------------------
# cloudsc2_tl -- tangent_linear/_stencils/cloudsc2.py:23-774
# ----------------------------------------------------------------------------------------
def cloudsc2_tl(s, dt, P):
    """s: NL inputs + their `_i` twins.  Returns (tendencies, diagnostics) with the names of
    tangent_linear/microphysics.py:132-160."""
    ap, aph, eta = s["f_ap"], s["f_aph"], s["f_eta"]
    ap_i, aph_i = s["f_ap_i"], s["f_aph_i"]
    dtype = ap.dtype
    nz = ap.shape[0] - 1
    nx = ap.shape[1]
    dt = dtype.type(dt)
    RTT, RG, RETV, RCPD, RVTMP2 = P["RTT"], P["RG"], P["RETV"], P["RCPD"], P["RVTMP2"]
    RLVTT, RLSTT, RLMLT = P["RLVTT"], P["RLSTT"], P["RLMLT"]
    R4LES, R4IES, R5LES, R5IES = P["R4LES"], P["R4IES"], P["R5LES"], P["R5IES"]
    ZEPS2 = P["ZEPS2"]
    LREGCL = bool(P["LREGCL"])
    evap_on = bool(P["LEVAPLS2"] or P["LDRAIN1D"])
    NLEV = nz  # tangent_linear/microphysics.py:67,88

    z = lambda: np.zeros((nz + 1, nx), dtype=dtype)  # noqa: E731
    zc = lambda: np.zeros(nx, dtype=dtype)  # noqa: E731
    names = ("clc", "covptot", "fhpsl", "fhpsn", "fplsl", "fplsn", "tq", "tqi", "tql", "tt")
    o = {n: z() for n in names}
    oi = {n: z() for n in names}

    with _errstate():
        # :124-135
        rfl, rfl_i, sfl, sfl_i, covptot, covptot_i = zc(), zc(), zc(), zc(), zc(), zc()
        aph_s = aph[nz].copy()
        aph_s_i = aph_i[nz].copy()
        # :137-140
        t3d = s["f_t"][:nz] + dt * s["f_tnd_cml_t"][:nz]
        t3d_i = s["f_t_i"][:nz] + dt * s["f_tnd_cml_t_i"][:nz]
        # :142-147
        trpaus = _trpaus(t3d, eta, nz)

        fplsl, fplsl_i = np.zeros((nz, nx), dtype), np.zeros((nz, nx), dtype)
        fplsn, fplsn_i = np.zeros((nz, nx), dtype), np.zeros((nz, nx), dtype)
        for k in range(nz):  # :149-753
            t, t_i = t3d[k], t3d_i[k]
            apk, apk_i = ap[k], ap_i[k]
            qsk, qsk_i = s["f_qsat"][k], s["f_qsat_i"][k]
            q = s["f_q"][k] + dt * s["f_tnd_cml_q"][k] + s["f_supsat"][k]
            q_i = s["f_q_i"][k] + dt * s["f_tnd_cml_q_i"][k] + s["f_supsat_i"][k]
            ql = s["f_ql"][k] + dt * s["f_tnd_cml_ql"][k]
            ql_i = s["f_ql_i"][k] + dt * s["f_tnd_cml_ql_i"][k]
            qi = s["f_qi"][k] + dt * s["f_tnd_cml_qi"][k]
            qi_i = s["f_qi_i"][k] + dt * s["f_tnd_cml_qi_i"][k]

            ckcodtl = 2.0 * P["RKCONV"] * dt
            ckcodti = 5.0 * P["RKCONV"] * dt
            ckcodtla = ckcodtl / 100.0
            ckcodtia = ckcodti / 100.0
            cons2 = 1.0 / (RG * dt)
            cons3 = RLVTT / RCPD
            meltp2 = RTT + 2.0

            scalm = dtype.type(P["ZSCAL"] * max(eta[k] - dtype.type(0.2), dtype.type(P["ZEPS1"])) ** dtype.type(0.2))

            dp = aph[k + 1] - aph[k]
            dp_i = aph_i[k + 1] - aph_i[k]
            zz = 1.0 / (RCPD + RCPD * RVTMP2 * q)
            zz_i = -RCPD * RVTMP2 * q_i / (RCPD + RCPD * RVTMP2 * q) ** 2.0
            lfdcp = RLMLT * zz
            lfdcp_i = RLMLT * zz_i
            lsdcp = RLSTT * zz
            lsdcp_i = RLSTT * zz_i
            lvdcp = RLVTT * zz
            lvdcp_i = RLVTT * zz_i

            # :188-205
            cold = t < RTT
            arg = 0.17 * (t - P["RLPTRC"])
            fwat = W(cold, 0.545 * (np.tanh(arg) + 1.0), 1.0).astype(dtype)
            fwat_i = W(cold, 0.545 * 0.17 * t_i / np.cosh(arg) ** 2.0, 0.0).astype(dtype)
            z3es = W(cold, P["R3IES"], P["R3LES"]).astype(dtype)
            z4es = W(cold, R4IES, R4LES).astype(dtype)
            foeew = P["R2ES"] * np.exp(z3es * (t - RTT) / (t - z4es))
            foeew_i = z3es * (RTT - z4es) * t_i * foeew / (t - z4es) ** 2.0
            esdp = foeew / apk
            esdp_i = foeew_i / apk - foeew * apk_i / (apk**2.0)
            clip = esdp > P["ZQMAX"]
            esdp = W(clip, P["ZQMAX"], esdp).astype(dtype)
            esdp_i = W(clip, 0.0, esdp_i).astype(dtype)

            # :207-222
            facw = R5LES / (t - R4LES) ** 2.0
            facw_i = -2.0 * R5LES * t_i / (t - R4LES) ** 3.0
            faci = R5IES / (t - R4IES) ** 2.0
            faci_i = -2.0 * R5IES * t_i / (t - R4IES) ** 3.0
            fac = fwat * facw + (1.0 - fwat) * faci
            fac_i = fwat_i * (facw - faci) + fwat * facw_i + (1.0 - fwat) * faci_i
            cor = 1.0 / (1.0 - RETV * esdp)
            cor_i = RETV * esdp_i / (1.0 - RETV * esdp) ** 2.0
            dqsdtemp = fac * cor * qsk
            dqsdtemp_i = fac_i * cor * qsk + fac * cor_i * qsk + fac * cor * qsk_i
            corqs = 1.0 + cons3 * dqsdtemp
            corqs_i = cons3 * dqsdtemp_i

            # :224-230
            qclip = q > qsk
            qlim = W(qclip, qsk, q)
            qlim_i = W(qclip, qsk_i, q_i)

            crh2 = _crh2(eta[k], trpaus, dtype)  # :232-253 (uses **0.5 instead of sqrt)

            # :255-265
            ice = t < P["RTICE"]
            supsat = W(ice, 1.8 - 0.003 * t, 1.0).astype(dtype)
            supsat_i = W(ice, -0.003 * t_i, 0.0).astype(dtype)
            qsat = qsk * supsat
            qsat_i = qsk_i * supsat + qsk * supsat_i
            qcrit = crh2 * qsat
            qcrit_i = crh2 * qsat_i

            # :267-306
            qt = q + ql + qi
            qt_i = q_i + ql_i + qi_i
            b1 = qt < qcrit
            b2 = ~b1 & (qt >= qsat)
            qpd = qsat - qt
            qpd_i = qsat_i - qt_i
            qcd = qsat - qcrit
            qcd_i = qsat_i - qcrit_i
            den = qcd - scalm * (qt - qcrit)
            tmp1 = np.sqrt(qpd / den)
            clc3 = 1.0 - tmp1
            clc3_i = -0.5 / tmp1 * (qpd_i * den - qpd * (qcd_i - scalm * (qt_i - qcrit_i))) / den**2.0
            if LREGCL:
                rat = qpd / qcd
                yyy = np.minimum(0.3, 3.5 * np.sqrt(rat * (1.0 - scalm * (1.0 - rat)) ** 3.0) / (1.0 - scalm))
                clc3_i = clc3_i * yyy
            qc3 = (scalm * qpd + (1.0 - scalm) * qcd) * clc3**2.0
            qc3_i = (scalm * qpd_i + (1.0 - scalm) * qcd_i) * clc3**2.0 + 2.0 * (
                scalm * qpd + (1.0 - scalm) * qcd
            ) * clc3 * clc3_i
            clc = W(b1, 0.0, W(b2, 1.0, clc3)).astype(dtype)
            clc_i = W(b1 | b2, 0.0, clc3_i).astype(dtype)
            qc = W(b1, 0.0, W(b2, (1.0 - scalm) * (qsat - qcrit), qc3)).astype(dtype)
            qc_i = W(b1, 0.0, W(b2, (1.0 - scalm) * (qsat_i - qcrit_i), qc3_i)).astype(dtype)

            # :308-325
            gdp = RG / (aph[k + 1] - aph[k])
            gdp_i = -RG * (aph_i[k + 1] - aph_i[k]) / (aph[k + 1] - aph[k]) ** 2.0
            ludek, ludek_i = s["f_lude"][k], s["f_lude_i"][k]
            lude = dt * ludek * gdp
            lude_i = dt * (ludek_i * gdp + ludek * gdp_i)
            lu1, lu1_i = s["f_lu"][k + 1], s["f_lu_i"][k + 1]
            lo1 = (k < NLEV - 1) & (lude >= P["RLMIN"]) & (lu1 >= ZEPS2)
            tmp2 = np.exp(-lude / lu1)
            clc_i = W(
                lo1,
                clc_i + (-clc_i * (1 - tmp2) + (1.0 - clc) * tmp2 * (lude_i / lu1 - lude * lu1_i / lu1**2.0)),
                clc_i,
            )
            clc = W(lo1, clc + (1.0 - clc) * (1.0 - tmp2), clc)
            qc = W(lo1, qc + lude, qc)
            qc_i = W(lo1, qc_i + lude_i, qc_i)

            # :327-373
            fac1 = 1.0 / (P["RD"] * t)
            rho = apk * fac1
            rho_i = (apk_i - apk * t_i / t) * fac1
            fac2 = 1.0 / (apk - RETV * foeew)
            rodqsdp = -rho * qsk * fac2
            rodqsdp_i = (-rho_i * qsk - rho * qsk_i + rho * qsk * (apk_i - RETV * foeew_i) * fac2) * fac2
            ldcp = fwat * lvdcp + (1.0 - fwat) * lsdcp
            ldcp_i = fwat_i * (lvdcp - lsdcp) + fwat * lvdcp_i + (1.0 - fwat) * lsdcp_i
            fac3 = 1.0 / (1.0 + ldcp * dqsdtemp)
            dtdzmo = RG * (1.0 / RCPD - ldcp * rodqsdp) * fac3
            dtdzmo_i = (
                -(RG * (ldcp_i * rodqsdp + ldcp * rodqsdp_i) + dtdzmo * (ldcp_i * dqsdtemp + ldcp * dqsdtemp_i)) * fac3
            )
            dqsdz = dqsdtemp * dtdzmo - RG * rodqsdp
            dqsdz_i = dqsdtemp_i * dtdzmo + dqsdtemp * dtdzmo_i - RG * rodqsdp_i
            mfu, mfd, mfu_i, mfd_i = s["f_mfu"][k], s["f_mfd"][k], s["f_mfu_i"][k], s["f_mfd_i"][k]
            tmp3 = dt * dqsdz * (mfu + mfd) / rho
            lo3 = tmp3 < qc
            dqc_a = tmp3
            dqc_a_i = (dt * (dqsdz_i * (mfu + mfd) + dqsdz * (mfu_i + mfd_i)) - dqc_a * rho_i) / rho
            if LREGCL:
                dqc_a_i = dqc_a_i * 0.1
            dqc = W(lo3, dqc_a, qc)
            dqc_i = W(lo3, dqc_a_i, qc_i)
            qc = qc - dqc
            qc_i = qc_i - dqc_i

            # :375-386
            qlwc = qc * fwat
            qlwc_i = qc_i * fwat + qc * fwat_i
            qiwc = qc * (1.0 - fwat)
            qiwc_i = qc_i * (1.0 - fwat) - qc * fwat_i
            condl = (qlwc - ql) / dt
            condl_i = (qlwc_i - ql_i) / dt
            condi = (qiwc - qi) / dt
            condi_i = (qiwc_i - qi_i) / dt

            # :388-397
            up = clc > covptot
            covptot = W(up, clc, covptot)
            covptot_i = W(up, clc_i, covptot_i)
            covpclr = covptot - clc
            covpclr_i = covptot_i - clc_i
            neg = covpclr < 0.0
            covpclr = W(neg, 0.0, covpclr).astype(dtype)
            covpclr_i = W(neg, 0.0, covpclr_i).astype(dtype)

            # :399-427
            melt = sfl != 0.0
            cons = cons2 * dp / lfdcp
            cons_i = cons2 * (dp_i * lfdcp - dp * lfdcp_i) / lfdcp**2
            warm2 = t > meltp2
            z2s = W(warm2, cons * (t - meltp2), 0.0).astype(dtype)
            z2s_i = W(warm2, cons_i * (t - meltp2) + cons * t_i, 0.0).astype(dtype)
            allm = sfl <= z2s
            snmlt = W(allm, sfl, z2s)
            snmlt_i = W(allm, sfl_i, z2s_i)
            rfln = W(melt, rfl + snmlt, rfl)
            rfln_i = W(melt, rfl_i + snmlt_i, rfl_i)
            sfln = W(melt, sfl - snmlt, sfl)
            sfln_i = W(melt, sfl_i - snmlt_i, sfl_i)
            t_i = W(melt, t_i - (snmlt_i * cons - snmlt * cons_i) / cons**2, t_i)
            t = W(melt, t - snmlt / cons, t)

            # :429-503
            cloudy = clc > ZEPS2
            lcrit = (1.9 if evap_on else 2.0) * P["RCLCRIT"]
            icrit = 0.0001 if evap_on else 2.0 * P["RCLCRIT"]
            cldl = qlwc / clc
            cldl_i = qlwc_i / clc - qlwc * clc_i / clc**2.0
            ltmp4 = np.exp(-((cldl / lcrit) ** 2.0))
            dl = ckcodtl * (1.0 - ltmp4)
            ltmp5 = np.exp(-dl)
            dl_i = (2.0 * (ckcodtla if LREGCL else ckcodtl) / lcrit**2.0) * ltmp4 * cldl * cldl_i
            qlnew = clc * cldl * ltmp5
            qlnew_i = clc_i * cldl * ltmp5 + clc * cldl_i * ltmp5 - clc * cldl * ltmp5 * dl_i
            prr = W(cloudy, qlwc - qlnew, 0.0).astype(dtype)
            prr_i = W(cloudy, qlwc_i - qlnew_i, 0.0).astype(dtype)
            qlwc = W(cloudy, qlwc - prr, qlwc)
            qlwc_i = W(cloudy, qlwc_i - prr_i, qlwc_i)

            cldi = qiwc / clc
            cldi_i = qiwc_i / clc - qiwc * clc_i / clc**2.0
            itmp41 = np.exp(-((cldi / icrit) ** 2.0))
            itmp42 = np.exp(0.025 * (t - RTT))
            di = ckcodti * itmp42 * (1.0 - itmp41)
            itmp5 = np.exp(-di)
            di_i = (
                (ckcodtia if LREGCL else ckcodti)
                * itmp42
                * (itmp41 * (2.0 * cldi * cldi_i / icrit**2.0 - 0.025 * t_i) + 0.025 * t_i)
            )
            qinew = clc * cldi * itmp5
            qinew_i = clc_i * cldi * itmp5 + clc * cldi_i * itmp5 - clc * cldi * itmp5 * di_i
            prs = W(cloudy, qiwc - qinew, 0.0).astype(dtype)
            prs_i = W(cloudy, qiwc_i - qinew_i, 0.0).astype(dtype)
            qiwc = W(cloudy, qiwc - prs, qiwc)
            qiwc_i = W(cloudy, qiwc_i - prs_i, qiwc_i)

            # :505-523
            dr = cons2 * dp * (prr + prs)
            dr_i = cons2 * (dp_i * (prr + prs) + dp * (prr_i + prs_i))
            frz = t < RTT
            rfreeze = W(frz, cons2 * dp * prr, 0.0).astype(dtype)
            rfreeze_i = W(frz, cons2 * (dp_i * prr + dp * prr_i), 0.0).astype(dtype)
            fwatr = W(frz, 0.0, 1.0).astype(dtype)
            fwatr_i = 0.0
            rfln = rfln + fwatr * dr
            rfln_i = rfln_i + (fwatr_i * dr + fwatr * dr_i)
            sfln = sfln + (1.0 - fwatr) * dr
            sfln_i = sfln_i + (-fwatr_i * dr + (1.0 - fwatr) * dr_i)

            # :525-616
            prtot = rfln + sfln
            prtot_i = rfln_i + sfln_i
            if evap_on:
                ev = (prtot > ZEPS2) & (covpclr > ZEPS2)
                preclr = prtot * covpclr / covptot
                preclr_i = (prtot_i * covpclr + prtot * covpclr_i) / covptot - prtot * covpclr * covptot_i / covptot**2.0
                qe = qsk - (qsk - qlim) * covpclr / (1.0 - clc) ** 2.0
                qe_i = (
                    qsk_i
                    - (qsk_i * covpclr - qlim_i * covpclr + (qsk - qlim) * covpclr_i) / (1.0 - clc) ** 2.0
                    - 2.0 * (qsk - qlim) * covpclr * clc_i / (1.0 - clc) ** 3.0
                )
                tmp6 = np.sqrt(apk / aph_s)
                beta = RG * P["RPECONS"] * (tmp6 * preclr / (0.00509 * covpclr)) ** 0.5777
                beta_i = (
                    0.5777
                    * RG
                    * P["RPECONS"]
                    / 0.00509
                    * (0.00509 * covpclr / (tmp6 * preclr)) ** 0.4223
                    * (
                        (tmp6 * preclr_i + 0.5 * preclr * apk_i / tmp6 - 0.5 * preclr * tmp6 * aph_s_i / aph_s)
                        / covpclr
                        - tmp6 * preclr * covpclr_i / covpclr**2
                    )
                )
                b = dt * beta * (qsk - qe) / (1.0 + dt * beta * corqs)
                b_i = dt * (beta_i * (qsk - qe) + beta * (qsk_i - qe_i)) / (1.0 + dt * beta * corqs) - dt**2.0 * b * (
                    beta_i * corqs + beta * corqs_i
                ) / (1 + dt * beta * corqs)
                dtgdp = dt * RG / (aph[k + 1] - aph[k])
                dtgdp_i = -dt * RG * (aph_i[k + 1] - aph_i[k]) / (aph[k + 1] - aph[k]) ** 2.0
                dpr = covpclr * b / dtgdp
                dpr_i = (covpclr_i * b + covpclr * b_i) / dtgdp - covpclr * b * dtgdp_i / dtgdp**2
                cap = dpr > preclr
                dpr = W(cap, preclr, dpr)
                dpr_i = W(cap, preclr_i, dpr_i)
                preclr = preclr - dpr
                preclr_i = preclr_i - dpr_i
                gone = ev & (preclr <= 0.0)
                covptot = W(gone, clc, covptot)
                covptot_i = W(gone, clc_i, covptot_i)
                o["covptot"][k] = W(ev, covptot, 0.0)
                oi["covptot"][k] = W(ev, covptot_i, 0.0)
                evapr = W(ev, dpr * rfln / prtot, 0.0).astype(dtype)
                evapr_i = W(ev, (dpr_i * rfln + dpr * rfln_i) / prtot - dpr * rfln * prtot_i / prtot**2, 0.0).astype(dtype)
                rfln = rfln - evapr
                rfln_i = rfln_i - evapr_i
                evaps = W(ev, dpr * sfln / prtot, 0.0).astype(dtype)
                evaps_i = W(ev, (dpr_i * sfln + dpr * sfln_i) / prtot - dpr * sfln * prtot_i / prtot**2, 0.0).astype(dtype)
                sfln = sfln - evaps
                sfln_i = sfln_i - evaps_i
            else:
                evapr, evapr_i, evaps, evaps_i = zc(), zc(), zc(), zc()

            # :618-651
            dqdt = -(condl + condi) + (ludek + evapr + evaps) * gdp
            dqdt_i = -(condl_i + condi_i) + (ludek_i + evapr_i + evaps_i) * gdp + (ludek + evapr + evaps) * gdp_i
            tmp7 = lvdcp * evapr + lsdcp * evaps + ludek * (fwat * lvdcp + (1.0 - fwat) * lsdcp) - (lsdcp - lvdcp) * rfreeze
            dtdt = lvdcp * condl + lsdcp * condi - tmp7 * gdp
            dtdt_i = (
                lvdcp_i * condl
                + lvdcp * condl_i
                + lsdcp_i * condi
                + lsdcp * condi_i
                - (
                    lvdcp_i * evapr
                    + lvdcp * evapr_i
                    + lsdcp_i * evaps
                    + lsdcp * evaps_i
                    + ludek_i * (fwat * lvdcp + (1.0 - fwat) * lsdcp)
                    + ludek * (fwat_i * (lvdcp - lsdcp) + fwat * lvdcp_i + (1.0 - fwat) * lsdcp_i)
                    - (lsdcp_i - lvdcp_i) * rfreeze
                    - (lsdcp - lvdcp) * rfreeze_i
                )
                * gdp
                - tmp7 * gdp_i
            )

            # :653-662
            t = t + dt * dtdt
            t_i = t_i + dt * dtdt_i
            q = q + dt * dqdt
            q_i = q_i + dt * dqdt_i
            qold, qold_i = q, q_i
            t, t_i, q, q_i = cuadjtqs_tl(apk, apk_i, t, t_i, q, q_i, P)

            # :664-673
            pos = qold >= q
            dq = W(pos, qold - q, 0.0).astype(dtype)
            dq_i = qold_i - q_i
            if LREGCL:
                dq_i = dq_i * 0.7
            dq_i = W(pos, dq_i, 0.0).astype(dtype)
            dr2 = cons2 * dp * dq
            dr2_i = cons2 * (dp_i * dq + dp * dq_i)

            # :675-703
            frz2 = t < RTT
            rfreeze2 = W(frz2, fwat * dr2, 0.0).astype(dtype)
            rfreeze2_i = W(frz2, fwat_i * dr2 + fwat * dr2_i, 0.0).astype(dtype)
            fwatr = W(frz2, 0.0, 1.0).astype(dtype)
            fwatr_i = 0.0
            rn = fwatr * dr2
            rn_i = fwatr_i * dr2 + fwatr * dr2_i
            sn = (1.0 - fwatr) * dr2
            sn_i = -fwatr_i * dr2 + (1.0 - fwatr) * dr2_i
            condl = condl + fwatr * dq / dt
            condl_i = condl_i + (fwatr_i * dq + fwatr * dq_i) / dt
            condi = condi + (1.0 - fwatr) * dq / dt
            condi_i = condi_i + (-fwatr_i * dq + (1.0 - fwatr) * dq_i) / dt
            rfln = rfln + rn
            rfln_i = rfln_i + rn_i
            sfln = sfln + sn
            sfln_i = sfln_i + sn_i
            rfreeze = rfreeze + rfreeze2
            rfreeze_i = rfreeze_i + rfreeze2_i

            # :705-741
            o["clc"][k], oi["clc"][k] = clc, clc_i
            o["tq"][k] = -(condl + condi) + (ludek + evapr + evaps) * gdp
            oi["tq"][k] = -(condl_i + condi_i) + (ludek_i + evapr_i + evaps_i) * gdp + (ludek + evapr + evaps) * gdp_i
            tmp8 = lvdcp * evapr + lsdcp * evaps + ludek * (fwat * lvdcp + (1.0 - fwat) * lsdcp) - (lsdcp - lvdcp) * rfreeze
            o["tt"][k] = lvdcp * condl + lsdcp * condi - tmp8 * gdp
            oi["tt"][k] = (
                lvdcp_i * condl
                + lvdcp * condl_i
                + lsdcp_i * condi
                + lsdcp * condi_i
                - (
                    lvdcp_i * evapr
                    + lvdcp * evapr_i
                    + lsdcp_i * evaps
                    + lsdcp * evaps_i
                    + ludek_i * (fwat * lvdcp + (1.0 - fwat) * lsdcp)
                    + ludek * (fwat_i * (lvdcp - lsdcp) + fwat * lvdcp_i + (1.0 - fwat) * lsdcp_i)
                    - (lsdcp_i - lvdcp_i) * rfreeze
                    - (lsdcp - lvdcp) * rfreeze_i
                )
                * gdp
                - tmp8 * gdp_i
            )
            o["tql"][k] = (qlwc - ql) / dt
            oi["tql"][k] = (qlwc_i - ql_i) / dt
            o["tqi"][k] = (qiwc - qi) / dt
            oi["tqi"][k] = (qiwc_i - qi_i) / dt

            # :743-753
            fplsl[k], fplsl_i[k], fplsn[k], fplsn_i[k] = rfln, rfln_i, sfln, sfln_i
            rfl, rfl_i, sfl, sfl_i = rfln, rfln_i, sfln, sfln_i

        # :755-774
        o["fplsl"][1:], oi["fplsl"][1:] = fplsl, fplsl_i
        o["fplsn"][1:], oi["fplsn"][1:] = fplsn, fplsn_i
        o["fhpsl"][1:] = -o["fplsl"][1:] * RLVTT
        oi["fhpsl"][1:] = -oi["fplsl"][1:] * RLVTT
        o["fhpsn"][1:] = -o["fplsn"][1:] * RLSTT
        oi["fhpsn"][1:] = -oi["fplsn"][1:] * RLSTT

    tends = {
        "f_q": o["tq"], "f_q_i": oi["tq"], "f_qi": o["tqi"], "f_qi_i": oi["tqi"],
        "f_ql": o["tql"], "f_ql_i": oi["tql"], "f_t": o["tt"], "f_t_i": oi["tt"],
    }
    diags = {}
    for n in ("clc", "covptot", "fhpsl", "fhpsn", "fplsl", "fplsn"):
        diags["f_" + n] = o[n]
        diags["f_" + n + "_i"] = oi[n]
    return tends, diags


# ----------------------------------------------------------------------------------------
# cuadjtqs_ad -- adjoint/_stencils/cuadjtqs.py:22-158
# ----------------------------------------------------------------------------------------
def cuadjtqs_ad(ap, ap_i, t, t_i, q, q_i, P):
    """Returns (ap_i, t, t_i, q, q_i) like the reference function."""
    R2ES, RETV, RTT, ZQMAX = P["R2ES"], P["RETV"], P["RTT"], P["ZQMAX"]
    dtype = t.dtype
    z3es, z4es, z5alcp, zaldcp = _cuadjtqs_phase(t, P)

    # first Newton step, trajectory saved with suffix _b / _d (:52-69)
    targ = t
    foeew = R2ES * np.exp(z3es * (targ - RTT) / (targ - z4es))
    foeew_b = foeew
    qsat = foeew / ap
    ltest2 = qsat > ZQMAX
    qsat = W(ltest2, ZQMAX, qsat).astype(dtype)
    cor = 1.0 / (1.0 - RETV * qsat)
    qsat_d = qsat
    qsat = qsat * cor
    targ_b = targ
    z2s = z5alcp / (targ - z4es) ** 2.0
    qsat_b, cor_b, z2s_b, q_b = qsat, cor, z2s, q
    cond1 = (q - qsat) / (1.0 + qsat * cor * z2s)
    t = t + zaldcp * cond1
    q = q - cond1

    # second Newton step, suffix _a / _c (:71-91)
    targ = t
    foeew = R2ES * np.exp(z3es * (targ - RTT) / (targ - z4es))
    foeew_a = foeew
    qsat = foeew / ap
    ltest1 = qsat > ZQMAX
    qsat = W(ltest1, ZQMAX, qsat).astype(dtype)
    cor = 1.0 / (1.0 - RETV * qsat)
    qsat_c = qsat
    qsat = qsat * cor
    targ_a = targ
    z2s = z5alcp / (targ - z4es) ** 2.0
    qsat_a, cor_a, z2s_a, q_a = qsat, cor, z2s, q
    cond1 = (q - qsat) / (1.0 + qsat * cor * z2s)
    t = t + zaldcp * cond1
    q = q - cond1

    # reverse of the second step (:93-124)
    cond1_i = -q_i + zaldcp * t_i
    qsat, cor, z2s = qsat_a, cor_a, z2s_a
    q_i = q_i + cond1_i / (1.0 + qsat * cor * z2s)
    qsat_i = -cond1_i / (1.0 + qsat * cor * z2s) - cond1_i * (q_a - qsat) * cor * z2s / (1.0 + qsat * cor * z2s) ** 2.0
    cor_i = -cond1_i * (q_a - qsat) * qsat * z2s / (1.0 + qsat * cor * z2s) ** 2.0
    z2s_i = -cond1_i * (q_a - qsat) * qsat * cor / (1.0 + qsat * cor * z2s) ** 2.0
    targ = targ_a
    targ_i = -2.0 * z2s_i * z5alcp / (targ - z4es) ** 3.0
    qsat = qsat_c
    cor_i = cor_i + qsat_i * qsat
    qsat_i = qsat_i * cor
    qsat_i = qsat_i + cor_i * RETV / (1.0 - RETV * qsat) ** 2.0
    qsat_i = W(ltest1, 0.0, qsat_i).astype(dtype)
    foeew_i = qsat_i / ap
    foeew = foeew_a
    qp_i = qsat_i * foeew
    targ_i = targ_i + (
        foeew_i * R2ES * z3es * (RTT - z4es) * np.exp(z3es * (targ - RTT) / (targ - z4es)) / (targ - z4es) ** 2.0
    )
    t_i = t_i + targ_i

    # reverse of the first step (:126-156)
    cond1_i = -q_i + zaldcp * t_i
    qsat, cor, z2s = qsat_b, cor_b, z2s_b
    q_i = q_i + cond1_i / (1.0 + qsat * cor * z2s)
    qsat_i = -cond1_i / (1.0 + qsat * cor * z2s) - cond1_i * (q_b - qsat) * cor * z2s / (1.0 + qsat * cor * z2s) ** 2.0
    cor_i = -cond1_i * (q_b - qsat) * qsat * z2s / (1.0 + qsat * cor * z2s) ** 2.0
    z2s_i = -cond1_i * (q_b - qsat) * qsat * cor / (1.0 + qsat * cor * z2s) ** 2.0
    targ = targ_b
    targ_i = -2.0 * z2s_i * z5alcp / (targ - z4es) ** 3.0
    qsat = qsat_d
    cor_i = cor_i + qsat_i * qsat
    qsat_i = qsat_i * cor
    qsat_i = qsat_i + cor_i * RETV / (1.0 - RETV * qsat) ** 2.0
    qsat_i = W(ltest2, 0.0, qsat_i).astype(dtype)
    foeew_i = qsat_i / ap
    foeew = foeew_b
    qp_i = qp_i + qsat_i * foeew
    targ_i = targ_i + (
        foeew_i * R2ES * z3es * (RTT - z4es) * np.exp(z3es * (targ - RTT) / (targ - z4es)) / (targ - z4es) ** 2.0
    )
    t_i = t_i + targ_i
    ap_i = ap_i - qp_i / ap**2.0
    return ap_i, t, t_i, q, q_i


class _NS:
    """Per-level bag of trajectory temporaries (the reference keeps them as 3-D temporaries)."""


# ----------------------------------------------------------------------------------------
# cloudsc2_ad -- adjoint/_stencils/cloudsc2.py:24-996
# ----------------------------------------------------------------------------------------
def cloudsc2_ad(s, dt, P, predicates="reference"):
    """s: NL inputs + the adjoint seeds f_tnd_{t,q,ql,qi}_i, f_clc_i, f_covptot_i,
    f_fhpsl_i, f_fhpsn_i, f_fplsl_i, f_fplsn_i (adjoint/microphysics.py:106-120).

    The seed arrays in `s` are **modified in place** exactly as the reference does (it zeroes
    them; adjoint/_stencils/cloudsc2.py:482-484,506-542,650,714,920,972-984).

    predicates="reference": literal restatement (second freezing test on the pre-adjustment
      `t3`, :427/:577; backward first freezing test on the post-adjustment `t`, :729).
    predicates="tl": every branch predicate equals the TL predicate on the same trajectory
      (tangent_linear/_stencils/cloudsc2.py:510,677) -- what an exact adjoint of the TL needs
      when a level crosses RTT during the saturation adjustment.

    Returns (tendencies, diagnostics) with the names of adjoint/microphysics.py:123-157."""
    assert predicates in ("reference", "tl")
    ap, aph, eta = s["f_ap"], s["f_aph"], s["f_eta"]
    dtype = ap.dtype
    nz = ap.shape[0] - 1
    nx = ap.shape[1]
    dt = dtype.type(dt)
    RTT, RG, RETV, RCPD, RVTMP2 = P["RTT"], P["RG"], P["RETV"], P["RCPD"], P["RVTMP2"]
    RLVTT, RLSTT, RLMLT = P["RLVTT"], P["RLSTT"], P["RLMLT"]
    R4LES, R4IES, R5LES, R5IES = P["R4LES"], P["R4IES"], P["R5LES"], P["R5IES"]
    ZEPS2, RPECONS = P["ZEPS2"], P["RPECONS"]
    LREGCL = bool(P["LREGCL"])
    evap_on = bool(P["LEVAPLS2"] or P["LDRAIN1D"])
    NLEV = nz

    z = lambda: np.zeros((nz + 1, nx), dtype=dtype)  # noqa: E731
    zc = lambda: np.zeros(nx, dtype=dtype)  # noqa: E731
    o_clc, o_covptot, o_fhpsl, o_fhpsn, o_fplsl, o_fplsn = z(), z(), z(), z(), z(), z()
    o_tq, o_tqi, o_tql, o_tt = z(), z(), z(), z()
    o_ap_i, o_aph_i, o_lu_i, o_lude_i, o_mfd_i, o_mfu_i = z(), z(), z(), z(), z(), z()
    o_q_i, o_qi_i, o_ql_i, o_qsat_i, o_supsat_i, o_t_i = z(), z(), z(), z(), z(), z()
    o_cml_q_i, o_cml_qi_i, o_cml_ql_i, o_cml_t_i = z(), z(), z(), z()

    in_clc_i, in_covptot_i = s["f_clc_i"], s["f_covptot_i"]
    in_fhpsl_i, in_fhpsn_i = s["f_fhpsl_i"], s["f_fhpsn_i"]
    in_fplsl_i, in_fplsn_i = s["f_fplsl_i"], s["f_fplsn_i"]
    in_tq_i, in_tqi_i, in_tql_i, in_tt_i = s["f_tnd_q_i"], s["f_tnd_qi_i"], s["f_tnd_ql_i"], s["f_tnd_t_i"]

    with _errstate():
        # ============================== forward (trajectory) ==============================
        # :124-131
        covptotp, rfln, sfln = zc(), zc(), zc()
        aph_s = aph[nz].copy()
        # :133-137
        t3d = s["f_t"][:nz] + dt * s["f_tnd_cml_t"][:nz]
        # :139-144
        trpaus = _trpaus(t3d, eta, nz)

        ckcodtl = 2.0 * P["RKCONV"] * dt
        ckcodti = 5.0 * P["RKCONV"] * dt
        cons2 = 1.0 / (RG * dt)
        cons3 = RLVTT / RCPD
        meltp2 = RTT + 2.0
        lcrit = (1.9 if evap_on else 2.0) * P["RCLCRIT"]
        icrit = 0.0001 if evap_on else 2.0 * P["RCLCRIT"]

        L = []
        rfl3d = np.zeros((nz + 1, nx), dtype)
        sfl3d = np.zeros((nz + 1, nx), dtype)
        for k in range(nz):  # :146-458
            n = _NS()
            L.append(n)
            t = t3d[k]
            n.t2 = t2 = t
            apk, qsk = ap[k], s["f_qsat"][k]
            n.rfl = rfl = rfln
            n.sfl = sfl = sfln
            rfl3d[k], sfl3d[k] = rfl, sfl

            q = s["f_q"][k] + dt * s["f_tnd_cml_q"][k] + s["f_supsat"][k]
            n.ql = ql = s["f_ql"][k] + dt * s["f_tnd_cml_ql"][k]
            n.qi = qi = s["f_qi"][k] + dt * s["f_tnd_cml_qi"][k]
            n.q2 = q2 = q

            n.scalm = scalm = dtype.type(
                P["ZSCAL"] * max(eta[k] - dtype.type(0.2), dtype.type(P["ZEPS1"])) ** dtype.type(0.2)
            )

            n.dp = dp = aph[k + 1] - aph[k]
            zz = RCPD + RCPD * RVTMP2 * q
            n.lfdcp = lfdcp = RLMLT / zz
            n.lsdcp = lsdcp = RLSTT / zz
            n.lvdcp = lvdcp = RLVTT / zz

            # :180-197
            cold = t < RTT
            n.fwat = fwat = W(cold, 0.545 * (np.tanh(0.17 * (t2 - P["RLPTRC"])) + 1.0), 1.0).astype(dtype)
            z3es = W(cold, P["R3IES"], P["R3LES"]).astype(dtype)
            z4es = W(cold, R4IES, R4LES).astype(dtype)
            n.foeew = foeew = P["R2ES"] * np.exp(z3es * (t2 - RTT) / (t2 - z4es))
            n.esdp1 = esdp1 = foeew / apk
            esdp = np.minimum(esdp1, P["ZQMAX"])
            n.facw = facw = R5LES / (t2 - R4LES) ** 2.0
            n.faci = faci = R5IES / (t2 - R4IES) ** 2.0
            n.fac = fac = fwat * facw + (1.0 - fwat) * faci
            n.cor = cor = 1.0 / (1.0 - RETV * esdp)
            n.dqsdtemp = dqsdtemp = fac * cor * qsk
            n.corqs = corqs = 1.0 + cons3 * dqsdtemp

            n.qlim = qlim = np.minimum(q2, qsk)  # :200

            n.crh2 = crh2 = _crh2(eta[k], trpaus, dtype)  # :202-223

            # :225-231
            n.supsat = supsat = W(t2 < P["RTICE"], 1.8 - 0.003 * t2, 1.0).astype(dtype)
            n.qsat = qsat = qsk * supsat
            n.qcrit = qcrit = crh2 * qsat

            # :233-252
            n.qt = qt = q + ql + qi
            b1 = qt <= qcrit
            b2 = ~b1 & (qt >= qsat)
            b3 = ~b1 & ~b2
            qcd3 = qsat - qcrit
            qpd3 = qsat - qt
            tmp33 = np.sqrt(qpd3 / (qcd3 - scalm * (qt - qcrit)))
            clc3 = 1.0 - tmp33
            n.qcd = qcd = W(b3, qcd3, 0.0).astype(dtype)
            n.qpd = qpd = W(b3, qpd3, 0.0).astype(dtype)
            n.tmp3 = W(b3, tmp33, 0.0).astype(dtype)
            n.clc = clc = W(b1, 0.0, W(b2, 1.0, clc3)).astype(dtype)
            qc1 = W(
                b1, 0.0, W(b2, (1.0 - scalm) * (qsat - qcrit), (scalm * qpd3 + (1.0 - scalm) * qcd3) * clc3**2.0)
            ).astype(dtype)

            # :254-263
            n.gdp = gdp = RG / (aph[k + 1] - aph[k])
            ludek = s["f_lude"][k]
            n.lude = lude = dt * ludek * gdp
            lu1 = s["f_lu"][k + 1]
            lo1 = (lude >= P["RLMIN"]) & (lu1 >= ZEPS2)
            n.oclc = oclc = W(lo1, clc + (1.0 - clc) * (1.0 - np.exp(-lude / lu1)), clc)
            qc2 = W(lo1, qc1 + lude, qc1)

            # :265-277
            n.fac1 = fac1 = 1.0 / (P["RD"] * t2)
            n.rho = rho = apk * fac1
            n.fac2 = fac2 = 1.0 / (apk - RETV * foeew)
            n.rodqsdp = rodqsdp = -rho * qsk * fac2
            n.ldcp = ldcp = fwat * lvdcp + (1.0 - fwat) * lsdcp
            n.fac3 = fac3 = 1.0 / (1.0 + ldcp * dqsdtemp)
            n.dtdzmo = dtdzmo = RG * (1.0 / RCPD - ldcp * rodqsdp) * fac3
            n.dqsdz = dqsdz = dqsdtemp * dtdzmo - RG * rodqsdp
            n.fac4 = fac4 = 1.0 / rho
            mfu, mfd = s["f_mfu"][k], s["f_mfd"][k]
            n.lo3 = dt * dqsdz * (mfu + mfd) * fac4 < qc2
            n.dqc = dqc = np.minimum(dt * dqsdz * (mfu + mfd) * fac4, qc2)
            n.qc3 = qc3 = qc2 - dqc

            # :279-283
            n.qlwc1 = qlwc1 = qc3 * fwat
            n.qiwc1 = qiwc1 = qc3 * (1.0 - fwat)
            n.condl1 = condl1 = (qlwc1 - ql) / dt
            n.condi1 = condi1 = (qiwc1 - qi) / dt

            # :285-290
            n.covptot1 = covptot1 = np.maximum(covptotp, oclc)
            covptot = covptot1
            n.covpclr1 = covpclr1 = covptot - oclc
            n.covpclr = covpclr = np.maximum(covpclr1, 0.0)

            # :292-302
            n.melt = melt = sfl != 0.0
            n.cons = cons = cons2 * dp / lfdcp
            n.z2s = z2s = cons * np.maximum(t2 - meltp2, 0.0)
            n.snmlt = snmlt = np.minimum(sfl, z2s)
            rfln = W(melt, rfl + snmlt, rfl)
            sfln = W(melt, sfl - snmlt, sfl)
            t = W(melt, t2 - snmlt / cons, t)
            n.tmelt = t

            # :304-337
            n.cloudy = cloudy = oclc > ZEPS2
            n.cldl = cldl = qlwc1 / oclc
            n.ltmp1 = ltmp1 = np.exp(-((cldl / lcrit) ** 2.0))
            dl = ckcodtl * (1.0 - ltmp1)
            n.ltmp2 = ltmp2 = np.exp(-dl)
            qlnew = oclc * cldl * ltmp2
            n.prr = prr = W(cloudy, qlwc1 - qlnew, 0.0).astype(dtype)
            qlwc = W(cloudy, qlwc1 - prr, qlwc1)
            n.cldi = cldi = qiwc1 / oclc
            n.itmp11 = itmp11 = np.exp(-((cldi / icrit) ** 2.0))
            n.itmp12 = itmp12 = np.exp(0.025 * (t - RTT))
            di = ckcodti * itmp12 * (1.0 - itmp11)
            n.itmp2 = itmp2 = np.exp(-di)
            qinew = oclc * cldi * itmp2
            n.prs = prs = W(cloudy, qiwc1 - qinew, 0.0).astype(dtype)
            qiwc = W(cloudy, qiwc1 - prs, qiwc1)

            # :339-353
            dr1 = cons2 * dp * (prr + prs)
            n.frz1 = frz1 = t < RTT
            n.rfreeze1 = rfreeze1 = W(frz1, cons2 * dp * prr, 0.0).astype(dtype)
            n.fwatr1 = fwatr1 = W(frz1, 0.0, 1.0).astype(dtype)
            rfln = rfln + fwatr1 * dr1
            sfln = sfln + (1.0 - fwatr1) * dr1
            n.rfln2, n.sfln2 = rfln2, sfln2 = rfln, sfln

            # :355-394
            n.prtot = prtot = rfln + sfln
            if evap_on:
                n.ev = ev = (prtot > ZEPS2) & (covpclr > ZEPS2)
                n.preclr1 = preclr1 = prtot * covpclr / covptot1
                n.qe = qe = qsk - (qsk - qlim) * covpclr / (1.0 - oclc) ** 2.0
                n.beta = beta = RG * RPECONS * (np.sqrt(apk / aph_s) / 0.00509 * preclr1 / covpclr) ** 0.5777
                n.b = b = dt * beta * (qsk - qe) / (1.0 + dt * beta * corqs)
                n.dtgdp = dtgdp = dt * RG / (aph[k + 1] - aph[k])
                n.dpr1 = dpr1 = covpclr * b / dtgdp
                n.dpr = dpr = np.minimum(dpr1, preclr1)
                n.preclr = preclr = preclr1 - dpr
                covptot = W(ev & (preclr <= 0.0), oclc, covptot)
                o_covptot[k] = W(ev, covptot, 0.0)
                n.evapr = evapr = W(ev, dpr * rfln2 / prtot, 0.0).astype(dtype)
                rfln = rfln - evapr
                n.evaps = evaps = W(ev, dpr * sfln2 / prtot, 0.0).astype(dtype)
                sfln = sfln - evaps
            else:
                n.ev = np.zeros(nx, dtype=bool)
                n.evapr = evapr = zc()
                n.evaps = evaps = zc()
            n.covptot = covptot

            # :396-419
            dqdt = -(condl1 + condi1) + (ludek + evapr + evaps) * gdp
            dtdt = (
                lvdcp * condl1
                + lsdcp * condi1
                - (
                    lvdcp * evapr
                    + lsdcp * evaps
                    + ludek * (fwat * lvdcp + (1.0 - fwat) * lsdcp)
                    - (lsdcp - lvdcp) * rfreeze1
                )
                * gdp
            )
            n.t3 = t3 = t + dt * dtdt
            q = q2 + dt * dqdt
            n.told = t3
            n.qold = n.qold1 = qold1 = q

            # :421-439
            t, q = cuadjtqs_nl(apk, t3, q, P)
            n.tpost, n.qpost = t, q
            n.dq = dq = np.maximum(qold1 - q, 0.0)
            n.dr2 = dr2 = cons2 * dp * dq
            n.frz2 = frz2 = (t3 < RTT) if predicates == "reference" else (t < RTT)
            rfreeze2 = W(frz2, fwat * dr2, 0.0).astype(dtype)
            n.fwatr2 = fwatr2 = W(frz2, 0.0, 1.0).astype(dtype)
            rn = fwatr2 * dr2
            sn = (1.0 - fwatr2) * dr2
            n.condl2 = condl2 = condl1 + fwatr2 * dq / dt
            n.condi2 = condi2 = condi1 + (1.0 - fwatr2) * dq / dt
            rfln = rfln + rn
            sfln = sfln + sn
            n.rfreeze3 = rfreeze3 = rfreeze1 + rfreeze2

            # :441-458
            o_clc[k] = oclc
            o_tq[k] = -(condl2 + condi2) + (ludek + evapr + evaps) * gdp
            o_tt[k] = (
                lvdcp * condl2
                + lsdcp * condi2
                - (
                    lvdcp * evapr
                    + lsdcp * evaps
                    + ludek * (fwat * lvdcp + (1.0 - fwat) * lsdcp)
                    - (lsdcp - lvdcp) * rfreeze3
                )
                * gdp
            )
            o_tql[k] = (qlwc - ql) / dt
            o_tqi[k] = (qiwc - qi) / dt
            covptotp = covptot
        rfl3d[nz], sfl3d[nz] = rfln, sfln  # :459-462

        # :464-475
        o_fplsl[1:] = rfl3d[1:]
        o_fplsn[1:] = sfl3d[1:]
        o_fhpsl[1:] = -o_fplsl[1:] * RLVTT
        o_fhpsn[1:] = -o_fplsn[1:] * RLSTT

        # ================================ adjoint computations ============================
        # :479-484
        in_fplsn_i -= in_fhpsn_i * RLSTT
        in_fhpsn_i[...] = 0.0
        in_fplsl_i -= in_fhpsl_i * RLVTT
        in_fhpsl_i[...] = 0.0

        # :486-493
        covptot_i3d = np.zeros((nz + 1, nx), dtype)
        rfl_i3d = np.zeros((nz + 1, nx), dtype)  # zero where the melt branch does not assign it
        sfl_i3d = np.zeros((nz + 1, nx), dtype)
        aph_s_i, rfln_i, sfln_i = zc(), zc(), zc()
        daph_i3d = np.zeros((nz, nx), dtype)
        dp_i3d = np.zeros((nz, nx), dtype)
        dlu_i3d = np.zeros((nz, nx), dtype)
        lvdcp_i3d = np.zeros((nz, nx), dtype)
        lsdcp_i3d = np.zeros((nz, nx), dtype)
        lfdcp_i3d = np.zeros((nz, nx), dtype)

        ckcodtla = ckcodtl / 100.0
        ckcodtia = ckcodti / 100.0

        for k in range(nz - 1, -1, -1):  # :494-967
            n = L[k]
            apk, qsk = ap[k], s["f_qsat"][k]
            ludek = s["f_lude"][k]
            mfu, mfd = s["f_mfu"][k], s["f_mfd"][k]
            lu1 = s["f_lu"][k + 1]
            fwat, lvdcp, lsdcp, lfdcp, gdp, dp = n.fwat, n.lvdcp, n.lsdcp, n.lfdcp, n.gdp, n.dp
            evapr, evaps = n.evapr, n.evaps
            oclc = n.oclc
            t2 = n.t2

            # :499-501
            rfln_i = rfln_i + (rfl_i3d[k + 1] + in_fplsl_i[k + 1])
            sfln_i = sfln_i + (sfl_i3d[k + 1] + in_fplsn_i[k + 1])

            # :503-511
            oqi_i = -in_tqi_i[k] / dt
            qiwc_i = in_tqi_i[k] / dt
            in_tqi_i[k] = 0.0
            oql_i = -in_tql_i[k] / dt
            qlwc_i = in_tql_i[k] / dt
            in_tql_i[k] = 0.0

            # :513-533
            tti = in_tt_i[k].copy()
            gdp_i = -tti * (
                lvdcp * evapr + lsdcp * evaps + ludek * (fwat * lvdcp + (1.0 - fwat) * lsdcp) - (lsdcp - lvdcp) * n.rfreeze3
            )
            condl_i = tti * lvdcp
            condi_i = tti * lsdcp
            evapr_i = -tti * lvdcp * gdp
            evaps_i = -tti * lsdcp * gdp
            lvdcp_i = tti * (n.condl2 - evapr * gdp)
            lsdcp_i = tti * (n.condi2 - evaps * gdp)
            olude_i = o_lude_i[k] - tti * gdp * (fwat * lvdcp + (1.0 - fwat) * lsdcp)
            lvdcp_i = lvdcp_i - tti * ludek * gdp * fwat
            lsdcp_i = lsdcp_i - tti * ludek * gdp * (1.0 - fwat)
            fwat_i = -tti * ludek * gdp * (lvdcp - lsdcp)
            lvdcp_i = lvdcp_i - tti * n.rfreeze3 * gdp
            lsdcp_i = lsdcp_i + tti * n.rfreeze3 * gdp
            rfreeze_i = tti * (lsdcp - lvdcp) * gdp
            in_tt_i[k] = 0.0

            # :535-542
            tqi = in_tq_i[k].copy()
            gdp_i = gdp_i + tqi * (ludek + evapr + evaps)
            olude_i = olude_i + tqi * gdp
            evapr_i = evapr_i + tqi * gdp
            evaps_i = evaps_i + tqi * gdp
            condl_i = condl_i - tqi
            condi_i = condi_i - tqi
            in_tq_i[k] = 0.0

            # :565-582
            rn_i = rfln_i
            sn_i = sfln_i
            fwatr2 = n.fwatr2
            dq_i = (fwatr2 * condl_i + (1.0 - fwatr2) * condi_i) / dt
            dr2_i = fwatr2 * rn_i + (1.0 - fwatr2) * sn_i
            fwat_i = W(n.frz2, fwat_i + n.dr2 * rfreeze_i, fwat_i)
            dr2_i = W(n.frz2, dr2_i + fwat * rfreeze_i, dr2_i)
            dq_i = dq_i + cons2 * dp * dr2_i
            dp_i = cons2 * n.dq * dr2_i

            # :584-592
            pos = n.qold1 >= n.qpost
            if LREGCL:
                dq_i = W(pos, dq_i * 0.7, dq_i)
            qold_i = W(pos, dq_i, 0.0).astype(dtype)
            oq_i = W(pos, -dq_i, 0.0).astype(dtype)

            # :594-598
            oap_i, _, ot_i, _, oq_i = cuadjtqs_ad(apk, zc(), n.told, zc(), n.qold, oq_i, P)

            # :600-603
            oq_i = oq_i + qold_i
            dqdt_i = dt * oq_i
            dtdt_i = dt * ot_i

            # :605-633
            gdp_i = gdp_i - dtdt_i * (
                lvdcp * evapr + lsdcp * evaps + ludek * (fwat * lvdcp + (1.0 - fwat) * lsdcp) - (lsdcp - lvdcp) * n.rfreeze1
            )
            condl_i = condl_i + dtdt_i * lvdcp
            condi_i = condi_i + dtdt_i * lsdcp
            evapr_i = evapr_i - dtdt_i * lvdcp * gdp
            evaps_i = evaps_i - dtdt_i * lsdcp * gdp
            lvdcp_i = lvdcp_i + dtdt_i * (n.condl1 - evapr * gdp)
            lsdcp_i = lsdcp_i + dtdt_i * (n.condi1 - evaps * gdp)
            olude_i = olude_i - dtdt_i * gdp * (fwat * lvdcp + (1.0 - fwat) * lsdcp)
            lvdcp_i = lvdcp_i - dtdt_i * ludek * gdp * fwat
            lsdcp_i = lsdcp_i - dtdt_i * ludek * gdp * (1.0 - fwat)
            fwat_i = fwat_i - dtdt_i * ludek * gdp * (lvdcp - lsdcp)
            lvdcp_i = lvdcp_i - dtdt_i * n.rfreeze1 * gdp
            lsdcp_i = lsdcp_i + dtdt_i * n.rfreeze1 * gdp
            rfreeze_i = rfreeze_i + dtdt_i * (lsdcp - lvdcp) * gdp
            gdp_i = gdp_i + dqdt_i * (ludek + evapr + evaps)
            olude_i = olude_i + dqdt_i * gdp
            evapr_i = evapr_i + dqdt_i * gdp
            evaps_i = evaps_i + dqdt_i * gdp
            condl_i = condl_i - dqdt_i
            condi_i = condi_i - dqdt_i

            # :635-719
            if evap_on:
                ev = n.ev
                prtot, dpr, covpclr, covptot1 = n.prtot, n.dpr, n.covpclr, n.covptot1
                e_evaps_i = evaps_i - sfln_i
                e_sfln_i = sfln_i + dpr * e_evaps_i / prtot
                dpr_i = n.sfln2 * e_evaps_i / prtot
                prtot_i = -dpr * n.sfln2 * e_evaps_i / prtot**2.0
                e_evapr_i = evapr_i - rfln_i
                e_rfln_i = rfln_i + dpr * e_evapr_i / prtot
                dpr_i = dpr_i + n.rfln2 * e_evapr_i / prtot
                prtot_i = prtot_i - dpr * n.rfln2 * e_evapr_i / prtot**2.0

                e_covptot_i = covptot_i3d[k + 1] + in_covptot_i[k]
                gone = n.preclr <= 0
                e_clc_add = W(gone, e_covptot_i, 0.0)
                e_covptot_i = W(gone, 0.0, e_covptot_i)

                capped = n.dpr1 > n.preclr1
                preclr_i = W(capped, dpr_i, 0.0)
                dpr_i = W(capped, 0.0, dpr_i)

                b_i = covpclr * dpr_i / n.dtgdp
                covpclr_i = n.b * dpr_i / n.dtgdp
                dtgdp_i = -covpclr * n.b * dpr_i / n.dtgdp**2.0
                e_daph_i = dt * RG * dtgdp_i / (aph[k + 1] - aph[k])

                tmp1 = 1.0 + dt * n.beta * n.corqs
                beta_i = dt * (qsk - n.qe) * b_i / tmp1 - (dt**2.0) * n.beta * (qsk - n.qe) * n.corqs * b_i / tmp1**2.0
                e_oqsat_i = dt * n.beta * b_i / tmp1
                qe_i = -dt * n.beta * b_i / tmp1
                e_corqs_i = -(dt**2.0) * n.beta * (qsk - n.qe) * n.beta * b_i / tmp1**2.0

                sq = np.sqrt(apk / aph_s)
                xx = 0.5777 * (RG * RPECONS / 0.00509) * (0.00509 * covpclr / (n.preclr1 * sq)) ** 0.4223
                preclr_i = preclr_i + xx * sq * beta_i / covpclr
                e_oap_add = 0.5 * xx * n.preclr1 * beta_i / (covpclr * np.sqrt(apk * aph_s))
                e_aph_s_sub = 0.5 * xx * n.preclr1 * sq * beta_i / (covpclr * aph_s)
                covpclr_i = covpclr_i + ((
                    -(xx * n.preclr1 * sq * beta_i / covpclr**2.0) - (qsk - n.qlim) * qe_i / (1.0 - oclc) ** 2.0
                ) + prtot * preclr_i / covptot1)  # `+=` of the whole right-hand side, :693-696
                e_oqsat_i = e_oqsat_i + (qe_i - covpclr * qe_i / (1.0 - oclc) ** 2.0)
                e_qlim_i = covpclr * qe_i / (1.0 - oclc) ** 2.0
                e_clc_sub = 2.0 * (qsk - n.qlim) * covpclr * qe_i / (1.0 - oclc) ** 3.0
                prtot_i = prtot_i + covpclr * preclr_i / covptot1
                e_covptot_i = e_covptot_i - prtot * covpclr * preclr_i / covptot1**2.0

                evaps_i = W(ev, e_evaps_i, evaps_i)
                evapr_i = W(ev, e_evapr_i, evapr_i)
                sfln_i = W(ev, e_sfln_i, sfln_i)
                rfln_i = W(ev, e_rfln_i, rfln_i)
                in_clc_i[k] = W(ev, (in_clc_i[k] + e_clc_add) - e_clc_sub, in_clc_i[k])  # :656 then :700-706
                corqs_i = W(ev, e_corqs_i, 0.0).astype(dtype)
                covpclr_i = W(ev, covpclr_i, 0.0).astype(dtype)
                covptot_i = W(ev, e_covptot_i, 0.0).astype(dtype)
                daph_i = W(ev, e_daph_i, 0.0).astype(dtype)
                oqsat_i = W(ev, e_oqsat_i, 0.0).astype(dtype)
                prtot_i = W(ev, prtot_i, 0.0).astype(dtype)
                qlim_i = W(ev, e_qlim_i, 0.0).astype(dtype)
                oap_i = W(ev, oap_i + e_oap_add, oap_i)
                aph_s_i = W(ev, aph_s_i - e_aph_s_sub, aph_s_i)
                in_covptot_i[k] = 0.0
            else:
                corqs_i, covpclr_i, covptot_i, daph_i, oqsat_i, prtot_i, qlim_i = (
                    zc(), zc(), zc(), zc(), zc(), zc(), zc()
                )
                in_covptot_i[k] = 0.0

            # :721-736
            rfln_i = rfln_i + prtot_i
            sfln_i = sfln_i + prtot_i
            dr_i = n.fwatr1 * rfln_i + (1.0 - n.fwatr1) * sfln_i
            frz_b = (n.tpost < RTT) if predicates == "reference" else n.frz1
            dp_i = W(frz_b, dp_i + rfreeze_i * cons2 * n.prr, dp_i)
            prr_i = W(frz_b, rfreeze_i * cons2 * dp, 0.0).astype(dtype)
            prr_i = prr_i + cons2 * dp * dr_i
            prs_i = cons2 * dp * dr_i
            dp_i = dp_i + cons2 * (n.prr + n.prs) * dr_i

            # :738-782
            cloudy = n.cloudy
            clc_acc = in_clc_i[k].copy()
            c_prs_i = prs_i - qiwc_i
            c_qiwc_i = qiwc_i + c_prs_i
            qinew_i = -c_prs_i
            c_clc = clc_acc + qinew_i * n.cldi * n.itmp2
            cldi_i = qinew_i * oclc * n.itmp2
            di_i = -qinew_i * oclc * n.cldi * n.itmp2
            itmp4 = ckcodtia if LREGCL else ckcodti
            c_ot_i = ot_i + 0.025 * itmp4 * n.itmp12 * (1.0 - n.itmp11) * di_i
            cldi_i = cldi_i + 2.0 * itmp4 * n.itmp12 * n.itmp11 * n.cldi * di_i / icrit**2.0
            c_qiwc_i = c_qiwc_i + cldi_i / oclc
            c_clc = c_clc - n.qiwc1 * cldi_i / oclc**2.0
            c_prr_i = prr_i - qlwc_i
            c_qlwc_i = qlwc_i + c_prr_i
            qlnew_i = -c_prr_i
            c_clc = c_clc + qlnew_i * n.cldl * n.ltmp2
            cldl_i = qlnew_i * oclc * n.ltmp2
            dl_i = -qlnew_i * oclc * n.cldl * n.ltmp2
            ltmp4 = ckcodtla if LREGCL else ckcodtl
            cldl_i = cldl_i + 2.0 * ltmp4 * n.ltmp1 * n.cldl * dl_i / lcrit**2.0
            c_qlwc_i = c_qlwc_i + cldl_i / oclc
            c_clc = c_clc - n.qlwc1 * cldl_i / oclc**2.0
            qiwc_i = W(cloudy, c_qiwc_i, qiwc_i)
            qlwc_i = W(cloudy, c_qlwc_i, qlwc_i)
            ot_i = W(cloudy, c_ot_i, ot_i)
            clc_acc = W(cloudy, c_clc, clc_acc)

            # :784-806
            melt = n.melt
            cons, snmlt = n.cons, n.snmlt
            snmlt_i = -ot_i / cons + rfln_i - sfln_i
            cons_i = ot_i * snmlt / cons**2.0
            m_rfl_i = rfln_i
            m_sfl_i = sfln_i
            allm = n.sfl <= n.z2s
            m_sfl_i = W(allm, m_sfl_i + snmlt_i, m_sfl_i)
            z2s_i = W(allm, 0.0, snmlt_i).astype(dtype)
            warm2 = t2 > meltp2
            m_ot_i = W(warm2, ot_i + cons * z2s_i, ot_i)
            cons_i = W(warm2, cons_i + (t2 - meltp2) * z2s_i, cons_i)
            m_dp_i = dp_i + cons2 * cons_i / lfdcp
            m_lfdcp_i = -cons2 * dp * cons_i / lfdcp**2.0
            rfl_i3d[k] = W(melt, m_rfl_i, 0.0)
            sfl_i3d[k] = W(melt, m_sfl_i, 0.0)
            rfln_i = W(melt, 0.0, rfln_i).astype(dtype)
            sfln_i = W(melt, 0.0, sfln_i).astype(dtype)
            ot_i = W(melt, m_ot_i, ot_i)
            dp_i = W(melt, m_dp_i, dp_i)
            lfdcp_i = W(melt, m_lfdcp_i, 0.0).astype(dtype)

            # :808-817
            covpclr_i = W(n.covpclr1 < 0.0, 0.0, covpclr_i).astype(dtype)
            covptot_i = covptot_i + covpclr_i
            clc_acc = clc_acc - covpclr_i
            up = oclc > n.covptot
            clc_acc = W(up, clc_acc + covptot_i, clc_acc)
            covptot_i = W(up, 0.0, covptot_i).astype(dtype)
            covptot_i3d[k] = covptot_i

            # :819-825
            qiwc_i = qiwc_i + condi_i / dt
            oqi_i = oqi_i - condi_i / dt
            qlwc_i = qlwc_i + condl_i / dt
            oql_i = oql_i - condl_i / dt
            qc_i = fwat * qlwc_i + (1.0 - fwat) * qiwc_i
            fwat_i = fwat_i + n.qc3 * (qlwc_i - qiwc_i)

            # :827-842
            dqc_i = -qc_i
            lo3 = n.lo3
            dqc_i_reg = dqc_i * 0.1 if LREGCL else dqc_i
            dqsdz_i = W(lo3, dt * dqc_i_reg * (mfd + mfu) * n.fac4, 0.0).astype(dtype)
            omfd_i = W(lo3, dt * dqc_i_reg * n.dqsdz * n.fac4, 0.0).astype(dtype)
            omfu_i = W(lo3, dt * dqc_i_reg * n.dqsdz * n.fac4, 0.0).astype(dtype)
            rho_i = W(lo3, -dqc_i_reg * n.dqc * n.fac4, 0.0).astype(dtype)
            qc_i = W(lo3, qc_i, qc_i + dqc_i)

            # :844-855
            dtdzmo_i = dqsdz_i * n.dqsdtemp
            dqsdtemp_i = dqsdz_i * n.dtdzmo - n.dtdzmo * dtdzmo_i * n.ldcp * n.fac3
            rodqsdp_i = -RG * (dqsdz_i + dtdzmo_i * n.ldcp * n.fac3)
            ldcp_i = -dtdzmo_i * (RG * n.rodqsdp + n.dtdzmo * n.dqsdtemp) * n.fac3
            fwat_i = fwat_i + ldcp_i * (lvdcp - lsdcp)
            lvdcp_i = lvdcp_i + fwat * ldcp_i
            lsdcp_i = lsdcp_i + (1.0 - fwat) * ldcp_i
            rho_i = rho_i - rodqsdp_i * qsk * n.fac2
            oqsat_i = oqsat_i - rodqsdp_i * n.rho * n.fac2
            oap_i = oap_i + (rodqsdp_i * n.rho * qsk * n.fac2**2.0 + rho_i * n.fac1)
            foeew_i = -RETV * rodqsdp_i * n.rho * qsk * n.fac2**2.0
            ot_i = ot_i - rho_i * apk * n.fac1 / t2

            # :857-877
            lude, clc = n.lude, n.clc
            lo1b = (k < NLEV - 1) & (lude >= P["RLMIN"]) & (lu1 >= ZEPS2)
            ex = np.exp(-lude / lu1)
            lude_i = W(lo1b, qc_i + (1.0 - clc) / lu1 * ex * clc_acc, 0.0).astype(dtype)
            dlu_i = W(lo1b, (1.0 - clc) * lude / lu1**2.0 * ex * clc_acc, 0.0).astype(dtype)
            clc_acc = W(lo1b, clc_acc * (1.0 - (1.0 - ex)), clc_acc)
            olude_i = olude_i + dt * gdp * lude_i
            gdp_i = gdp_i + dt * ludek * lude_i
            daph_i = daph_i + RG * gdp_i / (aph[k + 1] - aph[k]) ** 2.0

            # :879-923
            qt, qcrit, qsat, scalm = n.qt, n.qcrit, n.qsat, n.scalm
            b1 = qt < qcrit
            b2 = ~b1 & (qt >= qsat)
            b3 = ~b1 & ~b2
            qpd, qcd, tmp3 = n.qpd, n.qcd, n.tmp3
            den = qcd - scalm * (qt - qcrit)
            p_qpd_i = scalm * qc_i * clc**2.0
            p_qcd_i = (1.0 - scalm) * qc_i * clc**2.0
            p_clc = clc_acc + 2.0 * (scalm * qpd + (1.0 - scalm) * qcd) * clc * qc_i
            if LREGCL:
                rat = qpd / qcd
                yyy = np.minimum(0.3, 3.5 * np.sqrt(rat * (1.0 - scalm * (1.0 - rat)) ** 3.0) / (1.0 - scalm))
                p_clc = p_clc * yyy
            p_qpd_i = p_qpd_i - 0.5 / tmp3 * p_clc / den
            p_qcd_i = p_qcd_i + 0.5 / tmp3 * qpd * p_clc / den**2.0
            p_qt_i = (-0.5 / tmp3 * (qpd * scalm * p_clc) / den**2.0) - p_qpd_i
            p_qcrit_i = (0.5 / tmp3 * (qpd * scalm * p_clc) / den**2.0) - p_qcd_i
            p_qsat_i = p_qcd_i + p_qpd_i
            qt_i = W(b3, p_qt_i, 0.0).astype(dtype)
            qsat_i = W(b1, 0.0, W(b2, (1.0 - scalm) * qc_i, p_qsat_i)).astype(dtype)
            qcrit_i = W(b1, 0.0, W(b2, -(1.0 - scalm) * qc_i, p_qcrit_i)).astype(dtype)
            in_clc_i[k] = 0.0
            oq_i = oq_i + qt_i
            oql_i = oql_i + qt_i
            oqi_i = oqi_i + qt_i

            # :925-938
            qsat_i = qsat_i + qcrit_i * n.crh2
            oqsat_i = oqsat_i + qsat_i * n.supsat
            supsat_i = qsat_i * qsk
            ot_i = W(t2 < P["RTICE"], ot_i - 0.003 * supsat_i, ot_i)
            qclip = n.q2 > qsk
            oqsat_i = W(qclip, oqsat_i + qlim_i, oqsat_i)
            oq_i = W(qclip, oq_i, oq_i + qlim_i)

            # :940-967
            dqsdtemp_i = dqsdtemp_i + cons3 * corqs_i
            oqsat_i = oqsat_i + n.fac * n.cor * dqsdtemp_i
            cor_i = n.fac * qsk * dqsdtemp_i
            fac_i = n.cor * qsk * dqsdtemp_i
            esdp_i = RETV * cor_i * n.cor**2.0
            facw_i = fwat * fac_i
            faci_i = (1.0 - fwat) * fac_i
            fwat_i = fwat_i + (n.facw - n.faci) * fac_i
            ot_i = ot_i - 2.0 * (R5IES * faci_i / (t2 - R4IES) ** 3.0 + R5LES * facw_i / (t2 - R4LES) ** 3.0)
            esdp_i = W(n.esdp1 > P["ZQMAX"], 0.0, esdp_i).astype(dtype)
            foeew_i = foeew_i + esdp_i / apk
            oap_i = oap_i - esdp_i * n.foeew / apk**2.0
            cold = t2 < RTT
            z3es = W(cold, P["R3IES"], P["R3LES"]).astype(dtype)
            z4es = W(cold, R4IES, R4LES).astype(dtype)
            ot_i = ot_i + z3es * (RTT - z4es) * foeew_i * n.foeew / (t2 - z4es) ** 2.0
            ot_i = W(cold, ot_i + 0.545 * 0.17 * fwat_i / np.cosh(0.17 * (t2 - P["RLPTRC"])) ** 2.0, ot_i)

            o_ap_i[k], o_t_i[k], o_q_i[k], o_ql_i[k], o_qi_i[k] = oap_i, ot_i, oq_i, oql_i, oqi_i
            o_qsat_i[k], o_lude_i[k], o_mfd_i[k], o_mfu_i[k] = oqsat_i, olude_i, omfd_i, omfu_i
            daph_i3d[k], dp_i3d[k], dlu_i3d[k] = daph_i, dp_i, dlu_i
            lvdcp_i3d[k], lsdcp_i3d[k], lfdcp_i3d[k] = lvdcp_i, lsdcp_i, lfdcp_i

        # :969-986
        in_fplsl_i[...] = 0.0
        in_fplsn_i[...] = 0.0
        aph_s_i = aph_s_i + (-daph_i3d[nz - 1] + dp_i3d[nz - 1])
        o_aph_i[nz] = aph_s_i
        o_lu_i[nz] = -dlu_i3d[nz - 1]
        for k in range(1, nz):
            o_aph_i[k] = daph_i3d[k] - daph_i3d[k - 1] - dp_i3d[k] + dp_i3d[k - 1]
            o_lu_i[k] = -dlu_i3d[k - 1]
        o_aph_i[0] = daph_i3d[0] - dp_i3d[0]
        o_lu_i[0] = 0.0

        # :988-996
        for k in range(nz):
            zz = RLVTT * lvdcp_i3d[k] + RLSTT * lsdcp_i3d[k] + RLMLT * lfdcp_i3d[k]
            o_q_i[k] = o_q_i[k] + (-zz * RCPD * RVTMP2 / (RCPD + RCPD * RVTMP2 * L[k].qpost) ** 2.0)
            o_supsat_i[k] = dt * o_q_i[k]
            o_cml_t_i[k] = dt * o_t_i[k]
            o_cml_q_i[k] = dt * o_q_i[k]
            o_cml_ql_i[k] = dt * o_ql_i[k]
            o_cml_qi_i[k] = dt * o_qi_i[k]

    tends = {
        "f_t": o_tt, "f_q": o_tq, "f_ql": o_tql, "f_qi": o_tqi,
        "f_cml_t_i": o_cml_t_i, "f_cml_q_i": o_cml_q_i, "f_cml_ql_i": o_cml_ql_i, "f_cml_qi_i": o_cml_qi_i,
    }
    diags = {
        "f_aph_i": o_aph_i, "f_ap_i": o_ap_i, "f_q_i": o_q_i, "f_qsat_i": o_qsat_i, "f_t_i": o_t_i,
        "f_ql_i": o_ql_i, "f_qi_i": o_qi_i, "f_lude_i": o_lude_i, "f_lu_i": o_lu_i, "f_mfu_i": o_mfu_i,
        "f_mfd_i": o_mfd_i, "f_supsat_i": o_supsat_i,
        "f_clc": o_clc, "f_covptot": o_covptot, "f_fhpsl": o_fhpsl, "f_fhpsn": o_fhpsn,
        "f_fplsl": o_fplsl, "f_fplsn": o_fplsn,
    }
    return tends, diags


# ----------------------------------------------------------------------------------------
# validation maths
# ----------------------------------------------------------------------------------------
TAYLOR_TEND_NAMES = ("f_t", "f_q", "f_ql", "f_qi")  # tangent_linear/validation.py:225
TAYLOR_DIAG_NAMES = ("f_clc", "f_fhpsl", "f_fhpsn", "f_fplsl", "f_fplsn", "f_covptot")  # :232


def taylor_field_norm(f2, field_nl, field_nl_p, field_tl):
    """tangent_linear/validation.py:252-261."""
    den = np.abs(f2 * np.sum(field_tl))
    if den > sys.float_info.epsilon:
        return np.abs(np.sum(field_nl_p - field_nl)) / den
    return 0


def taylor_norm(f2, tends_nl, diags_nl, tends_nl_p, diags_nl_p, tends_tl, diags_tl):
    """tangent_linear/validation.py:219-238: mean of the non-zero per-field norms."""
    total_count, total_norm = 0, 0.0
    for name in TAYLOR_TEND_NAMES:
        norm = taylor_field_norm(f2, tends_nl[name], tends_nl_p[name], tends_tl[name + "_i"])
        total_count += norm > 0
        total_norm += norm
    for name in TAYLOR_DIAG_NAMES:
        norm = taylor_field_norm(f2, diags_nl[name], diags_nl_p[name], diags_tl[name + "_i"])
        total_count += norm > 0
        total_norm += norm
    return total_norm / total_count if total_count > 0 else 0


def taylor_score(norms_in):
    """tangent_linear/validation.py:183-217.  Returns (passed, code, start)."""
    norms = np.abs(1 - np.asarray(norms_in, dtype=np.float64))
    start = -1
    for i in range(norms.size):
        if start == -1 and norms[i] < 0.5:
            start = i
    if start == -1 or start > 3:
        return False, 13, start
    test = -10
    negat = 1
    for i in range(start, norms.size - 1):
        tmp_negat = int(norms[i + 1] < norms[i])
        if negat > tmp_negat:
            test += 10
        negat = tmp_negat
    if test == -10:
        test = 11
    if np.min(norms[start:]) > 1e-5:
        test += 7
    if np.min(norms[start:]) > 1e-6:
        test += 5
    return test <= 5, test, start


SYM_TL_TENDS = ("f_t_i", "f_q_i", "f_ql_i", "f_qi_i")  # adjoint/validation.py:170
SYM_TL_DIAGS = ("f_clc_i", "f_fhpsl_i", "f_fhpsn_i", "f_fplsl_i", "f_fplsn_i", "f_covptot_i")  # :176
SYM_AD_TENDS = ("f_cml_t_i", "f_cml_q_i", "f_cml_ql_i", "f_cml_qi_i")  # :188
SYM_AD_DIAGS = (
    "f_ap_i", "f_aph_i", "f_t_i", "f_q_i", "f_qsat_i", "f_ql_i", "f_qi_i", "f_lu_i", "f_lude_i",
    "f_mfd_i", "f_mfu_i", "f_supsat_i",
)  # :195-208


def symmetry_norm1(tends_tl, diags_tl):
    """adjoint/validation.py:167-181 (per column: sum over levels, axis 0 in [nz+1, nx])."""
    out = None
    for name in SYM_TL_TENDS:
        f = tends_tl[name]
        out = np.zeros(f.shape[1]) if out is None else out
        out += np.sum(f**2, axis=0)
    for name in SYM_TL_DIAGS:
        out += np.sum(diags_tl[name] ** 2, axis=0)
    return out


def symmetry_norm2(state_i, tends_ad, diags_ad):
    """adjoint/validation.py:183-215."""
    out = None
    for name in SYM_AD_TENDS:
        a = state_i["f_tnd_" + name[2:]]
        b = tends_ad[name]
        out = np.zeros(a.shape[1]) if out is None else out
        out += np.sum(a * b, axis=0)
    for name in SYM_AD_DIAGS:
        out += np.sum(state_i[name] * diags_ad[name], axis=0)
    return out


def symmetry_norm3(norm1, norm2, dtype):
    """adjoint/validation.py:157-160."""
    eps = np.finfo(dtype).eps
    with _errstate():
        return np.where(norm2 == 0, abs(norm1 - norm2) / eps, abs(norm1 - norm2) / (eps * norm2))
