// CLOUDSC2 level physics: one model level of one column, in the three formulations.
//
//   level_fwd  -- nonlinear trajectory of a level   (reference nonlinear/_stencils/cloudsc2.py
//                 :113-388 with nonlinear/_stencils/cuadjtqs.py:22-68 inlined)
//   level_tl   -- tangent of level_fwd              (tangent_linear/_stencils/cloudsc2.py:149-753,
//                 tangent_linear/_stencils/cuadjtqs.py:22-84)
//   level_ad   -- transpose of level_tl             (adjoint/_stencils/cloudsc2.py:494-967,
//                 adjoint/_stencils/cuadjtqs.py:93-156)
//
// This is a restatement designed for the FP64 pipe of the B200, not a transcription: the NL, TL
// and AD sweeps all evaluate the SAME trajectory function (so TL and AD linearise exactly the
// trajectory NL produces), divisions are folded into shared reciprocals (1/ap, 1/dp, 1/(t-R4xES),
// 1/clc, 1/dt, ...), `x**2.0`/`x**3.0` are products, and everything that depends on the level
// only (scalm, crh2) or on (externals, dt) only comes in precomputed.  The TL statements are the
// derivative of this trajectory including the reference's LREGCL regularisations; the AD
// statements are the exact transpose of the TL statements.
#pragma once

#include "cs2_common.cuh"

namespace cs2 {

template <bool EVAP_, bool TETENS_>
struct Cfg {
  static constexpr bool EVAP = EVAP_;      // LEVAPLS2 or LDRAIN1D
  static constexpr bool TETENS = TETENS_;  // LPHYLIN or LDRAIN1D (NL only; TL/AD always true)
};

// Inputs of one level (names = stencil arguments without the in_ prefix; aph0/aph1 are the
// half levels k and k+1, lu1 is lu at level k+1).  The same struct carries perturbations (TL)
// and adjoints (AD).
template <class R>
struct LevelIn {
  R ap, aph0, aph1, lu1, lude, mfd, mfu, q, qi, ql, qsat, supsat, t, tnd_q, tnd_qi, tnd_ql, tnd_t;
};

// Precipitation fluxes / overlap entering a level (and, on return, leaving it).
template <class R>
struct Carry {
  R rfl, sfl, covptot;
};

template <class R>
struct LevelOut {
  R clc, covptot, tnd_q, tnd_qi, tnd_ql, tnd_t;
};

// One Newton step of the saturation adjustment with everything its tangent / adjoint needs.
template <class R>
struct AdjStep {
  R t, q;               // state entering the step
  R rt;                 // 1 / (t - z4es)
  R foeew, qsc, cor, qs, z2s, rden, cond;
  bool clipped;         // foeew / ap > ZQMAX
  R ct, cap;            // local Jacobian (adj_step_coef): cond_i = rden q_i + ct t_i + cap ap_i
};

// Trajectory of a level.  Everything is a plain local; members that a caller does not read
// are eliminated by the compiler.
template <class R>
struct Traj {
  R t0, q0, ql0, qi0, dp, rdp, rap;
  R zzinv, lfdcp, lsdcp, lvdcp, rlfdcp;
  bool cold, ice, clip_esdp;
  R tp1, fwat, z3es, z4es, rtw, rti, rtm4, foeew, facw, faci, fac, cor, dqsdtemp, corqs, qlim;
  R scalm, crh2, supsat, qsat, qcrit, qt;
  int branch;  // 0: qt < qcrit, 1: qt >= qsat, 2: partial cloud
  R qpd, qcd, rden, tmp3, clc, qc1;
  R gdp, lude;
  bool lo1;
  R rlu1, ex, clc_o, qc2;
  R fac1, rho, fac2, rodqsdp, ldcp, fac3, dtdzmo, dqsdz, fac4, mfsum;
  bool lo3;
  R dqc, qc3, qlwc1, qiwc1, condl1, condi1;
  R covptotp, covptot1, covpclr1, covpclr;
  bool melt, allm, warm2;
  R cons, rcons, snmlt, tmelt;
  bool cloudy;
  R rclc, cldl, ltmp1, ltmp2, cldi, itmp11, itmp12, itmp2, prr, prs, qlwc, qiwc;
  bool frz1;
  R rfreeze1;
  R evapr, evaps;
  // precipitation-evaporation branch (only with Cfg::EVAP; read by level_ad)
  bool ev, ev_capped, ev_gone;
  R ev_prtot, ev_rfln, ev_sfln, ev_preclr, ev_qe, ev_beta, ev_b, ev_dtgdp, ev_dpr, covptot_out;
  R t3, qa;
  bool warmc;
  R z3c, z4c, z5c, zalc;
  AdjStep<R> sb, sa;  // first and second Newton step
  R tpost, qpost;
  bool pos, frz2;
  R dq, dr2, rfreeze3, condl2, condi2;
};

// Transcendental results of a level (the expensive part of the trajectory).  MODE 0: compute;
// MODE 1: compute and record (AD forward sweep, CS2_AD_CHECKPOINT); MODE 2: replay recorded values
// (AD backward sweep, CS2_AD_CHECKPOINT) -- the cheap algebra around them is always recomputed.
enum { CK_TP1, CK_FOEEW, CK_LTMP1, CK_LTMP2, CK_ITMP11, CK_ITMP12, CK_ITMP2, CK_SB, CK_SA, CK_N };

template <class R, int MODE>
struct Trans {
  R v[CK_N];
  CS2_HD void defaults() {
    v[CK_TP1] = R(2);
#pragma unroll
    for (int n = 1; n < CK_N; ++n) v[n] = R(1);
  }
  CS2_HD R exp(int idx, R x) {
    if (MODE == 2) return v[idx];
    const R e = exp_(x);
    if (MODE == 1) v[idx] = e;
    return e;
  }
  // N independent exponentials at once (checkpoint slots idx[n]): evaluated in lockstep on the device (exp_batch)
  template <int N>
  CS2_HD void exps(const int (&idx)[N], const R (&x)[N], R (&e)[N]) {
    if (MODE == 2) {
#pragma unroll
      for (int n = 0; n < N; ++n) e[n] = v[idx[n]];
      return;
    }
    exp_batch<N>(x, e);
    if (MODE == 1) {
#pragma unroll
      for (int n = 0; n < N; ++n) v[idx[n]] = e[n];
    }
  }
  CS2_HD R tp1(R x) {
    if (MODE == 2) return v[CK_TP1];
    const R e = one_plus_tanh<R>(x);
    if (MODE == 1) v[CK_TP1] = e;
    return e;
  }
};

// ---------------------------------------------------------------------------------------
// thermodynamic functions (common/_stencils/fcttre.py:22-57) -- only the non-LPHYLIN NL path
// ---------------------------------------------------------------------------------------
template <class R>
CS2_HD R foealfa(const DevParams<R>& p, R t) {
  const R x = (max_(p.RTICE, min_(p.RTWAT, t)) - p.RTICE) * p.RTWAT_RTICE_R;
  return min_(R(1), x * x);
}
template <class R>
CS2_HD R foealfcu(const DevParams<R>& p, R t) {
  const R x = (max_(p.RTICECU, min_(p.RTWAT, t)) - p.RTICECU) * p.RTWAT_RTICECU_R;
  return min_(R(1), x * x);
}
template <class R>
CS2_HD R foeew_mixed(const DevParams<R>& p, R t, R alfa) {
  const R el = exp_(p.R3LES * (t - p.RTT) / (t - p.R4LES));
  const R ei = exp_(p.R3IES * (t - p.RTT) / (t - p.R4IES));
  return p.R2ES * (alfa * el + (R(1) - alfa) * ei);
}

// saturation stencil, one point (common/_stencils/saturation.py:30-42).  The exponential of a phase whose
// weight is exactly 0 is skipped (alfa is 0 below RTICE and 1 above RTWAT: 0 * finite + x == x bit for bit).
template <class R>
CS2_HD R saturation_point(const DevParams<R>& p, bool lphylin, R ap, R t) {
  const R alfa = (lphylin || p.kflag != 1) ? foealfa(p, t) : foealfcu(p, t);
  const R dtt = t - p.RTT;
  R el = R(0), ei = R(0);
  if (alfa > R(0)) el = exp_(p.R3LES * dtt * rcp(t - p.R4LES));
  if (alfa < R(1)) ei = exp_(p.R3IES * dtt * rcp(t - p.R4IES));
  R foeew;
  if (lphylin)
    foeew = alfa * (p.R2ES * el) + (R(1) - alfa) * (p.R2ES * ei);
  else
    foeew = p.R2ES * (alfa * el + (R(1) - alfa) * ei);
  const R qs = min_(foeew * rcp(ap), p.QMAX);
  return qs * rcp(R(1) - p.RETV * qs);
}

// N points at once (the saturation kernel): same statements as saturation_point, the exponentials of the N points in
// lockstep per phase (a phase no point needs is skipped: at one level the columns of a thread are all cold, all warm or
// all mixed), and qs / (1 - RETV qs) = foeew / (ap - RETV foeew) where the clip does not bind -- one reciprocal for the
// last two divisions (saturation.py:35,42; the clip test foeew / ap > QMAX is foeew > QMAX ap, ap > 0).
template <class R, int N>
CS2_HD void saturation_points(const DevParams<R>& p, bool lphylin, const R (&ap)[N], const R (&t)[N], R (&qsat)[N]) {
  R alfa[N], xl[N], xi[N], el[N], ei[N];
  bool need_l = false, need_i = false;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    alfa[n] = (lphylin || p.kflag != 1) ? foealfa(p, t[n]) : foealfcu(p, t[n]);
    const R dtt = t[n] - p.RTT;
    xl[n] = p.R3LES * dtt * rcp(t[n] - p.R4LES);
    xi[n] = p.R3IES * dtt * rcp(t[n] - p.R4IES);
    need_l = need_l || (alfa[n] > R(0));
    need_i = need_i || (alfa[n] < R(1));
  }
#pragma unroll
  for (int n = 0; n < N; ++n) el[n] = ei[n] = R(0);
  if (need_l) exp_batch<N>(xl, el);
  if (need_i) exp_batch<N>(xi, ei);
#pragma unroll
  for (int n = 0; n < N; ++n) {
    // a phase with weight exactly 0 contributes an exact 0, as in saturation_point (alfa is 0 below RTICE, 1 above RTWAT)
    const R wl = (alfa[n] > R(0)) ? el[n] : R(0), wi = (alfa[n] < R(1)) ? ei[n] : R(0);
    R foeew;
    if (lphylin)
      foeew = alfa[n] * (p.R2ES * wl) + (R(1) - alfa[n]) * (p.R2ES * wi);
    else
      foeew = p.R2ES * (alfa[n] * wl + (R(1) - alfa[n]) * wi);
    const bool clip = foeew > p.QMAX * ap[n];
    qsat[n] = clip ? p.qsat_clip : foeew * rcp(ap[n] - p.RETV * foeew);
  }
}

// Local Jacobian of one Newton step, formed from the trajectory alone (pre-accumulation):
//   cond_i = rden q_i + ct t_i + cap ap_i
// -- the statements of tangent_linear/_stencils/cuadjtqs.py:22-55 collected by input (with 1 + RETV qs = cor), whose
// transpose is adjoint/_stencils/cuadjtqs.py:93-124.  The adjoint of the step is then
//   a_cond = zal a_t - a_q;  a_q += rden a_cond;  a_t += ct a_cond;  a_ap += cap a_cond
// i.e. four dependent FMAs on the backward sweep's critical path instead of a chain of a dozen products; the coefficients are
// evaluated by the forward recomputation (adj_step<LIN>), off that path.  Measured: AD 1.180 -> 1.138 ms at 65 536 columns
// together with the shared products in level_ad (profiles/r2q_ad_preaccumulation.md).
template <class R>
CS2_HD void adj_step_coef(const DevParams<R>& p, R rap, R r5, const AdjStep<R>& s, R& ct, R& cap) {
  const R h = s.cond * s.z2s;
  const R e = s.rden * (s.cor * s.cor) * (R(1) + h * (s.cor + p.RETV * s.qs));  // -d cond / d qsc
  const R f = R(2) * s.rden * h * s.qs * s.cor * s.rt;                            // d cond / d t through z2s
  const R fr = s.clipped ? R(0) : s.foeew * rap;
  ct = f - e * (fr * r5 * (s.rt * s.rt));
  cap = e * fr * rap;
}

// ---------------------------------------------------------------------------------------
// saturation adjustment step (nonlinear/_stencils/cuadjtqs.py:22-35)
// ---------------------------------------------------------------------------------------
// LIN = the caller linearises about this trajectory (TL / AD): keep cor, qs and rden.  The NL kernels (LIN = false)
// use an algebraically merged form with one reciprocal on the critical path instead of two in sequence.
template <class R, bool LIN, class X>
CS2_HD void adj_step(const DevParams<R>& p, R rap, R z3, R z4, R z5, R zal, R& t, R& q, AdjStep<R>& s, X& x, int ck) {
  s.ct = s.cap = R(0);
  s.t = t;
  s.q = q;
  s.rt = rcp(t - z4);
  s.foeew = p.R2ES * x.exp(ck, z3 * (t - p.RTT) * s.rt);
  const R qs1 = s.foeew * rap;
  s.clipped = qs1 > p.ZQMAX;
  s.qsc = s.clipped ? p.ZQMAX : qs1;
  s.z2s = z5 * s.rt * s.rt;
  if (LIN) {
    // the two reciprocals side by side (cor feeds rden in the literal form): rden = 1 / (1 + qsc cor^2 z2s) = a^2 / (a^2 + qsc z2s),
    // a = 1 - RETV qsc = 1 / cor.  adj_step_fwd_tl (cs2_physics_tl.cuh) has the same statements: TL and AD trajectories stay
    // bit-identical
    const R a = R(1) - p.RETV * s.qsc;
    s.cor = rcp(a);
    const R a2 = a * a;
    s.rden = a2 * rcp(a2 + s.qsc * s.z2s);
    s.qs = s.qsc * s.cor;
    s.cond = (q - s.qs) * s.rden;
    adj_step_coef(p, rap, z3 * (p.RTT - z4), s, s.ct, s.cap);
  } else {
    // cond = (q - qsc/a) / (1 + qsc z2s / a^2) = a (q a - qsc) / (a^2 + qsc z2s),  a = 1 - RETV qsc
    const R a = R(1) - p.RETV * s.qsc;
    s.cond = a * (q * a - s.qsc) * rcp(a * a + s.qsc * s.z2s);
    s.cor = s.qs = s.rden = R(0);
  }
  t = t + zal * s.cond;
  q = q - s.cond;
}

// tangent of adj_step (tangent_linear/_stencils/cuadjtqs.py:22-55)
template <class R>
CS2_HD void adj_step_tl(const DevParams<R>& p, R rap, R ap_i, R z3, R z4, R z5, R zal, const AdjStep<R>& s,
                        R& t_i, R& q_i) {
  const R foeew_i = s.foeew * z3 * (p.RTT - z4) * t_i * s.rt * s.rt;
  R qsc_i = -ap_i * rap * rap * s.foeew + rap * foeew_i;
  if (s.clipped) qsc_i = R(0);
  const R cor_i = p.RETV * qsc_i * s.cor * s.cor;
  const R qs_i = qsc_i * s.cor + s.qsc * cor_i;
  const R z2s_i = R(-2) * z5 * t_i * s.rt * s.rt * s.rt;
  const R cond_i = (q_i - qs_i) * s.rden -
                   (s.q - s.qs) * (qs_i * s.cor * s.z2s + s.qs * cor_i * s.z2s + s.qs * s.cor * z2s_i) * s.rden * s.rden;
  t_i += zal * cond_i;
  q_i -= cond_i;
}

// ---------------------------------------------------------------------------------------
// level_fwd
//   c       : in = fluxes / overlap entering the level, out = leaving it
//   conv_ok : k < nlev-1 (the level below exists, so lu[k+1] is a physical value; the TL and
//             AD stencils carry this guard explicitly, tangent_linear/_stencils/cloudsc2.py:317)
//   ad_ref  : second freezing test on the pre-adjustment temperature (AD stencil literal,
//             adjoint/_stencils/cloudsc2.py:427); false for NL / TL / consistent AD.
// ---------------------------------------------------------------------------------------
template <class R, class C, bool LIN, class X>
CS2_HD void level_fwd(const DevParams<R>& p, const LevelIn<R>& in, R scalm, R crh2, bool conv_ok, R aph_s,
                      bool ad_ref, Carry<R>& c, LevelOut<R>& o, Traj<R>& tr, X& x) {
  const R one = R(1), zero = R(0);
  // first guess (:104,115-117)
  tr.t0 = in.t + p.dt * in.tnd_t;
  tr.q0 = in.q + p.dt * in.tnd_q + in.supsat;
  tr.ql0 = in.ql + p.dt * in.tnd_ql;
  tr.qi0 = in.qi + p.dt * in.tnd_qi;
  tr.scalm = scalm;
  tr.crh2 = crh2;
  const R t0 = tr.t0;

  // thermodynamic constants (:130-134)
  tr.dp = in.aph1 - in.aph0;
  tr.rdp = rcp(tr.dp);
  tr.rap = rcp(in.ap);
  if (p.rvtmp2_zero) {
    tr.zzinv = p.rcpd;
    tr.lfdcp = p.lfdcp0;
    tr.lsdcp = p.lsdcp0;
    tr.lvdcp = p.lvdcp0;
    tr.rlfdcp = p.rlfdcp0;
  } else {
    tr.zzinv = rcp(p.RCPD + p.RCPD * p.RVTMP2 * tr.q0);
    tr.lfdcp = p.RLMLT * tr.zzinv;
    tr.lsdcp = p.RLSTT * tr.zzinv;
    tr.lvdcp = p.RLVTT * tr.zzinv;
    tr.rlfdcp = rcp(tr.lfdcp);
  }

  // dqs/dT correction factor (:141-160)
  tr.rtw = rcp(t0 - p.R4LES);
  tr.rti = rcp(t0 - p.R4IES);
  tr.cold = t0 < p.RTT;
  if (C::TETENS) {
    if (tr.cold) {
      tr.tp1 = x.tp1(R(0.17) * (t0 - p.RLPTRC));
      tr.fwat = R(0.545) * tr.tp1;
      tr.z3es = p.R3IES;
      tr.z4es = p.R4IES;
      tr.rtm4 = tr.rti;
    } else {
      tr.tp1 = R(2);
      tr.fwat = one;
      tr.z3es = p.R3LES;
      tr.z4es = p.R4LES;
      tr.rtm4 = tr.rtw;
    }
    tr.foeew = p.R2ES * x.exp(CK_FOEEW, tr.z3es * (t0 - p.RTT) * tr.rtm4);
    tr.clip_esdp = tr.foeew * tr.rap > p.ZQMAX;  // esdp = min(foeew / ap, ZQMAX) binds
  } else {
    tr.tp1 = R(2);
    tr.fwat = foealfa(p, t0);
    tr.foeew = foeew_mixed(p, t0, tr.fwat);
    tr.z3es = p.R3LES;
    tr.z4es = p.R4LES;
    tr.rtm4 = tr.rtw;
    tr.clip_esdp = false;
  }
  tr.facw = p.R5LES * tr.rtw * tr.rtw;
  tr.faci = p.R5IES * tr.rti * tr.rti;
  tr.fac = tr.fwat * tr.facw + (one - tr.fwat) * tr.faci;
  // cor = 1 / (1 - RETV esdp); with esdp = foeew / ap unclipped this is ap / (ap - RETV foeew) = ap * fac2,
  // and fac2 is needed by the subsidence term anyway: one reciprocal less per level
  tr.fac2 = rcp(in.ap - p.RETV * tr.foeew);
  tr.cor = tr.clip_esdp ? p.cor_clip : in.ap * tr.fac2;
  tr.dqsdtemp = tr.fac * tr.cor * in.qsat;
  tr.corqs = one + p.cons3 * tr.dqsdtemp;
  tr.qlim = min_(tr.q0, in.qsat);

  // ice supersaturation, critical humidity (:188-193)
  tr.ice = t0 < p.RTICE;
  tr.supsat = tr.ice ? (R(1.8) - R(0.003) * t0) : one;
  tr.qsat = in.qsat * tr.supsat;
  tr.qcrit = crh2 * tr.qsat;

  // uniform total-water distribution (:196-207)
  tr.qt = tr.q0 + tr.ql0 + tr.qi0;
  if (tr.qt < tr.qcrit) {
    tr.branch = 0;
    tr.clc = zero;
    tr.qc1 = zero;
    tr.qpd = tr.qcd = tr.rden = tr.tmp3 = zero;
  } else if (tr.qt >= tr.qsat) {
    tr.branch = 1;
    tr.clc = one;
    tr.qc1 = (one - scalm) * (tr.qsat - tr.qcrit);
    tr.qpd = tr.qcd = tr.rden = tr.tmp3 = zero;
  } else {
    tr.branch = 2;
    tr.qpd = tr.qsat - tr.qt;
    tr.qcd = tr.qsat - tr.qcrit;
    tr.rden = rcp(tr.qcd - scalm * (tr.qt - tr.qcrit));
    tr.tmp3 = sqrt_(tr.qpd * tr.rden);
    tr.clc = one - tr.tmp3;
    tr.qc1 = (scalm * tr.qpd + (one - scalm) * tr.qcd) * (tr.clc * tr.clc);
  }

  // convective component (:210-215)
  tr.gdp = p.RG * tr.rdp;
  tr.lude = p.dt * in.lude * tr.gdp;
  tr.lo1 = conv_ok && (tr.lude >= p.RLMIN) && (in.lu1 >= p.ZEPS2);
  if (tr.lo1) {
    tr.rlu1 = rcp(in.lu1);
    tr.ex = exp_(-tr.lude * tr.rlu1);
    tr.clc_o = tr.clc + (one - tr.clc) * (one - tr.ex);
    tr.qc2 = tr.qc1 + tr.lude;
  } else {
    tr.rlu1 = zero;
    tr.ex = one;
    tr.clc_o = tr.clc;
    tr.qc2 = tr.qc1;
  }

  // compensating subsidence (:218-224)
  tr.fac1 = rcp(p.RD * t0);
  tr.rho = in.ap * tr.fac1;
  tr.rodqsdp = -tr.rho * in.qsat * tr.fac2;
  tr.ldcp = tr.fwat * tr.lvdcp + (one - tr.fwat) * tr.lsdcp;
  tr.fac3 = rcp(one + tr.ldcp * tr.dqsdtemp);
  tr.dtdzmo = p.RG * (p.rcpd - tr.ldcp * tr.rodqsdp) * tr.fac3;
  tr.dqsdz = tr.dqsdtemp * tr.dtdzmo - p.RG * tr.rodqsdp;
  tr.fac4 = p.RD * t0 * tr.rap;  // 1 / rho
  tr.mfsum = in.mfu + in.mfd;
  const R sub = p.dt * tr.dqsdz * tr.mfsum * tr.fac4;
  tr.lo3 = sub < tr.qc2;
  tr.dqc = tr.lo3 ? sub : tr.qc2;
  tr.qc3 = tr.qc2 - tr.dqc;

  // new liquid / ice and condensation rates (:227-230)
  tr.qlwc1 = tr.qc3 * tr.fwat;
  tr.qiwc1 = tr.qc3 * (one - tr.fwat);
  tr.condl1 = (tr.qlwc1 - tr.ql0) * p.rdt;
  tr.condi1 = (tr.qiwc1 - tr.qi0) * p.rdt;

  // maximum overlap (:234-235)
  tr.covptotp = c.covptot;
  tr.covptot1 = max_(c.covptot, tr.clc_o);
  tr.covpclr1 = tr.covptot1 - tr.clc_o;
  tr.covpclr = max_(tr.covpclr1, zero);
  R covptot = tr.covptot1;

  // melting of incoming snow (:238-246)
  R rfln = c.rfl, sfln = c.sfl;
  tr.melt = c.sfl != zero;
  tr.tmelt = t0;
  tr.cons = tr.rcons = tr.snmlt = zero;
  tr.allm = tr.warm2 = false;
  if (tr.melt) {
    tr.cons = p.cons2 * tr.dp * tr.rlfdcp;
    tr.rcons = tr.lfdcp * p.rgdt * tr.rdp;
    tr.warm2 = t0 > p.meltp2;
    const R z2s = tr.warm2 ? tr.cons * (t0 - p.meltp2) : zero;
    tr.allm = c.sfl <= z2s;
    tr.snmlt = tr.allm ? c.sfl : z2s;
    rfln = c.rfl + tr.snmlt;
    sfln = c.sfl - tr.snmlt;
    tr.tmelt = t0 - tr.snmlt * tr.rcons;
  }

  // autoconversion of cloud liquid and ice (:249-272)
  tr.cloudy = tr.clc_o > p.ZEPS2;
  if (tr.cloudy) {
    tr.rclc = rcp(tr.clc_o);
    tr.cldl = tr.qlwc1 * tr.rclc;
    const R xl = tr.cldl * p.rlcrit;
    tr.cldi = tr.qiwc1 * tr.rclc;
    const R xi = tr.cldi * p.ricrit;
    // the five exponentials of the block form two groups of independent ones: evaluated in lockstep
    {
      const int i3[3] = {CK_LTMP1, CK_ITMP11, CK_ITMP12};
      const R a3[3] = {-(xl * xl), -(xi * xi), R(0.025) * (tr.tmelt - p.RTT)};
      R e3[3];
      x.template exps<3>(i3, a3, e3);
      tr.ltmp1 = e3[0];
      tr.itmp11 = e3[1];
      tr.itmp12 = e3[2];
      const int i2[2] = {CK_LTMP2, CK_ITMP2};
      const R a2[2] = {-(p.ckcodtl * (one - tr.ltmp1)), -(p.ckcodti * tr.itmp12 * (one - tr.itmp11))};
      R e2[2];
      x.template exps<2>(i2, a2, e2);
      tr.ltmp2 = e2[0];
      tr.itmp2 = e2[1];
    }
    tr.qlwc = tr.clc_o * tr.cldl * tr.ltmp2;
    tr.prr = tr.qlwc1 - tr.qlwc;
    tr.qiwc = tr.clc_o * tr.cldi * tr.itmp2;
    tr.prs = tr.qiwc1 - tr.qiwc;
  } else {
    tr.rclc = tr.cldl = tr.cldi = zero;
    tr.ltmp1 = tr.ltmp2 = tr.itmp11 = tr.itmp12 = tr.itmp2 = one;
    tr.prr = tr.prs = zero;
    tr.qlwc = tr.qlwc1;
    tr.qiwc = tr.qiwc1;
  }

  // new precipitation and its phase (:275-285)
  const R dr1 = p.cons2 * tr.dp * (tr.prr + tr.prs);
  tr.frz1 = tr.tmelt < p.RTT;
  if (tr.frz1) {
    tr.rfreeze1 = p.cons2 * tr.dp * tr.prr;
    sfln += dr1;
  } else {
    tr.rfreeze1 = zero;
    rfln += dr1;
  }

  // precipitation evaporation (:288-321) -- dead unless LEVAPLS2 or LDRAIN1D
  tr.evapr = tr.evaps = zero;
  tr.ev = tr.ev_capped = tr.ev_gone = false;
  o.covptot = zero;
  if (C::EVAP) {
    const R prtot = rfln + sfln;
    tr.ev_prtot = prtot;
    tr.ev_rfln = rfln;
    tr.ev_sfln = sfln;
    if (prtot > p.ZEPS2 && tr.covpclr > p.ZEPS2) {
      tr.ev = true;
      R preclr = prtot * tr.covpclr / covptot;
      tr.ev_preclr = preclr;
      const R omc = one - tr.clc_o;
      const R qe = in.qsat - (in.qsat - tr.qlim) * tr.covpclr / (omc * omc);
      const R beta =
          p.RG * p.RPECONS * pow_(sqrt_(in.ap / aph_s) / R(0.00509) * preclr / tr.covpclr, R(0.5777));
      const R b = p.dt * beta * (in.qsat - qe) / (one + p.dt * beta * tr.corqs);
      const R dtgdp = p.dt * p.RG * tr.rdp;
      const R dpr1 = tr.covpclr * b / dtgdp;
      tr.ev_capped = dpr1 > preclr;
      const R dpr = min_(dpr1, preclr);
      tr.ev_qe = qe;
      tr.ev_beta = beta;
      tr.ev_b = b;
      tr.ev_dtgdp = dtgdp;
      tr.ev_dpr = dpr;
      preclr -= dpr;
      tr.ev_gone = preclr <= zero;
      if (tr.ev_gone) covptot = tr.clc_o;
      o.covptot = covptot;
      tr.evapr = dpr * rfln / prtot;
      rfln -= tr.evapr;
      tr.evaps = dpr * sfln / prtot;
      sfln -= tr.evaps;
    }
  }
  tr.covptot_out = covptot;

  // first-guess T and q (:328-344)
  R dqdt, dtdt;
  if (C::EVAP) {
    dqdt = -(tr.condl1 + tr.condi1) + (in.lude + tr.evapr + tr.evaps) * tr.gdp;
    dtdt = tr.lvdcp * tr.condl1 + tr.lsdcp * tr.condi1 -
           (tr.lvdcp * tr.evapr + tr.lsdcp * tr.evaps + in.lude * tr.ldcp - (tr.lsdcp - tr.lvdcp) * tr.rfreeze1) * tr.gdp;
  } else {
    dqdt = -(tr.condl1 + tr.condi1) + in.lude * tr.gdp;
    dtdt = tr.lvdcp * tr.condl1 + tr.lsdcp * tr.condi1 -
           (in.lude * tr.ldcp - (tr.lsdcp - tr.lvdcp) * tr.rfreeze1) * tr.gdp;
  }
  tr.t3 = tr.tmelt + p.dt * dtdt;
  tr.qa = tr.q0 + p.dt * dqdt;

  // saturation adjustment, two Newton steps (:347; cuadjtqs.py:38-68)
  tr.warmc = tr.t3 > p.RTT;
  tr.z3c = tr.warmc ? p.R3LES : p.R3IES;
  tr.z4c = tr.warmc ? p.R4LES : p.R4IES;
  tr.z5c = tr.warmc ? p.R5ALVCP : p.R5ALSCP;
  tr.zalc = tr.warmc ? p.RALVDCP : p.RALSDCP;
  R t = tr.t3, q = tr.qa;
  adj_step<R, LIN>(p, tr.rap, tr.z3c, tr.z4c, tr.z5c, tr.zalc, t, q, tr.sb, x, CK_SB);
  adj_step<R, LIN>(p, tr.rap, tr.z3c, tr.z4c, tr.z5c, tr.zalc, t, q, tr.sa, x, CK_SA);
  tr.tpost = t;
  tr.qpost = q;

  // rain fraction and freezing after the adjustment (:350-364)
  tr.pos = tr.qa >= tr.qpost;
  tr.dq = tr.pos ? (tr.qa - tr.qpost) : zero;
  tr.dr2 = p.cons2 * tr.dp * tr.dq;
  tr.frz2 = (ad_ref ? tr.t3 : tr.tpost) < p.RTT;
  tr.condl2 = tr.condl1;
  tr.condi2 = tr.condi1;
  tr.rfreeze3 = tr.rfreeze1;
  if (tr.frz2) {
    tr.rfreeze3 += tr.fwat * tr.dr2;
    tr.condi2 += tr.dq * p.rdt;
    sfln += tr.dr2;
  } else {
    tr.condl2 += tr.dq * p.rdt;
    rfln += tr.dr2;
  }

  // outputs (:367-388)
  o.clc = tr.clc_o;
  if (C::EVAP) {
    o.tnd_q = -(tr.condl2 + tr.condi2) + (in.lude + tr.evapr + tr.evaps) * tr.gdp;
    o.tnd_t = tr.lvdcp * tr.condl2 + tr.lsdcp * tr.condi2 -
              (tr.lvdcp * tr.evapr + tr.lsdcp * tr.evaps + in.lude * tr.ldcp - (tr.lsdcp - tr.lvdcp) * tr.rfreeze3) *
                  tr.gdp;
  } else {
    o.tnd_q = -(tr.condl2 + tr.condi2) + in.lude * tr.gdp;
    o.tnd_t = tr.lvdcp * tr.condl2 + tr.lsdcp * tr.condi2 -
              (in.lude * tr.ldcp - (tr.lsdcp - tr.lvdcp) * tr.rfreeze3) * tr.gdp;
  }
  o.tnd_ql = (tr.qlwc - tr.ql0) * p.rdt;
  o.tnd_qi = (tr.qiwc - tr.qi0) * p.rdt;
  c.rfl = rfln;
  c.sfl = sfln;
  c.covptot = covptot;
}

// ---------------------------------------------------------------------------------------
// level_tl: tangent of level_fwd about the trajectory `tr` (evaporation branch off).
//   d  : perturbations of the level inputs;  ci : perturbation carry (in/out);  oi: outputs.
// ---------------------------------------------------------------------------------------
template <class R>
CS2_HD void level_tl(const DevParams<R>& p, const LevelIn<R>& in, const LevelIn<R>& d, const Traj<R>& tr,
                     Carry<R>& ci, LevelOut<R>& oi) {
  const R one = R(1), zero = R(0);
  const R scalm = tr.scalm;
  R t_i = d.t + p.dt * d.tnd_t;
  R q_i = d.q + p.dt * d.tnd_q + d.supsat;
  const R ql_i = d.ql + p.dt * d.tnd_ql;
  const R qi_i = d.qi + p.dt * d.tnd_qi;
  const R dp_i = d.aph1 - d.aph0;

  R lfdcp_i = zero, lsdcp_i = zero, lvdcp_i = zero;
  if (!p.rvtmp2_zero) {
    const R zz_i = -p.RCPD * p.RVTMP2 * q_i * tr.zzinv * tr.zzinv;
    lfdcp_i = p.RLMLT * zz_i;
    lsdcp_i = p.RLSTT * zz_i;
    lvdcp_i = p.RLVTT * zz_i;
  }

  // dqs/dT correction factor (TL :188-222)
  const R fwat_i = tr.cold ? R(0.545) * R(0.17) * t_i * (tr.tp1 * (R(2) - tr.tp1)) : zero;
  const R foeew_i = tr.z3es * (p.RTT - tr.z4es) * t_i * tr.foeew * tr.rtm4 * tr.rtm4;
  R esdp_i = foeew_i * tr.rap - tr.foeew * d.ap * tr.rap * tr.rap;
  if (tr.clip_esdp) esdp_i = zero;
  const R facw_i = R(-2) * p.R5LES * t_i * tr.rtw * tr.rtw * tr.rtw;
  const R faci_i = R(-2) * p.R5IES * t_i * tr.rti * tr.rti * tr.rti;
  const R fac_i = fwat_i * (tr.facw - tr.faci) + tr.fwat * facw_i + (one - tr.fwat) * faci_i;
  const R cor_i = p.RETV * esdp_i * tr.cor * tr.cor;
  const R dqsdtemp_i = fac_i * tr.cor * in.qsat + tr.fac * cor_i * in.qsat + tr.fac * tr.cor * d.qsat;

  // critical humidity (TL :255-265)
  const R supsat_i = tr.ice ? R(-0.003) * t_i : zero;
  const R qsat_i = d.qsat * tr.supsat + in.qsat * supsat_i;
  const R qcrit_i = tr.crh2 * qsat_i;

  // cloud fraction and condensate (TL :267-306)
  const R qt_i = q_i + ql_i + qi_i;
  R clc_i = zero, qc_i = zero;
  if (tr.branch == 1) {
    qc_i = (one - scalm) * (qsat_i - qcrit_i);
  } else if (tr.branch == 2) {
    const R qpd_i = qsat_i - qt_i;
    const R qcd_i = qsat_i - qcrit_i;
    const R den = tr.qcd - scalm * (tr.qt - tr.qcrit);
    clc_i = R(-0.5) * rcp(tr.tmp3) * (qpd_i * den - tr.qpd * (qcd_i - scalm * (qt_i - qcrit_i))) * tr.rden * tr.rden;
    if (p.lregcl) {
      const R rat = tr.qpd * rcp(tr.qcd);
      const R u = one - scalm * (one - rat);
      const R yyy = min_(R(0.3), R(3.5) * sqrt_(rat * (u * u * u)) * rcp(one - scalm));
      clc_i *= yyy;
    }
    const R wq = scalm * tr.qpd + (one - scalm) * tr.qcd;
    qc_i = (scalm * qpd_i + (one - scalm) * qcd_i) * (tr.clc * tr.clc) + R(2) * wq * tr.clc * clc_i;
  }

  // convective component (TL :308-325)
  const R gdp_i = -p.RG * dp_i * tr.rdp * tr.rdp;
  const R lude_i = p.dt * (d.lude * tr.gdp + in.lude * gdp_i);
  if (tr.lo1) {
    clc_i += -clc_i * (one - tr.ex) +
             (one - tr.clc) * tr.ex * (lude_i * tr.rlu1 - tr.lude * d.lu1 * tr.rlu1 * tr.rlu1);
    qc_i += lude_i;
  }

  // subsidence (TL :327-373)
  const R rho_i = (d.ap - in.ap * t_i * (p.RD * tr.fac1)) * tr.fac1;
  const R rodqsdp_i =
      (-rho_i * in.qsat - tr.rho * d.qsat + tr.rho * in.qsat * (d.ap - p.RETV * foeew_i) * tr.fac2) * tr.fac2;
  const R ldcp_i = fwat_i * (tr.lvdcp - tr.lsdcp) + tr.fwat * lvdcp_i + (one - tr.fwat) * lsdcp_i;
  const R dtdzmo_i = -(p.RG * (ldcp_i * tr.rodqsdp + tr.ldcp * rodqsdp_i) +
                       tr.dtdzmo * (ldcp_i * tr.dqsdtemp + tr.ldcp * dqsdtemp_i)) *
                     tr.fac3;
  const R dqsdz_i = dqsdtemp_i * tr.dtdzmo + tr.dqsdtemp * dtdzmo_i - p.RG * rodqsdp_i;
  R dqc_i;
  if (tr.lo3) {
    dqc_i = (p.dt * (dqsdz_i * tr.mfsum + tr.dqsdz * (d.mfu + d.mfd)) - tr.dqc * rho_i) * tr.fac4;
    if (p.lregcl) dqc_i *= R(0.1);
  } else {
    dqc_i = qc_i;
  }
  qc_i -= dqc_i;

  // liquid / ice and condensation rates (TL :375-386)
  R qlwc_i = qc_i * tr.fwat + tr.qc3 * fwat_i;
  R qiwc_i = qc_i * (one - tr.fwat) - tr.qc3 * fwat_i;
  R condl_i = (qlwc_i - ql_i) * p.rdt;
  R condi_i = (qiwc_i - qi_i) * p.rdt;

  // overlap (TL :388-397); covpclr only matters to the evaporation branch
  if (tr.clc_o > tr.covptotp) ci.covptot = clc_i;

  // melting (TL :399-427)
  R rfln_i = ci.rfl, sfln_i = ci.sfl;
  if (tr.melt) {
    const R cons_i = tr.cons * (dp_i * tr.rdp - lfdcp_i * tr.rlfdcp);
    const R z2s_i = tr.warm2 ? cons_i * (tr.t0 - p.meltp2) + tr.cons * t_i : zero;
    const R snmlt_i = tr.allm ? ci.sfl : z2s_i;
    rfln_i = ci.rfl + snmlt_i;
    sfln_i = ci.sfl - snmlt_i;
    t_i -= (snmlt_i * tr.cons - tr.snmlt * cons_i) * tr.rcons * tr.rcons;
  }

  // autoconversion (TL :429-503)
  R prr_i = zero, prs_i = zero;
  if (tr.cloudy) {
    const R cldl_i = qlwc_i * tr.rclc - tr.qlwc1 * clc_i * tr.rclc * tr.rclc;
    const R dl_i = R(2) * p.ckl_tl * p.rlcrit * p.rlcrit * tr.ltmp1 * tr.cldl * cldl_i;
    const R qlnew_i =
        clc_i * tr.cldl * tr.ltmp2 + tr.clc_o * cldl_i * tr.ltmp2 - tr.clc_o * tr.cldl * tr.ltmp2 * dl_i;
    prr_i = qlwc_i - qlnew_i;
    qlwc_i = qlnew_i;
    const R cldi_i = qiwc_i * tr.rclc - tr.qiwc1 * clc_i * tr.rclc * tr.rclc;
    const R di_i = p.cki_tl * tr.itmp12 *
                   (tr.itmp11 * (R(2) * tr.cldi * cldi_i * p.ricrit * p.ricrit - R(0.025) * t_i) + R(0.025) * t_i);
    const R qinew_i =
        clc_i * tr.cldi * tr.itmp2 + tr.clc_o * cldi_i * tr.itmp2 - tr.clc_o * tr.cldi * tr.itmp2 * di_i;
    prs_i = qiwc_i - qinew_i;
    qiwc_i = qinew_i;
  }

  // new precipitation (TL :505-523)
  const R dr_i = p.cons2 * (dp_i * (tr.prr + tr.prs) + tr.dp * (prr_i + prs_i));
  R rfreeze_i = zero;
  if (tr.frz1) {
    rfreeze_i = p.cons2 * (dp_i * tr.prr + tr.dp * prr_i);
    sfln_i += dr_i;
  } else {
    rfln_i += dr_i;
  }

  // first-guess T and q (TL :618-659)
  const R dqdt_i = -(condl_i + condi_i) + d.lude * tr.gdp + in.lude * gdp_i;
  const R dlv = tr.lsdcp - tr.lvdcp, dlv_i = lsdcp_i - lvdcp_i;
  const R tmp7 = in.lude * tr.ldcp - dlv * tr.rfreeze1;
  const R dtdt_i = lvdcp_i * tr.condl1 + tr.lvdcp * condl_i + lsdcp_i * tr.condi1 + tr.lsdcp * condi_i -
                   (d.lude * tr.ldcp + in.lude * ldcp_i - dlv_i * tr.rfreeze1 - dlv * rfreeze_i) * tr.gdp -
                   tmp7 * gdp_i;
  t_i += p.dt * dtdt_i;
  q_i += p.dt * dqdt_i;
  const R qold_i = q_i;

  // saturation adjustment (TL :662)
  adj_step_tl(p, tr.rap, d.ap, tr.z3c, tr.z4c, tr.z5c, tr.zalc, tr.sb, t_i, q_i);
  adj_step_tl(p, tr.rap, d.ap, tr.z3c, tr.z4c, tr.z5c, tr.zalc, tr.sa, t_i, q_i);

  // after the adjustment (TL :664-703)
  R dq_i = zero;
  if (tr.pos) {
    dq_i = qold_i - q_i;
    if (p.lregcl) dq_i *= R(0.7);
  }
  const R dr2_i = p.cons2 * (dp_i * tr.dq + tr.dp * dq_i);
  if (tr.frz2) {
    rfreeze_i += fwat_i * tr.dr2 + tr.fwat * dr2_i;
    condi_i += dq_i * p.rdt;
    sfln_i += dr2_i;
  } else {
    condl_i += dq_i * p.rdt;
    rfln_i += dr2_i;
  }

  // outputs (TL :705-753)
  oi.clc = clc_i;
  oi.covptot = zero;
  oi.tnd_q = -(condl_i + condi_i) + d.lude * tr.gdp + in.lude * gdp_i;
  const R tmp8 = in.lude * tr.ldcp - dlv * tr.rfreeze3;
  oi.tnd_t = lvdcp_i * tr.condl2 + tr.lvdcp * condl_i + lsdcp_i * tr.condi2 + tr.lsdcp * condi_i -
             (d.lude * tr.ldcp + in.lude * ldcp_i - dlv_i * tr.rfreeze3 - dlv * rfreeze_i) * tr.gdp - tmp8 * gdp_i;
  oi.tnd_ql = (qlwc_i - ql_i) * p.rdt;
  oi.tnd_qi = (qiwc_i - qi_i) * p.rdt;
  ci.rfl = rfln_i;
  ci.sfl = sfln_i;
}

// ---------------------------------------------------------------------------------------
// level_ad: exact transpose of level_tl about the trajectory `tr` (evaporation branch off).
//   so          : adjoint seeds of the level outputs (clc, tnd_*; covptot is ignored exactly as
//                 the reference drops it when the evaporation branch is off, AD :710-719)
//   a_rfln/a_sfln : in  = adjoint of the fluxes LEAVING the level (already including the seeds
//                   of out_fplsl/out_fplsn at level k+1);
//                   out = adjoint of the fluxes ENTERING the level.
//   a           : adjoints of the level inputs (a.aph0 = -a_dp, a.aph1 = +a_dp, a.lu1 = adjoint
//                 of lu[k+1]); every member is overwritten.
//   ad_ref      : backward first freezing test on the post-adjustment temperature (AD :729) and
//                 the RVTMP2 term on the post-adjustment q (AD :991).
// ---------------------------------------------------------------------------------------
// C::EVAP (LEVAPLS2 or LDRAIN1D): also the reference's adjoint statements of the precipitation-evaporation branch
// (adjoint/_stencils/cloudsc2.py:635-719,808-817,936-941), restated as they are -- they are NOT the transpose of the TL
// statements of that branch (TL :577-579 vs AD :668-674; the pressure-thickness adjoint of dtgdp at :664 carries one
// power of dp; the max-overlap test at :815 cannot fire; a level without evaporation drops the overlap adjoint coming from
// below, :710-719), so the symmetry test does not hold with these flags in the reference either.
//   aph_s   : surface pressure of the column (EVAP only)
//   a_cov   : in = adjoint of the overlap carry LEAVING the level (covptot_i of level k+1), out = entering it
//   a_aph_s : accumulates the adjoint of aph_s over the levels (added to the aph adjoint of the lowest half level)
template <class R, class C = Cfg<false, true>>
CS2_HD void level_ad(const DevParams<R>& p, const LevelIn<R>& in, const Traj<R>& tr, const LevelOut<R>& so,
                     bool ad_ref, R& a_rfln, R& a_sfln, LevelIn<R>& a, R aph_s = R(1), R* a_cov = nullptr,
                     R* a_aph_s = nullptr) {
  const R one = R(1), zero = R(0);
  const R scalm = tr.scalm;
  const R dlv = tr.lsdcp - tr.lvdcp;

  // ---- outputs (AD :503-542)
  R a_qiwc = so.tnd_qi * p.rdt;
  R a_qi0 = -so.tnd_qi * p.rdt;
  R a_qlwc = so.tnd_ql * p.rdt;
  R a_ql0 = -so.tnd_ql * p.rdt;

  R a_gdp = -so.tnd_t * (in.lude * tr.ldcp - dlv * tr.rfreeze3) + so.tnd_q * in.lude;
  R a_condl = so.tnd_t * tr.lvdcp - so.tnd_q;
  R a_condi = so.tnd_t * tr.lsdcp - so.tnd_q;
  R a_lvdcp = so.tnd_t * (tr.condl2 - tr.rfreeze3 * tr.gdp);
  R a_lsdcp = so.tnd_t * (tr.condi2 + tr.rfreeze3 * tr.gdp);
  R a_lude_in = (-so.tnd_t * tr.ldcp + so.tnd_q) * tr.gdp;
  R a_ldcp = -so.tnd_t * tr.gdp * in.lude;
  R a_rfreeze = so.tnd_t * dlv * tr.gdp;
  R a_fwat = zero;
  R a_clc = so.clc;  // adjoint of out_clc
  R a_evapr = zero, a_evaps = zero;
  if (C::EVAP) {  // (AD :513-542)
    a_gdp += -so.tnd_t * (tr.lvdcp * tr.evapr + tr.lsdcp * tr.evaps) + so.tnd_q * (tr.evapr + tr.evaps);
    a_evapr = -so.tnd_t * tr.lvdcp * tr.gdp + so.tnd_q * tr.gdp;
    a_evaps = -so.tnd_t * tr.lsdcp * tr.gdp + so.tnd_q * tr.gdp;
    a_lvdcp -= so.tnd_t * tr.evapr * tr.gdp;
    a_lsdcp -= so.tnd_t * tr.evaps * tr.gdp;
  }

  // ---- after the adjustment (AD :565-592)
  R a_dr2, a_dq;
  if (tr.frz2) {
    a_dr2 = a_sfln + tr.fwat * a_rfreeze;
    a_fwat += tr.dr2 * a_rfreeze;
    a_dq = a_condi * p.rdt;
  } else {
    a_dr2 = a_rfln;
    a_dq = a_condl * p.rdt;
  }
  a_dq += p.cons2 * tr.dp * a_dr2;
  R a_dp = p.cons2 * tr.dq * a_dr2;
  R a_qold = zero, a_q = zero;
  if (tr.pos) {
    if (p.lregcl) a_dq *= R(0.7);
    a_qold = a_dq;
    a_q = -a_dq;
  }

  // ---- saturation adjustment (AD :594-598)
  R a_t = zero, a_ap = zero;
  {
    R a_cond = tr.zalc * a_t - a_q;  // second step
    a_q += tr.sa.rden * a_cond;
    a_t += tr.sa.ct * a_cond;
    a_ap += tr.sa.cap * a_cond;
    a_cond = tr.zalc * a_t - a_q;    // first step
    a_q += tr.sb.rden * a_cond;
    a_t += tr.sb.ct * a_cond;
    a_ap += tr.sb.cap * a_cond;
  }

  // ---- first-guess T and q (AD :600-633)
  a_q += a_qold;
  const R a_dqdt = p.dt * a_q;
  const R a_dtdt = p.dt * a_t;
  R a_q0 = a_q;
  R a_tm = a_t;  // adjoint of the post-melt temperature
  a_gdp += -a_dtdt * (in.lude * tr.ldcp - dlv * tr.rfreeze1) + a_dqdt * in.lude;
  a_condl += a_dtdt * tr.lvdcp - a_dqdt;
  a_condi += a_dtdt * tr.lsdcp - a_dqdt;
  a_lvdcp += a_dtdt * (tr.condl1 - tr.rfreeze1 * tr.gdp);
  a_lsdcp += a_dtdt * (tr.condi1 + tr.rfreeze1 * tr.gdp);
  a_lude_in += (-a_dtdt * tr.ldcp + a_dqdt) * tr.gdp;
  a_ldcp += -a_dtdt * tr.gdp * in.lude;
  a_rfreeze += a_dtdt * dlv * tr.gdp;
  if (C::EVAP) {  // (AD :605-633)
    a_gdp += -a_dtdt * (tr.lvdcp * tr.evapr + tr.lsdcp * tr.evaps) + a_dqdt * (tr.evapr + tr.evaps);
    a_evapr += -a_dtdt * tr.lvdcp * tr.gdp + a_dqdt * tr.gdp;
    a_evaps += -a_dtdt * tr.lsdcp * tr.gdp + a_dqdt * tr.gdp;
    a_lvdcp -= a_dtdt * tr.evapr * tr.gdp;
    a_lsdcp -= a_dtdt * tr.evaps * tr.gdp;
  }

  // ---- precipitation evaporation (AD :635-719)
  R a_corqs = zero, a_covpclr = zero, a_qlim = zero, a_qs_ev = zero, a_cov_lvl = zero;
  if (C::EVAP) {
    R a_prtot = zero;
    if (tr.ev) {
      const R prtot = tr.ev_prtot, dpr = tr.ev_dpr, covpclr = tr.covpclr, covptot1 = tr.covptot1;
      const R beta = tr.ev_beta, b = tr.ev_b, dtgdp = tr.ev_dtgdp, preclr1 = tr.ev_preclr;
      // ice, then warm proportion
      const R e_evaps = a_evaps - a_sfln;
      const R n_sfln = a_sfln + dpr * e_evaps / prtot;
      R a_dpr = tr.ev_sfln * e_evaps / prtot;
      a_prtot = -dpr * tr.ev_sfln * e_evaps / (prtot * prtot);
      const R e_evapr = a_evapr - a_rfln;
      const R n_rfln = a_rfln + dpr * e_evapr / prtot;
      a_dpr += tr.ev_rfln * e_evapr / prtot;
      a_prtot -= dpr * tr.ev_rfln * e_evapr / (prtot * prtot);
      // overlap leaving the level: carry from below + the out_covptot seed
      R a_cv = *a_cov + so.covptot;
      if (tr.ev_gone) {
        a_clc += a_cv;
        a_cv = zero;
      }
      R a_preclr = zero;
      if (tr.ev_capped) {
        a_preclr = a_dpr;
        a_dpr = zero;
      }
      const R a_b = covpclr * a_dpr / dtgdp;
      a_covpclr = b * a_dpr / dtgdp;
      const R a_dtgdp = -covpclr * b * a_dpr / (dtgdp * dtgdp);
      // implicit solution
      const R tmp1 = one + p.dt * beta * tr.corqs;
      const R dqs = in.qsat - tr.ev_qe;
      const R a_beta = p.dt * dqs * a_b / tmp1 - p.dt * p.dt * beta * dqs * tr.corqs * a_b / (tmp1 * tmp1);
      a_qs_ev = p.dt * beta * a_b / tmp1;
      const R a_qe = -p.dt * beta * a_b / tmp1;
      a_corqs = -(p.dt * p.dt) * beta * dqs * beta * a_b / (tmp1 * tmp1);
      // humidity in the moistest covpclr region
      const R sq = sqrt_(in.ap / aph_s);
      const R xx = R(0.5777) * (p.RG * p.RPECONS / R(0.00509)) * pow_(R(0.00509) * covpclr / (preclr1 * sq), R(0.4223));
      a_preclr += xx * sq * a_beta / covpclr;
      const R a_ap_ev = R(0.5) * xx * preclr1 * a_beta / (covpclr * sqrt_(in.ap * aph_s));
      *a_aph_s -= R(0.5) * xx * preclr1 * sq * a_beta / (covpclr * aph_s);
      const R omc = one - tr.clc_o;
      const R romc2 = one / (omc * omc);
      a_covpclr += -(xx * preclr1 * sq * a_beta / (covpclr * covpclr)) - (in.qsat - tr.qlim) * a_qe * romc2 +
                   prtot * a_preclr / covptot1;
      a_qs_ev += a_qe - covpclr * a_qe * romc2;
      a_qlim = covpclr * a_qe * romc2;
      a_clc -= R(2) * (in.qsat - tr.qlim) * covpclr * a_qe / (omc * omc * omc);
      a_prtot += covpclr * a_preclr / covptot1;
      a_cv -= prtot * covpclr * a_preclr / (covptot1 * covptot1);
      a_rfln = n_rfln;
      a_sfln = n_sfln;
      a_cov_lvl = a_cv;
      // pressure thickness through dtgdp (AD :664: one power of dp) and the layer pressure through beta
      a_dp -= p.dt * p.RG * a_dtgdp * tr.rdp;
      a_ap += a_ap_ev;
    }
    a_rfln += a_prtot;
    a_sfln += a_prtot;
  }

  // ---- new precipitation (AD :721-736)
  const bool frz_b = ad_ref ? (tr.tpost < p.RTT) : tr.frz1;
  const R a_dr1 = tr.frz1 ? a_sfln : a_rfln;
  R a_prr = zero;
  if (frz_b) {
    a_dp += a_rfreeze * p.cons2 * tr.prr;
    a_prr = a_rfreeze * p.cons2 * tr.dp;
  }
  a_prr += p.cons2 * tr.dp * a_dr1;
  R a_prs = p.cons2 * tr.dp * a_dr1;
  a_dp += p.cons2 * (tr.prr + tr.prs) * a_dr1;

  // ---- autoconversion (AD :738-782)
  R a_qlwc1, a_qiwc1;
  if (tr.cloudy) {
    const R a_qinew = a_qiwc - a_prs;
    a_qiwc1 = a_prs;
    a_clc += a_qinew * tr.cldi * tr.itmp2;
    R a_cldi = a_qinew * tr.clc_o * tr.itmp2;
    const R a_di = -a_qinew * tr.clc_o * tr.cldi * tr.itmp2;
    a_tm += R(0.025) * p.cki_tl * tr.itmp12 * (one - tr.itmp11) * a_di;
    a_cldi += R(2) * p.cki_tl * tr.itmp12 * tr.itmp11 * tr.cldi * p.ricrit * p.ricrit * a_di;
    a_qiwc1 += a_cldi * tr.rclc;
    a_clc -= tr.qiwc1 * a_cldi * tr.rclc * tr.rclc;

    const R a_qlnew = a_qlwc - a_prr;
    a_qlwc1 = a_prr;
    a_clc += a_qlnew * tr.cldl * tr.ltmp2;
    R a_cldl = a_qlnew * tr.clc_o * tr.ltmp2;
    const R a_dl = -a_qlnew * tr.clc_o * tr.cldl * tr.ltmp2;
    a_cldl += R(2) * p.ckl_tl * tr.ltmp1 * tr.cldl * p.rlcrit * p.rlcrit * a_dl;
    a_qlwc1 += a_cldl * tr.rclc;
    a_clc -= tr.qlwc1 * a_cldl * tr.rclc * tr.rclc;
  } else {
    a_qlwc1 = a_qlwc;
    a_qiwc1 = a_qiwc;
  }

  // ---- melting (AD :784-806)
  R a_t0 = a_tm;
  R a_lfdcp = zero;
  R a_rfl = a_rfln, a_sfl = a_sfln;
  if (tr.melt) {
    const R a_snmlt = a_rfln - a_sfln - a_tm * tr.rcons;
    R a_cons = a_tm * tr.snmlt * tr.rcons * tr.rcons;
    R a_z2s = zero;
    if (tr.allm) {
      a_sfl += a_snmlt;
    } else {
      a_z2s = a_snmlt;
    }
    if (tr.warm2) {
      a_t0 += tr.cons * a_z2s;
      a_cons += (tr.t0 - p.meltp2) * a_z2s;
    }
    a_dp += tr.cons * tr.rdp * a_cons;
    a_lfdcp = -tr.cons * tr.rlfdcp * a_cons;
  }
  a_rfln = a_rfl;
  a_sfln = a_sfl;

  // ---- maximum overlap (AD :808-817)
  if (C::EVAP) {
    if (tr.covpclr1 < zero) a_covpclr = zero;
    R a_cv = a_cov_lvl + a_covpclr;
    a_clc -= a_covpclr;
    if (tr.clc_o > tr.covptot_out) {  // sic (AD :815): covptot_out >= clc_o always, this never fires
      a_clc += a_cv;
      a_cv = zero;
    }
    *a_cov = a_cv;
  }

  // ---- condensation rates, liquid / ice split (AD :819-825)
  a_qlwc1 += a_condl * p.rdt;
  a_ql0 -= a_condl * p.rdt;
  a_qiwc1 += a_condi * p.rdt;
  a_qi0 -= a_condi * p.rdt;
  const R a_qc3 = tr.fwat * a_qlwc1 + (one - tr.fwat) * a_qiwc1;
  a_fwat += tr.qc3 * (a_qlwc1 - a_qiwc1);

  // ---- subsidence (AD :827-855)
  R a_qc2 = zero, a_dqsdz = zero, a_mf = zero, a_rho = zero;
  if (tr.lo3) {
    R a_dqc = -a_qc3;
    if (p.lregcl) a_dqc *= R(0.1);
    a_qc2 = a_qc3;
    const R wd = a_dqc * tr.fac4, wdt = p.dt * wd;
    a_dqsdz = wdt * tr.mfsum;
    a_mf = wdt * tr.dqsdz;
    a_rho = -(wd * tr.dqc);
  }
  const R a_dtdzmo = a_dqsdz * tr.dqsdtemp;
  const R gz = a_dtdzmo * tr.fac3, hz = gz * tr.ldcp;
  R a_dqsdtemp = tr.dtdzmo * (a_dqsdz - hz);
  const R a_rodqsdp = -p.RG * (a_dqsdz + hz);
  a_ldcp -= gz * (p.RG * tr.rodqsdp + tr.dtdzmo * tr.dqsdtemp);
  a_fwat -= a_ldcp * dlv;
  a_lvdcp += tr.fwat * a_ldcp;
  a_lsdcp += (one - tr.fwat) * a_ldcp;
  const R mq = a_rodqsdp * tr.fac2;
  a_rho -= mq * in.qsat;
  const R mr = mq * tr.rho;
  R a_qs = -mr;
  if (C::EVAP) a_qs += a_qs_ev;
  const R rq2 = mr * in.qsat * tr.fac2;
  a_ap += rq2 + a_rho * tr.fac1;
  R a_foeew = -p.RETV * rq2;
  a_t0 -= a_rho * tr.rho * (p.RD * tr.fac1);

  // ---- convective component (AD :857-877)
  R a_lude = zero, a_lu1 = zero, a_qc1 = a_qc2;
  if (tr.lo1) {
    const R wc = (one - tr.clc) * tr.rlu1 * tr.ex * a_clc;
    a_lude = a_qc2 + wc;
    a_lu1 = -(wc * tr.lude * tr.rlu1);
    a_clc *= tr.ex;
  }
  const R dtl = p.dt * a_lude;
  a_lude_in += dtl * tr.gdp;
  a_gdp += dtl * in.lude;
  a_dp -= tr.gdp * tr.rdp * a_gdp;  // gdp = RG / dp

  // ---- cloud fraction and condensate (AD :879-923)
  R a_qt = zero, a_qsat = zero, a_qcrit = zero;
  if (tr.branch == 1) {
    a_qsat = (one - scalm) * a_qc1;
    a_qcrit = -(one - scalm) * a_qc1;
  } else if (tr.branch == 2) {
    R a_qpd = scalm * a_qc1 * (tr.clc * tr.clc);
    R a_qcd = (one - scalm) * a_qc1 * (tr.clc * tr.clc);
    a_clc += R(2) * (scalm * tr.qpd + (one - scalm) * tr.qcd) * tr.clc * a_qc1;
    if (p.lregcl) {
      const R rat = tr.qpd * rcp(tr.qcd);
      const R u = one - scalm * (one - rat);
      a_clc *= min_(R(0.3), R(3.5) * sqrt_(rat * (u * u * u)) * rcp(one - scalm));
    }
    const R h = R(0.5) * rcp(tr.tmp3) * a_clc * tr.rden;
    a_qpd -= h;
    const R a_den = h * tr.qpd * tr.rden;
    a_qcd += a_den;
    a_qt = -scalm * a_den - a_qpd;
    a_qcrit = scalm * a_den - a_qcd;
    a_qsat = a_qcd + a_qpd;
  }
  a_q0 += a_qt;
  a_ql0 += a_qt;
  a_qi0 += a_qt;

  // ---- critical humidity, ice supersaturation (AD :925-932)
  a_qsat += a_qcrit * tr.crh2;
  a_qs += a_qsat * tr.supsat;
  if (tr.ice) a_t0 -= R(0.003) * a_qsat * in.qsat;
  if (C::EVAP) {  // qlim = min(q, qsat) (AD :936-938)
    if (tr.q0 > in.qsat) a_qs += a_qlim; else a_q0 += a_qlim;
  }

  // ---- dqs/dT correction factor (AD :940-967)
  if (C::EVAP) a_dqsdtemp += p.cons3 * a_corqs;
  const R xq = a_dqsdtemp * in.qsat;
  a_qs += tr.fac * tr.cor * a_dqsdtemp;
  const R a_cor = tr.fac * xq;
  const R a_fac = tr.cor * xq;
  R a_esdp = p.RETV * (tr.cor * tr.cor) * a_cor;
  a_fwat += (tr.facw - tr.faci) * a_fac;
  // facw = R5LES rtw^2, faci = R5IES rti^2 (trajectory): their temperature derivatives are -2 facw rtw, -2 faci rti
  a_t0 -= R(2) * a_fac * ((one - tr.fwat) * tr.faci * tr.rti + tr.fwat * tr.facw * tr.rtw);
  if (tr.clip_esdp) a_esdp = zero;
  const R ue = a_esdp * tr.rap;
  a_foeew += ue;
  a_ap -= ue * tr.foeew * tr.rap;
  a_t0 += tr.z3es * (p.RTT - tr.z4es) * (tr.rtm4 * tr.rtm4) * tr.foeew * a_foeew;
  if (tr.cold) a_t0 += R(0.545) * R(0.17) * a_fwat * (tr.tp1 * (R(2) - tr.tp1));

  // ---- latent-heat ratios (AD :988-991)
  if (!p.rvtmp2_zero) {
    const R zz = p.RLVTT * a_lvdcp + p.RLSTT * a_lsdcp + p.RLMLT * a_lfdcp;
    const R den = p.RCPD + p.RCPD * p.RVTMP2 * (ad_ref ? tr.qpost : tr.q0);
    a_q0 += -zz * p.RCPD * p.RVTMP2 / (den * den);
  }

  // ---- level inputs (AD :992-996)
  a.t = a_t0;
  a.tnd_t = p.dt * a_t0;
  a.q = a_q0;
  a.tnd_q = p.dt * a_q0;
  a.supsat = p.dt * a_q0;  // sic: the reference scales the supsat adjoint by dt (AD :992)
  a.ql = a_ql0;
  a.tnd_ql = p.dt * a_ql0;
  a.qi = a_qi0;
  a.tnd_qi = p.dt * a_qi0;
  a.qsat = a_qs;
  a.ap = a_ap;
  a.lude = a_lude_in;
  a.mfu = a_mf;
  a.mfd = a_mf;
  a.lu1 = a_lu1;
  a.aph1 = a_dp;
  a.aph0 = -a_dp;
}

}  // namespace cs2
