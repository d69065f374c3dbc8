// NL level physics split at the point where the vertical dependency enters.
//
// A level of CLOUDSC2 (nonlinear/_stencils/cloudsc2.py:113-388) depends on the level above only through the
// precipitation fluxes (rfl, sfl) -- and, with LEVAPLS2 / LDRAIN1D, the overlap covptot.  Everything up to and
// including the condensation rates (:113-230) is a function of the level's own inputs:
//
//   level_nl_a  (:113-230)  first guess, thermodynamics, cloud fraction, convective + subsidence terms,
//                           condensation rates            -> 15 (17) handed-over values (`Mid`)
//   level_nl_b  (:238-388)  melting of the incoming snow, autoconversion, new precipitation, saturation
//                           adjustment (cuadjtqs.py:22-68), tendencies and outgoing fluxes
//
// level_nl_b(level_nl_a(in), carry) == level_fwd<LIN = false>(in, carry) up to the association of a few sums (the
// tendencies are assembled from A's partial sums).  The split kernel (cs2_split_columns.cuh) runs A and B of a
// column in two different warps that hand `Mid` over through shared memory, so the dependent FP64 chain a warp
// walks per level is half as long.  Evaporation branch off only (the default configuration).
#pragma once

#include "../cs2_physics.cuh"

namespace cs2 {

enum { M_T0, M_CONS, M_RCONS, M_CDP, M_CLC, M_QLWC1, M_QIWC1, M_QA, M_DQDT, M_DTA, M_DLVGDP, M_RAP, M_FWAT, M_QL0, M_QI0,
       M_NZ,               // number of handed-over values when RVTMP2 == 0
       M_LVDCP = M_NZ, M_LSDCP, M_N };

template <class R>
struct Mid {
  R v[M_N];
  CS2_HD R operator[](int n) const { return v[n]; }
};

// C::TETENS as in level_fwd; C::EVAP must be false
template <class R, class C>
CS2_HD void level_nl_a(const DevParams<R>& p, const LevelIn<R>& in, R scalm, R crh2, bool conv_ok, Mid<R>& m) {
  const R one = R(1), zero = R(0);
  // first guess (:104,115-117)
  const R t0 = in.t + p.dt * in.tnd_t;
  const R q0 = in.q + p.dt * in.tnd_q + in.supsat;
  const R ql0 = in.ql + p.dt * in.tnd_ql;
  const R qi0 = in.qi + p.dt * in.tnd_qi;

  // thermodynamic constants (:130-134)
  const R dp = in.aph1 - in.aph0;
  const R rdp = rcp(dp), rap = rcp(in.ap);
  R lfdcp, lsdcp, lvdcp, rlfdcp;
  if (p.rvtmp2_zero) {
    lfdcp = p.lfdcp0; lsdcp = p.lsdcp0; lvdcp = p.lvdcp0; rlfdcp = p.rlfdcp0;
  } else {
    const R zzinv = rcp(p.RCPD + p.RCPD * p.RVTMP2 * q0);
    lfdcp = p.RLMLT * zzinv; lsdcp = p.RLSTT * zzinv; lvdcp = p.RLVTT * zzinv;
    rlfdcp = rcp(lfdcp);
  }

  // dqs/dT correction factor (:141-160)
  const R rtw = rcp(t0 - p.R4LES), rti = rcp(t0 - p.R4IES);
  R fwat, foeew;
  bool clip_esdp = false;
  if (C::TETENS) {
    R z3es, rtm4;
    if (t0 < p.RTT) {
      fwat = R(0.545) * one_plus_tanh<R>(R(0.17) * (t0 - p.RLPTRC));
      z3es = p.R3IES; rtm4 = rti;
    } else {
      fwat = one;
      z3es = p.R3LES; rtm4 = rtw;
    }
    foeew = p.R2ES * exp_(z3es * (t0 - p.RTT) * rtm4);
    clip_esdp = foeew * rap > p.ZQMAX;
  } else {
    fwat = foealfa(p, t0);
    foeew = foeew_mixed(p, t0, fwat);
  }
  const R fac = fwat * (p.R5LES * rtw * rtw) + (one - fwat) * (p.R5IES * rti * rti);
  const R fac2 = rcp(in.ap - p.RETV * foeew);
  const R cor = clip_esdp ? p.cor_clip : in.ap * fac2;
  const R dqsdtemp = fac * cor * in.qsat;

  // ice supersaturation, critical humidity (:188-193)
  const R supsat = (t0 < p.RTICE) ? (R(1.8) - R(0.003) * t0) : one;
  const R qsat = in.qsat * supsat;
  const R qcrit = crh2 * qsat;

  // uniform total-water distribution (:196-207)
  const R qt = q0 + ql0 + qi0;
  R clc, qc;
  if (qt < qcrit) {
    clc = zero;
    qc = zero;
  } else if (qt >= qsat) {
    clc = one;
    qc = (one - scalm) * (qsat - qcrit);
  } else {
    const R qpd = qsat - qt, qcd = qsat - qcrit;
    clc = one - sqrt_(qpd * rcp(qcd - scalm * (qt - qcrit)));
    qc = (scalm * qpd + (one - scalm) * qcd) * (clc * clc);
  }

  // convective component (:210-215)
  const R gdp = p.RG * rdp;
  const R lude = p.dt * in.lude * gdp;
  if (conv_ok && (lude >= p.RLMIN) && (in.lu1 >= p.ZEPS2)) {
    clc = clc + (one - clc) * (one - exp_(-lude * rcp(in.lu1)));
    qc += lude;
  }

  // compensating subsidence (:218-224)
  const R rho = in.ap * rcp(p.RD * t0);
  const R rodqsdp = -rho * in.qsat * fac2;
  const R ldcp = fwat * lvdcp + (one - fwat) * lsdcp;
  const R dtdzmo = p.RG * (p.rcpd - ldcp * rodqsdp) * rcp(one + ldcp * dqsdtemp);
  const R dqsdz = dqsdtemp * dtdzmo - p.RG * rodqsdp;
  const R sub = p.dt * dqsdz * (in.mfu + in.mfd) * (p.RD * t0 * rap);
  qc = (sub < qc) ? (qc - sub) : zero;

  // new liquid / ice and condensation rates (:227-230)
  const R qlwc1 = qc * fwat, qiwc1 = qc * (one - fwat);
  const R condl1 = (qlwc1 - ql0) * p.rdt, condi1 = (qiwc1 - qi0) * p.rdt;

  // partial sums of the first guess and of the tendencies (:328-344,367-388): everything but the freezing and
  // post-adjustment terms, which level_nl_b adds
  const R dqdt = -(condl1 + condi1) + in.lude * gdp;
  m.v[M_T0] = t0;
  m.v[M_CONS] = p.cons2 * dp * rlfdcp;
  m.v[M_RCONS] = lfdcp * p.rgdt * rdp;
  m.v[M_CDP] = p.cons2 * dp;
  m.v[M_CLC] = clc;
  m.v[M_QLWC1] = qlwc1;
  m.v[M_QIWC1] = qiwc1;
  m.v[M_QA] = q0 + p.dt * dqdt;
  m.v[M_DQDT] = dqdt;
  m.v[M_DTA] = lvdcp * condl1 + lsdcp * condi1 - in.lude * ldcp * gdp;
  m.v[M_DLVGDP] = (lsdcp - lvdcp) * gdp;
  m.v[M_RAP] = rap;
  m.v[M_FWAT] = fwat;
  m.v[M_QL0] = ql0;
  m.v[M_QI0] = qi0;
  m.v[M_LVDCP] = lvdcp;
  m.v[M_LSDCP] = lsdcp;
}

// M: anything with `R operator[](int) const` (a Mid<R>, or a view of the shared-memory hand-over slots)
template <class R, class M>
CS2_HD void level_nl_b(const DevParams<R>& p, const M& m, Carry<R>& c, LevelOut<R>& o) {
  const R one = R(1), zero = R(0);
  const R t0 = m[M_T0];
  // melting of incoming snow (:238-246)
  R rfln = c.rfl, sfln = c.sfl, tmelt = t0;
  if (c.sfl != zero) {
    const R z2s = (t0 > p.meltp2) ? m[M_CONS] * (t0 - p.meltp2) : zero;
    const R snmlt = (c.sfl <= z2s) ? c.sfl : z2s;
    rfln = c.rfl + snmlt;
    sfln = c.sfl - snmlt;
    tmelt = t0 - snmlt * m[M_RCONS];
  }

  // autoconversion of cloud liquid and ice (:249-272)
  const R clc = m[M_CLC];
  const R qlwc1 = m[M_QLWC1], qiwc1 = m[M_QIWC1];
  R qlwc = qlwc1, qiwc = qiwc1, prr = zero, prs = zero;
  if (clc > p.ZEPS2) {
    const R rclc = rcp(clc);
    const R cldl = qlwc1 * rclc;
    const R xl = cldl * p.rlcrit;
    const R ltmp1 = exp_(-(xl * xl));
    qlwc = clc * cldl * exp_(-(p.ckcodtl * (one - ltmp1)));
    prr = qlwc1 - qlwc;
    const R cldi = qiwc1 * rclc;
    const R xi = cldi * p.ricrit;
    const R itmp11 = exp_(-(xi * xi));
    const R itmp12 = exp_(R(0.025) * (tmelt - p.RTT));
    qiwc = clc * cldi * exp_(-(p.ckcodti * itmp12 * (one - itmp11)));
    prs = qiwc1 - qiwc;
  }

  // new precipitation and its phase (:275-285)
  const R cdp = m[M_CDP];
  const R dr1 = cdp * (prr + prs);
  R rfreeze = zero;
  if (tmelt < p.RTT) {
    rfreeze = cdp * prr;
    sfln += dr1;
  } else {
    rfln += dr1;
  }

  // first-guess T and q (:328-344)
  const R dta = m[M_DTA], dlvgdp = m[M_DLVGDP];
  const R qa = m[M_QA];
  R t = tmelt + p.dt * (dta + dlvgdp * rfreeze), q = qa;

  // saturation adjustment, two Newton steps (:347; cuadjtqs.py:38-68)
  const bool warmc = t > p.RTT;
  const R z3c = warmc ? p.R3LES : p.R3IES, z4c = warmc ? p.R4LES : p.R4IES;
  const R z5c = warmc ? p.R5ALVCP : p.R5ALSCP, zalc = warmc ? p.RALVDCP : p.RALSDCP;
  const R rap = m[M_RAP];
  AdjStep<R> s;
  Trans<R, 0> x;
  adj_step<R, false>(p, rap, z3c, z4c, z5c, zalc, t, q, s, x, CK_SB);
  adj_step<R, false>(p, rap, z3c, z4c, z5c, zalc, t, q, s, x, CK_SA);

  // rain fraction and freezing after the adjustment (:350-364), tendencies (:367-388)
  const R dq = (qa >= q) ? (qa - q) : zero;
  const R dr2 = cdp * dq;
  R lat;
  if (t < p.RTT) {
    rfreeze += m[M_FWAT] * dr2;
    sfln += dr2;
    lat = p.rvtmp2_zero ? p.lsdcp0 : m[M_LSDCP];
  } else {
    rfln += dr2;
    lat = p.rvtmp2_zero ? p.lvdcp0 : m[M_LVDCP];
  }
  o.clc = clc;
  o.covptot = zero;
  o.tnd_q = m[M_DQDT] - dq * p.rdt;
  o.tnd_t = dta + lat * (dq * p.rdt) + dlvgdp * rfreeze;
  o.tnd_ql = (qlwc - m[M_QL0]) * p.rdt;
  o.tnd_qi = (qiwc - m[M_QI0]) * p.rdt;
  c.rfl = rfln;
  c.sfl = sfln;
}

}  // namespace cs2
