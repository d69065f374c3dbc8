// NL level physics, software-pipelined across two levels inside one thread.
//
// A level of CLOUDSC2 (nonlinear/_stencils/cloudsc2.py:113-388) depends on the level above only through the
// precipitation fluxes (rfl, sfl).  Everything that does not read them -- first guess, thermodynamics, cloud fraction,
// detrainment, subsidence, condensation rates (:113-230) and the carry-independent part of the autoconversion
// (:249-259, the liquid branch, and the exponent of the ice branch that does not contain the post-melt temperature) --
// is half "A" of a level; melting, ice autoconversion, new precipitation, the first guess of T and q, the two Newton
// steps of the saturation adjustment (cuadjtqs.py:22-68) and the tendencies (:238-388) are half "B".
//
// With one thread per column and 65 536 columns on 148 SMs there are only 3.5 warps per scheduler, each walking ONE
// dependent FP64 chain (4 + 4 sequential exponentials, ~10 reciprocals per level): the kernel is bound by the
// dependent-issue latency of that chain (stall `wait` > 50 %, FP64 pipe 48 %, DRAM 60 %; profiles/r1g_final.md), not by
// HBM.  `pipe_step` evaluates B of level k together with A of level k+1 -- two independent chains in one basic block:
//   * no branches: every data-dependent `if` of the two halves is a select, exponentials are evaluated
//     unconditionally on a safe argument (a warp of mixed columns takes every branch anyway);
//   * the 10 exponentials of the two halves form 4 groups of mutually independent ones, each evaluated in lockstep
//     (`exp_lockstep`, the branch-free form of exp_batch): {A: tanh, foeew, detrainment | B: ice-fallspeed factor},
//     {B: ice autoconversion}, {A: liquid + ice thresholds | B: first Newton step}, {A: liquid autoconversion | B: second
//     Newton step};
//   * what A hands to the B of the same level one iteration later is `PipeMid` (16 values), kept in registers.
// pipe_b(pipe_a(in), carry) equals level_fwd<LIN = false>(in, carry) up to the association of a few sums (the
// tendencies are assembled from A's partial sums, as in cs2_physics_split.cuh).  Default flags only (evaporation branch
// off, LPHYLIN on, RVTMP2 == 0): the other configurations keep the one-level-at-a-time sweep.
#pragma once

#include "../cs2_physics.cuh"

namespace cs2 {

// The exponentials of the two halves are evaluated in lockstep groups without a range check (branch-free).  Arguments that
// are unbounded below -- the detrainment and the two autoconversion thresholds, all <= 0 and only used as 1 - exp(.) -- are
// clamped at -708 by pipe_step (3.3e-308 instead of a denormal or 0); the exponents of Tetens' formula, of the fallspeed
// factor and of the autoconversion rates are O(10) for any temperature a column can have.
template <int N, class R>
CS2_HD void exp_lockstep(const R (&x)[N], R (&e)[N]) {
  exp_batch<N, false>(x, e);
}

// What half A of a level hands to half B of the same level.
template <class R>
struct PipeMid {
  R t0;              // first-guess temperature
  R cons, rcons;     // melting: cons2 dp / lfdcp and its reciprocal (:239,243)
  R cdp;             // cons2 dp
  R qa, dqdt;        // q0 + dt dqdt, and dqdt without the post-adjustment term (:328,343)
  R dta;             // dtdt without the freezing term (:329-339)
  R dlvgdp;          // (lsdcp - lvdcp) gdp
  R rap, fwat;
  R qiwc1, qic;      // ice content before autoconversion; clc * cldi (= qiwc1 up to rounding) where cloudy
  R dic;             // ckcodti (1 - exp(-(cldi / icrit)^2))                       (:268, without the fallspeed factor)
  R prr;             // rain production (:258), final
  R qi0;
  bool cloudy;
};

// Level outputs that are final after half A (stored one iteration before the rest of the level).
template <class R>
struct PipeOutA {
  R clc, tnd_ql;
};

template <class R>
struct PipeOutB {
  R tnd_q, tnd_t, tnd_qi;
};

// B of the level described by `m` (carry `c` in/out, outputs `ob`) together with A of the level whose inputs are
// `in` (-> `mn`, `oa`).  ad_ref: the literal AD stencil's second freezing test (see level_fwd).
template <class R>
CS2_HD void pipe_step(const DevParams<R>& p, const PipeMid<R>& m, Carry<R>& c, bool ad_ref, PipeOutB<R>& ob,
                      const LevelIn<R>& in, R scalm, R crh2, bool conv_ok, PipeMid<R>& mn, PipeOutA<R>& oa) {
  const R one = R(1), zero = R(0);

  // ---- A1: first guess, reciprocals, exponents of the first group (:104,115-117,130-134,141-151,210-213)
  const R t0 = in.t + p.dt * in.tnd_t;
  const R q0 = in.q + p.dt * in.tnd_q + in.supsat;
  const R ql0 = in.ql + p.dt * in.tnd_ql;
  const R qi0 = in.qi + p.dt * in.tnd_qi;
  const R dp = in.aph1 - in.aph0;
  const R rdp = rcp(dp), rap = rcp(in.ap);
  const R lfdcp = p.lfdcp0, lsdcp = p.lsdcp0, lvdcp = p.lvdcp0, rlfdcp = p.rlfdcp0;  // RVTMP2 == 0 (launcher's condition)
  const R rtw = rcp(t0 - p.R4LES), rti = rcp(t0 - p.R4IES);
  const bool cold = t0 < p.RTT;
  const R z3es = cold ? p.R3IES : p.R3LES;
  const R rtm4 = cold ? rti : rtw;
  const R gdp = p.RG * rdp;
  const R lude = p.dt * in.lude * gdp;
  const bool lo1 = conv_ok && (lude >= p.RLMIN) && (in.lu1 >= p.ZEPS2);
  const R rlu1 = rcp(lo1 ? in.lu1 : one);

  // ---- B1: melting of the incoming snow (:238-246), branch-free: with sfl == 0 every term below vanishes
  //      (snmlt = 0, so rfl + 0, 0 - 0 and t0 - 0 * rcons are the untouched values)
  const R z2s = (m.t0 > p.meltp2) ? m.cons * (m.t0 - p.meltp2) : zero;
  const R snmlt = (c.sfl <= z2s) ? c.sfl : z2s;
  R rfln = c.rfl + snmlt;
  R sfln = c.sfl - snmlt;
  const R tmelt = m.t0 - snmlt * m.rcons;

  // ---- group 1: A tanh / foeew / detrainment, B ice-fallspeed factor
  R e1[4];
  {
    const R a1[4] = {R(0.34) * (t0 - p.RLPTRC), z3es * (t0 - p.RTT) * rtm4, lo1 ? max_(-lude * rlu1, R(-708)) : zero,
                     R(0.025) * (tmelt - p.RTT)};
    exp_lockstep<4>(a1, e1);
  }

  // ---- B2: ice autoconversion exponent (:262-271)
  const R itmp12 = e1[3];
  R e2[1];
  {
    const R a2[1] = {-(m.dic * itmp12)};
    exp_lockstep<1>(a2, e2);
  }

  // ---- A2: thermodynamics, cloud fraction, detrainment, subsidence, condensation rates (:141-230)
  const R tp1 = R(2) * e1[0] * rcp(one + e1[0]);  // 1 + tanh(0.17 (t0 - RLPTRC))
  const R fwat = cold ? R(0.545) * tp1 : one;
  const R foeew = p.R2ES * e1[1];
  const bool clip_esdp = foeew * rap > p.ZQMAX;
  const R fac = fwat * (p.R5LES * rtw * rtw) + (one - fwat) * (p.R5IES * rti * rti);
  const R fac2 = rcp(in.ap - p.RETV * foeew);
  const R cor = clip_esdp ? p.cor_clip : in.ap * fac2;
  const R dqsdtemp = fac * cor * in.qsat;
  const R supsat = (t0 < p.RTICE) ? (R(1.8) - R(0.003) * t0) : one;
  const R qsat = in.qsat * supsat;
  const R qcrit = crh2 * qsat;
  const R qt = q0 + ql0 + qi0;
  const bool br0 = qt < qcrit, br1 = !br0 && (qt >= qsat), br2 = !br0 && !br1;
  const R qpd = qsat - qt, qcd = qsat - qcrit;
  const R rden = rcp(br2 ? qcd - scalm * (qt - qcrit) : one);
  const R tmp3 = sqrt_(br2 ? qpd * rden : one);
  const R clc2 = one - tmp3;
  const R clc = br0 ? zero : (br1 ? one : clc2);
  const R qc1 = br0 ? zero : (br1 ? (one - scalm) * qcd : (scalm * qpd + (one - scalm) * qcd) * (clc2 * clc2));
  const R clc_o = clc + (one - clc) * (one - e1[2]);  // e1[2] == 1 exactly without detrainment
  const R qc2 = lo1 ? qc1 + lude : qc1;
  const R rho = in.ap * rcp(p.RD * t0);
  const R rodqsdp = -rho * in.qsat * fac2;
  const R ldcp = fwat * lvdcp + (one - fwat) * lsdcp;
  const R dtdzmo = p.RG * (p.rcpd - ldcp * rodqsdp) * rcp(one + ldcp * dqsdtemp);
  const R dqsdz = dqsdtemp * dtdzmo - p.RG * rodqsdp;
  const R sub = p.dt * dqsdz * (in.mfu + in.mfd) * (p.RD * t0 * rap);
  const R qc3 = (sub < qc2) ? (qc2 - sub) : zero;
  const R qlwc1 = qc3 * fwat, qiwc1 = qc3 * (one - fwat);
  const R condl1 = (qlwc1 - ql0) * p.rdt, condi1 = (qiwc1 - qi0) * p.rdt;
  const bool cloudy = clc_o > p.ZEPS2;
  const R rclc = rcp(cloudy ? clc_o : one);
  const R cldl = qlwc1 * rclc, cldi = qiwc1 * rclc;
  const R xl = cldl * p.rlcrit, xi = cldi * p.ricrit;

  // ---- B3: new precipitation and its phase, first guess of T and q, first Newton step up to its exponent
  //      (:267-285,328-347; cuadjtqs.py:22-26,54-63)
  const R qiwc = m.cloudy ? m.qic * e2[0] : m.qiwc1;
  const R prs = m.qiwc1 - qiwc;
  const R dr1 = m.cdp * (m.prr + prs);
  const bool frz1 = tmelt < p.RTT;
  const R rfreeze1 = frz1 ? m.cdp * m.prr : zero;
  sfln = frz1 ? sfln + dr1 : sfln;
  rfln = frz1 ? rfln : rfln + dr1;
  const R t3 = tmelt + p.dt * (m.dta + m.dlvgdp * rfreeze1);
  const bool warmc = t3 > p.RTT;
  const R z3c = warmc ? p.R3LES : p.R3IES, z4c = warmc ? p.R4LES : p.R4IES;
  const R z5c = warmc ? p.R5ALVCP : p.R5ALSCP, zalc = warmc ? p.RALVDCP : p.RALSDCP;
  const R rt_b = rcp(t3 - z4c);

  // ---- group 3: A autoconversion thresholds (:255,268), B first Newton step
  R e3[3];
  {
    const R a3[3] = {max_(-(xl * xl), R(-708)), max_(-(xi * xi), R(-708)), z3c * (t3 - p.RTT) * rt_b};
    exp_lockstep<3>(a3, e3);
  }

  // ---- B4: rest of the first Newton step (merged form, see adj_step<LIN = false>), second one up to its exponent
  R t = t3, q = m.qa;
  {
    const R qs1 = p.R2ES * e3[2] * m.rap;
    const R qsc = (qs1 > p.ZQMAX) ? p.ZQMAX : qs1;
    const R z2 = z5c * rt_b * rt_b;
    const R a = one - p.RETV * qsc;
    const R cond = a * (q * a - qsc) * rcp(a * a + qsc * z2);
    t = t + zalc * cond;
    q = q - cond;
  }
  const R rt_a = rcp(t - z4c);

  // ---- group 4: A liquid autoconversion (:255-257), B second Newton step
  R e4[2];
  {
    const R a4[2] = {-(p.ckcodtl * (one - e3[0])), z3c * (t - p.RTT) * rt_a};
    exp_lockstep<2>(a4, e4);
  }

  // ---- A3: rain production, what the level's half B needs
  const R qlwc = cloudy ? clc_o * cldl * e4[0] : qlwc1;
  const R dqdt = -(condl1 + condi1) + in.lude * gdp;
  mn.t0 = t0;
  mn.cons = p.cons2 * dp * rlfdcp;
  mn.rcons = lfdcp * p.rgdt * rdp;
  mn.cdp = p.cons2 * dp;
  mn.qa = q0 + p.dt * dqdt;
  mn.dqdt = dqdt;
  mn.dta = lvdcp * condl1 + lsdcp * condi1 - in.lude * ldcp * gdp;
  mn.dlvgdp = (lsdcp - lvdcp) * gdp;
  mn.rap = rap;
  mn.fwat = fwat;
  mn.qiwc1 = qiwc1;
  mn.qic = clc_o * cldi;
  mn.dic = p.ckcodti * (one - e3[1]);
  mn.prr = cloudy ? qlwc1 - qlwc : zero;
  mn.qi0 = qi0;
  mn.cloudy = cloudy;
  oa.clc = clc_o;
  oa.tnd_ql = (qlwc - ql0) * p.rdt;

  // ---- B5: rest of the second Newton step, rain fraction and freezing after the adjustment, tendencies (:350-388)
  const R tpre = t;
  {
    const R qs1 = p.R2ES * e4[1] * m.rap;
    const R qsc = (qs1 > p.ZQMAX) ? p.ZQMAX : qs1;
    const R z2 = z5c * rt_a * rt_a;
    const R a = one - p.RETV * qsc;
    const R cond = a * (q * a - qsc) * rcp(a * a + qsc * z2);
    t = tpre + zalc * cond;
    q = q - cond;
  }
  const R dq = (m.qa >= q) ? (m.qa - q) : zero;
  const R dr2 = m.cdp * dq;
  const bool frz2 = (ad_ref ? t3 : t) < p.RTT;
  const R rfreeze = frz2 ? rfreeze1 + m.fwat * dr2 : rfreeze1;
  sfln = frz2 ? sfln + dr2 : sfln;
  rfln = frz2 ? rfln : rfln + dr2;
  const R lat = frz2 ? p.lsdcp0 : p.lvdcp0;
  ob.tnd_q = m.dqdt - dq * p.rdt;
  ob.tnd_t = m.dta + lat * (dq * p.rdt) + m.dlvgdp * rfreeze;
  ob.tnd_qi = (qiwc - m.qi0) * p.rdt;
  c.rfl = rfln;
  c.sfl = sfln;
}

// A `PipeMid` whose half B is a no-op on a zero carry: what the prologue iteration of the sweep (A of level 0 with
// nothing to finish) feeds to pipe_step.
template <class R>
CS2_HD PipeMid<R> pipe_mid_idle(const DevParams<R>& p) {
  PipeMid<R> m;
  m.t0 = p.RTT + R(10);
  m.cons = m.rcons = m.cdp = m.qa = m.dqdt = m.dta = m.dlvgdp = R(0);
  m.rap = R(1e-5);
  m.fwat = R(1);
  m.qiwc1 = m.qic = m.dic = m.prr = m.qi0 = R(0);
  m.cloudy = false;
  return m;
}

}  // namespace cs2
