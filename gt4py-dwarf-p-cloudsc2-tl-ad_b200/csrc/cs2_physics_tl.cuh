// level_fwd_tl: trajectory and tangent of one level evaluated TOGETHER, statement by statement
// (reference tangent_linear/_stencils/cloudsc2.py:149-753 has the same interleaved structure).
//
// Same mathematics as level_fwd<LIN = true> followed by level_tl (cs2_physics.cuh), which remain the
// specification and the code the AD sweep transposes.  Evaluating each tangent right after its forward
// statement (a) lets most trajectory intermediates die immediately instead of staying live across the whole
// level, and (b) puts the forward chain and the lagging tangent chain into the same basic blocks, so the
// scheduler has two semi-independent FP64 dependency chains to interleave -- the TL kernel is bound by
// fixed-latency dependency stalls with only two warps per scheduler (profiles/r1e_kernels.md).
#pragma once

#include "cs2_physics.cuh"

namespace cs2 {

// one Newton step of the saturation adjustment with its tangent (cuadjtqs.py TL :22-55).  The trajectory statements are those of
// adj_step<LIN> (cs2_physics.cuh), so the TL and AD trajectories stay bit-identical; the tangent applies the step's local
// Jacobian (adj_step_coef: cond_i = rden q_i + ct t_i + cap ap_i), the very coefficients whose transpose the AD sweep applies.
template <class R>
CS2_HD void adj_step_fwd_tl(const DevParams<R>& p, R rap, R ap_i, R z3, R z4, R z5, R zal, R& t, R& q, R& t_i, R& q_i) {
  AdjStep<R> s;
  s.rt = rcp(t - z4);
  s.foeew = p.R2ES * exp_(z3 * (t - p.RTT) * s.rt);
  const R qs1 = s.foeew * rap;
  s.clipped = qs1 > p.ZQMAX;
  s.qsc = s.clipped ? p.ZQMAX : qs1;
  s.z2s = z5 * s.rt * s.rt;
  const R a = R(1) - p.RETV * s.qsc;
  s.cor = rcp(a);
  const R a2 = a * a;
  s.rden = a2 * rcp(a2 + s.qsc * s.z2s);
  s.qs = s.qsc * s.cor;
  s.cond = (q - s.qs) * s.rden;
  R ct, cap;
  adj_step_coef(p, rap, z3 * (p.RTT - z4), s, ct, cap);
  const R cond = s.cond;
  const R cond_i = s.rden * q_i + ct * t_i + cap * ap_i;
  t += zal * cond;
  t_i += zal * cond_i;
  q -= cond;
  q_i -= cond_i;
}

// C::EVAP (LEVAPLS2 or LDRAIN1D): also the precipitation-evaporation branch and its tangent, a statement-by-statement
// restatement of tangent_linear/_stencils/cloudsc2.py:224-230,388-397,525-616 -- including that stencil's `dt**2` in the
// tangent of the implicit solution (:577-579), where the quotient rule gives `dt`: the reference's TL of this branch is
// not the derivative of its NL, and the drop-in reproduces the reference.  aph_s / aph_s_i: surface
// pressure of the column and its perturbation (only read when C::EVAP).
template <class R, class C = Cfg<false, true>>
CS2_HD void level_fwd_tl(const DevParams<R>& p, const LevelIn<R>& in, const LevelIn<R>& d, R scalm, R crh2, bool conv_ok,
                         R aph_s, R aph_s_i, Carry<R>& c, Carry<R>& ci, LevelOut<R>& o, LevelOut<R>& oi) {
  const R one = R(1), zero = R(0);
  // ---- first guess (TL :149-156)
  const R t0 = in.t + p.dt * in.tnd_t;
  R t_i = d.t + p.dt * d.tnd_t;
  const R q0 = in.q + p.dt * in.tnd_q + in.supsat;
  R q_i = d.q + p.dt * d.tnd_q + d.supsat;
  const R ql0 = in.ql + p.dt * in.tnd_ql, ql_i = d.ql + p.dt * d.tnd_ql;
  const R qi0 = in.qi + p.dt * in.tnd_qi, qi_i = d.qi + p.dt * d.tnd_qi;

  // ---- thermodynamic constants (TL :170-180)
  const R dp = in.aph1 - in.aph0, dp_i = d.aph1 - d.aph0;
  const R rdp = rcp(dp), rap = rcp(in.ap);
  R lfdcp, lsdcp, lvdcp, rlfdcp, lfdcp_i = zero, lsdcp_i = zero, lvdcp_i = zero;
  if (p.rvtmp2_zero) {
    lfdcp = p.lfdcp0; lsdcp = p.lsdcp0; lvdcp = p.lvdcp0; rlfdcp = p.rlfdcp0;
  } else {
    const R zzinv = rcp(p.RCPD + p.RCPD * p.RVTMP2 * q0);
    const R zz_i = -p.RCPD * p.RVTMP2 * q_i * zzinv * zzinv;
    lfdcp = p.RLMLT * zzinv; lsdcp = p.RLSTT * zzinv; lvdcp = p.RLVTT * zzinv;
    lfdcp_i = p.RLMLT * zz_i; lsdcp_i = p.RLSTT * zz_i; lvdcp_i = p.RLVTT * zz_i;
    rlfdcp = rcp(lfdcp);
  }

  // ---- dqs/dT correction factor (TL :188-222)
  const R rtw = rcp(t0 - p.R4LES), rti = rcp(t0 - p.R4IES);
  const bool cold = t0 < p.RTT;
  R fwat, fwat_i, z3es, z4es, rtm4;
  if (cold) {
    const R tp1 = one_plus_tanh<R>(R(0.17) * (t0 - p.RLPTRC));
    fwat = R(0.545) * tp1;
    fwat_i = R(0.545) * R(0.17) * t_i * (tp1 * (R(2) - tp1));
    z3es = p.R3IES; z4es = p.R4IES; rtm4 = rti;
  } else {
    fwat = one; fwat_i = zero;
    z3es = p.R3LES; z4es = p.R4LES; rtm4 = rtw;
  }
  const R foeew = p.R2ES * exp_(z3es * (t0 - p.RTT) * rtm4);
  const R foeew_i = z3es * (p.RTT - z4es) * t_i * foeew * rtm4 * rtm4;
  const R esdp1 = foeew * rap;
  const bool clip_esdp = esdp1 > p.ZQMAX;
  const R esdp_i = clip_esdp ? zero : (foeew_i * rap - foeew * d.ap * rap * rap);
  const R facw = p.R5LES * rtw * rtw, faci = p.R5IES * rti * rti;
  const R facw_i = R(-2) * facw * t_i * rtw, faci_i = R(-2) * faci * t_i * rti;
  const R fac = fwat * facw + (one - fwat) * faci;
  const R fac_i = fwat_i * (facw - faci) + fwat * facw_i + (one - fwat) * faci_i;
  // 1 / (1 - RETV esdp) = ap / (ap - RETV foeew) where the clip does not bind: the reciprocal the subsidence term needs anyway
  const R fac2 = rcp(in.ap - p.RETV * foeew);
  const R cor = clip_esdp ? p.cor_clip : in.ap * fac2;
  const R cor_i = p.RETV * esdp_i * cor * cor;
  const R dqsdtemp = fac * cor * in.qsat;
  const R dqsdtemp_i = fac_i * cor * in.qsat + fac * cor_i * in.qsat + fac * cor * d.qsat;
  R corqs = one, corqs_i = zero, qlim = zero, qlim_i = zero;
  if (C::EVAP) {  // (TL :216-230)
    corqs = one + p.cons3 * dqsdtemp;
    corqs_i = p.cons3 * dqsdtemp_i;
    const bool qclip = q0 > in.qsat;
    qlim = qclip ? in.qsat : q0;
    qlim_i = qclip ? d.qsat : q_i;
  }

  // ---- critical humidity (TL :255-265)
  const bool ice = t0 < p.RTICE;
  const R supsat = ice ? (R(1.8) - R(0.003) * t0) : one;
  const R supsat_i = ice ? R(-0.003) * t_i : zero;
  const R qsat = in.qsat * supsat;
  const R qsat_i = d.qsat * supsat + in.qsat * supsat_i;
  const R qcrit = crh2 * qsat, qcrit_i = crh2 * qsat_i;

  // ---- cloud fraction and condensate (TL :267-306)
  const R qt = q0 + ql0 + qi0, qt_i = q_i + ql_i + qi_i;
  R clc, clc_i = zero, qc, qc_i = zero;
  if (qt < qcrit) {
    clc = zero;
    qc = zero;
  } else if (qt >= qsat) {
    clc = one;
    qc = (one - scalm) * (qsat - qcrit);
    qc_i = (one - scalm) * (qsat_i - qcrit_i);
  } else {
    const R qpd = qsat - qt, qpd_i = qsat_i - qt_i;
    const R qcd = qsat - qcrit, qcd_i = qsat_i - qcrit_i;
    const R den = qcd - scalm * (qt - qcrit);
    const R rden = rcp(den);
    const R tmp3 = sqrt_(qpd * rden);
    clc = one - tmp3;
    clc_i = R(-0.5) * rcp(tmp3) * (qpd_i * den - qpd * (qcd_i - scalm * (qt_i - qcrit_i))) * rden * rden;
    if (p.lregcl) {
      const R rat = qpd * rcp(qcd);
      const R u = one - scalm * (one - rat);
      clc_i *= min_(R(0.3), R(3.5) * sqrt_(rat * (u * u * u)) * rcp(one - scalm));
    }
    const R wq = scalm * qpd + (one - scalm) * qcd;
    qc = wq * (clc * clc);
    qc_i = (scalm * qpd_i + (one - scalm) * qcd_i) * (clc * clc) + R(2) * wq * clc * clc_i;
  }

  // ---- convective component (TL :308-325)
  const R gdp = p.RG * rdp;
  const R gdp_i = -gdp * dp_i * rdp;
  const R lude = p.dt * in.lude * gdp;
  const R lude_i = p.dt * (d.lude * gdp + in.lude * gdp_i);
  if (conv_ok && (lude >= p.RLMIN) && (in.lu1 >= p.ZEPS2)) {
    const R rlu1 = rcp(in.lu1);
    const R ex = exp_(-lude * rlu1);
    clc_i += -clc_i * (one - ex) + (one - clc) * ex * (lude_i * rlu1 - lude * d.lu1 * rlu1 * rlu1);
    clc = clc + (one - clc) * (one - ex);
    qc += lude;
    qc_i += lude_i;
  }

  // ---- compensating subsidence (TL :327-373)
  const R fac1 = rcp(p.RD * t0);
  const R rho = in.ap * fac1;
  const R rho_i = (d.ap - in.ap * t_i * (p.RD * fac1)) * fac1;
  const R rodqsdp = -rho * in.qsat * fac2;
  const R rodqsdp_i = (-rho_i * in.qsat - rho * d.qsat + rho * in.qsat * (d.ap - p.RETV * foeew_i) * fac2) * fac2;
  const R ldcp = fwat * lvdcp + (one - fwat) * lsdcp;
  const R ldcp_i = fwat_i * (lvdcp - lsdcp) + fwat * lvdcp_i + (one - fwat) * lsdcp_i;
  const R fac3 = rcp(one + ldcp * dqsdtemp);
  const R dtdzmo = p.RG * (p.rcpd - ldcp * rodqsdp) * fac3;
  const R dtdzmo_i =
      -(p.RG * (ldcp_i * rodqsdp + ldcp * rodqsdp_i) + dtdzmo * (ldcp_i * dqsdtemp + ldcp * dqsdtemp_i)) * fac3;
  const R dqsdz = dqsdtemp * dtdzmo - p.RG * rodqsdp;
  const R dqsdz_i = dqsdtemp_i * dtdzmo + dqsdtemp * dtdzmo_i - p.RG * rodqsdp_i;
  const R fac4 = p.RD * t0 * rap;
  const R mfsum = in.mfu + in.mfd;
  const R sub = p.dt * dqsdz * mfsum * fac4;
  if (sub < qc) {
    R dqc_i = (p.dt * (dqsdz_i * mfsum + dqsdz * (d.mfu + d.mfd)) - sub * rho_i) * fac4;
    if (p.lregcl) dqc_i *= R(0.1);
    qc -= sub;
    qc_i -= dqc_i;
  } else {
    qc = zero;  // qc - qc
    qc_i = zero;
  }

  // ---- liquid / ice and condensation rates (TL :375-386)
  const R qlwc1 = qc * fwat, qiwc1 = qc * (one - fwat);
  R qlwc_i = qc_i * fwat + qc * fwat_i;
  R qiwc_i = qc_i * (one - fwat) - qc * fwat_i;
  R condl = (qlwc1 - ql0) * p.rdt, condi = (qiwc1 - qi0) * p.rdt;
  R condl_i = (qlwc_i - ql_i) * p.rdt, condi_i = (qiwc_i - qi_i) * p.rdt;

  // ---- maximum overlap (TL :388-397): only the carry matters while the evaporation branch is off
  if (clc > c.covptot) {
    c.covptot = clc;
    ci.covptot = clc_i;
  }
  R covpclr = zero, covpclr_i = zero;
  if (C::EVAP) {
    covpclr = c.covptot - clc;
    covpclr_i = ci.covptot - clc_i;
    if (covpclr < zero) {
      covpclr = zero;
      covpclr_i = zero;
    }
  }

  // ---- melting of incoming snow (TL :399-427)
  R rfln = c.rfl, sfln = c.sfl, rfln_i = ci.rfl, sfln_i = ci.sfl;
  R tmelt = t0;
  if (c.sfl != zero) {
    const R cons = p.cons2 * dp * rlfdcp;
    const R rcons = lfdcp * p.rgdt * rdp;
    const R cons_i = cons * (dp_i * rdp - lfdcp_i * rlfdcp);
    const bool warm2 = t0 > p.meltp2;
    const R z2s = warm2 ? cons * (t0 - p.meltp2) : zero;
    const R z2s_i = warm2 ? cons_i * (t0 - p.meltp2) + cons * t_i : zero;
    const bool allm = c.sfl <= z2s;
    const R snmlt = allm ? c.sfl : z2s;
    const R snmlt_i = allm ? ci.sfl : z2s_i;
    rfln = c.rfl + snmlt;
    rfln_i = ci.rfl + snmlt_i;
    sfln = c.sfl - snmlt;
    sfln_i = ci.sfl - snmlt_i;
    tmelt = t0 - snmlt * rcons;
    t_i -= (snmlt_i * cons - snmlt * cons_i) * rcons * rcons;
  }

  // ---- autoconversion (TL :429-503)
  R qlwc = qlwc1, qiwc = qiwc1, prr = zero, prs = zero, prr_i = zero, prs_i = zero;
  if (clc > p.ZEPS2) {
    const R rclc = rcp(clc);
    const R cldl = qlwc1 * rclc;
    const R cldl_i = qlwc_i * rclc - qlwc1 * clc_i * rclc * rclc;
    const R xl = cldl * p.rlcrit;
    const R ltmp1 = exp_(-(xl * xl));
    const R ltmp2 = exp_(-(p.ckcodtl * (one - ltmp1)));
    const R dl_i = R(2) * p.ckl_tl * p.rlcrit * p.rlcrit * ltmp1 * cldl * cldl_i;
    qlwc = clc * cldl * ltmp2;
    const R qlnew_i = clc_i * cldl * ltmp2 + clc * cldl_i * ltmp2 - qlwc * dl_i;
    prr = qlwc1 - qlwc;
    prr_i = qlwc_i - qlnew_i;
    qlwc_i = qlnew_i;
    const R cldi = qiwc1 * rclc;
    const R cldi_i = qiwc_i * rclc - qiwc1 * clc_i * rclc * rclc;
    const R xi = cldi * p.ricrit;
    const R itmp11 = exp_(-(xi * xi));
    const R itmp12 = exp_(R(0.025) * (tmelt - p.RTT));
    const R itmp2 = exp_(-(p.ckcodti * itmp12 * (one - itmp11)));
    const R di_i =
        p.cki_tl * itmp12 * (itmp11 * (R(2) * cldi * cldi_i * p.ricrit * p.ricrit - R(0.025) * t_i) + R(0.025) * t_i);
    qiwc = clc * cldi * itmp2;
    const R qinew_i = clc_i * cldi * itmp2 + clc * cldi_i * itmp2 - qiwc * di_i;
    prs = qiwc1 - qiwc;
    prs_i = qiwc_i - qinew_i;
    qiwc_i = qinew_i;
  }

  // ---- new precipitation and its phase (TL :505-523)
  const R dr1 = p.cons2 * dp * (prr + prs);
  const R dr_i = p.cons2 * (dp_i * (prr + prs) + dp * (prr_i + prs_i));
  R rfreeze = zero, rfreeze_i = zero;
  if (tmelt < p.RTT) {
    rfreeze = p.cons2 * dp * prr;
    rfreeze_i = p.cons2 * (dp_i * prr + dp * prr_i);
    sfln += dr1;
    sfln_i += dr_i;
  } else {
    rfln += dr1;
    rfln_i += dr_i;
  }

  // ---- precipitation evaporation (TL :525-616) -- dead unless LEVAPLS2 or LDRAIN1D
  R evapr = zero, evapr_i = zero, evaps = zero, evaps_i = zero;
  o.covptot = zero;
  oi.covptot = zero;
  if (C::EVAP) {
    const R prtot = rfln + sfln, prtot_i = rfln_i + sfln_i;
    if (prtot > p.ZEPS2 && covpclr > p.ZEPS2) {
      // The trajectory statements are written exactly as in level_fwd (IEEE divisions, same association): when all
      // precipitation evaporates the fluxes must come out as exact zeros like the reference's, because `sfl != 0`
      // decides the melting branch of the level below.  Only the tangent statements use shared reciprocals.
      const R rcov = rcp(c.covptot);
      R preclr = prtot * covpclr / c.covptot;
      R preclr_i = (prtot_i * covpclr + prtot * covpclr_i) * rcov - prtot * covpclr * ci.covptot * rcov * rcov;
      // humidity in the moistest covpclr region
      const R omc = one - clc;
      const R romc2 = rcp(omc * omc);
      const R dqs = in.qsat - qlim;
      const R qe = in.qsat - dqs * covpclr / (omc * omc);
      const R qe_i = d.qsat - (d.qsat * covpclr - qlim_i * covpclr + dqs * covpclr_i) * romc2 -
                     R(2) * dqs * covpclr * clc_i * romc2 * rcp(omc);
      const R tmp6 = sqrt_(in.ap / aph_s);
      const R rcovpclr = rcp(covpclr);
      const R beta = p.RG * p.RPECONS * pow_(tmp6 / R(0.00509) * preclr / covpclr, R(0.5777));
      const R beta_i =
          R(0.5777) * p.RG * p.RPECONS / R(0.00509) * pow_(R(0.00509) * covpclr / (tmp6 * preclr), R(0.4223)) *
          ((tmp6 * preclr_i + R(0.5) * preclr * d.ap / tmp6 - R(0.5) * preclr * tmp6 * aph_s_i / aph_s) * rcovpclr -
           tmp6 * preclr * covpclr_i * rcovpclr * rcovpclr);
      // implicit solution; `dt * dt` is the reference's (TL :577-579), the derivative of b would have `dt`
      const R rden = rcp(one + p.dt * beta * corqs);
      const R b = p.dt * beta * (in.qsat - qe) / (one + p.dt * beta * corqs);
      const R b_i = p.dt * (beta_i * (in.qsat - qe) + beta * (d.qsat - qe_i)) * rden -
                    p.dt * p.dt * b * (beta_i * corqs + beta * corqs_i) * rden;
      const R dtgdp = p.dt * p.RG * rdp;
      const R rdtgdp = rcp(dtgdp);
      const R dtgdp_i = -p.rgdt * dp_i * rdp * rdp;
      R dpr = covpclr * b / dtgdp;
      R dpr_i = (covpclr_i * b + covpclr * b_i) * rdtgdp - covpclr * b * dtgdp_i * rdtgdp * rdtgdp;
      if (dpr > preclr) {
        dpr = preclr;
        dpr_i = preclr_i;
      }
      preclr -= dpr;
      preclr_i -= dpr_i;
      if (preclr <= zero) {
        c.covptot = clc;
        ci.covptot = clc_i;
      }
      o.covptot = c.covptot;
      oi.covptot = ci.covptot;
      const R rpr = rcp(prtot);
      evapr = dpr * rfln / prtot;  // warm proportion
      evapr_i = (dpr_i * rfln + dpr * rfln_i) * rpr - dpr * rfln * prtot_i * rpr * rpr;
      evaps = dpr * sfln / prtot;  // ice proportion
      evaps_i = (dpr_i * sfln + dpr * sfln_i) * rpr - dpr * sfln * prtot_i * rpr * rpr;
      rfln -= evapr;
      rfln_i -= evapr_i;
      sfln -= evaps;
      sfln_i -= evaps_i;
    }
  }

  // ---- first-guess T and q (TL :618-659)
  const R dlv = lsdcp - lvdcp, dlv_i = lsdcp_i - lvdcp_i;
  const R src = C::EVAP ? in.lude + evapr + evaps : in.lude;          // moisture sources per unit gdp
  const R src_i = C::EVAP ? d.lude + evapr_i + evaps_i : d.lude;
  const R lev = C::EVAP ? lvdcp * evapr + lsdcp * evaps : zero;       // latent heat of the evaporated precipitation
  const R lev_i = C::EVAP ? lvdcp_i * evapr + lvdcp * evapr_i + lsdcp_i * evaps + lsdcp * evaps_i : zero;
  const R dqdt = -(condl + condi) + src * gdp;
  const R dqdt_i = -(condl_i + condi_i) + src_i * gdp + src * gdp_i;
  const R dtdt = lvdcp * condl + lsdcp * condi - (lev + in.lude * ldcp - dlv * rfreeze) * gdp;
  const R dtdt_i = lvdcp_i * condl + lvdcp * condl_i + lsdcp_i * condi + lsdcp * condi_i -
                   (lev_i + d.lude * ldcp + in.lude * ldcp_i - dlv_i * rfreeze - dlv * rfreeze_i) * gdp -
                   (lev + in.lude * ldcp - dlv * rfreeze) * gdp_i;
  R t = tmelt + p.dt * dtdt;
  t_i += p.dt * dtdt_i;
  const R qa = q0 + p.dt * dqdt;
  q_i += p.dt * dqdt_i;
  const R qa_i = q_i;
  R q = qa;

  // ---- saturation adjustment (TL :662)
  const bool warmc = t > p.RTT;
  const R z3c = warmc ? p.R3LES : p.R3IES, z4c = warmc ? p.R4LES : p.R4IES;
  const R z5c = warmc ? p.R5ALVCP : p.R5ALSCP, zalc = warmc ? p.RALVDCP : p.RALSDCP;
  adj_step_fwd_tl(p, rap, d.ap, z3c, z4c, z5c, zalc, t, q, t_i, q_i);
  adj_step_fwd_tl(p, rap, d.ap, z3c, z4c, z5c, zalc, t, q, t_i, q_i);

  // ---- after the adjustment (TL :664-703)
  R dq = zero, dq_i = zero;
  if (qa >= q) {
    dq = qa - q;
    dq_i = qa_i - q_i;
    if (p.lregcl) dq_i *= R(0.7);
  }
  const R dr2 = p.cons2 * dp * dq;
  const R dr2_i = p.cons2 * (dp_i * dq + dp * dq_i);
  if (t < p.RTT) {
    rfreeze_i += fwat_i * dr2 + fwat * dr2_i;
    rfreeze += fwat * dr2;
    condi += dq * p.rdt;
    condi_i += dq_i * p.rdt;
    sfln += dr2;
    sfln_i += dr2_i;
  } else {
    condl += dq * p.rdt;
    condl_i += dq_i * p.rdt;
    rfln += dr2;
    rfln_i += dr2_i;
  }

  // ---- outputs (TL :705-753)
  o.clc = clc;
  oi.clc = clc_i;
  o.tnd_q = -(condl + condi) + src * gdp;
  oi.tnd_q = -(condl_i + condi_i) + src_i * gdp + src * gdp_i;
  o.tnd_t = lvdcp * condl + lsdcp * condi - (lev + in.lude * ldcp - dlv * rfreeze) * gdp;
  oi.tnd_t = lvdcp_i * condl + lvdcp * condl_i + lsdcp_i * condi + lsdcp * condi_i -
             (lev_i + d.lude * ldcp + in.lude * ldcp_i - dlv_i * rfreeze - dlv * rfreeze_i) * gdp -
             (lev + in.lude * ldcp - dlv * rfreeze) * gdp_i;
  o.tnd_ql = (qlwc - ql0) * p.rdt;
  oi.tnd_ql = (qlwc_i - ql_i) * p.rdt;
  o.tnd_qi = (qiwc - qi0) * p.rdt;
  oi.tnd_qi = (qiwc_i - qi_i) * p.rdt;
  c.rfl = rfln;
  c.sfl = sfln;
  ci.rfl = rfln_i;
  ci.sfl = sfln_i;
}

}  // namespace cs2
