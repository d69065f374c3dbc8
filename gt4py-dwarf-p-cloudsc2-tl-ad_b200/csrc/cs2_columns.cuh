// Column sweeps: one thread (GPU) / one loop iteration (host twin) owns one column and walks
// the vertical with the precipitation fluxes and the overlap carried in registers.
//
// Field storage is `[nlev+1][S]` with the column index fastest (S = ncol_stride), so the 32
// threads of a warp touch 32 consecutive elements of one level: one fully coalesced request
// per field per level.
#pragma once

#include "cs2_physics.cuh"
#include "cs2_physics_tl.cuh"

namespace cs2 {

template <class R>
struct NLFields {
  const R *ap, *aph, *lu, *lude, *mfd, *mfu, *q, *qi, *ql, *qsat, *supsat, *t, *tnd_q, *tnd_qi, *tnd_ql, *tnd_t;
  R *clc, *covptot, *fhpsl, *fhpsn, *fplsl, *fplsn, *o_tnd_q, *o_tnd_qi, *o_tnd_ql, *o_tnd_t;
};

template <class R>
inline NLFields<R> make_nl_fields(const cs2_nl_fields& f) {
  NLFields<R> o;
#define CS2_I(n) o.n = static_cast<const R*>(f.in_##n)
  CS2_I(ap); CS2_I(aph); CS2_I(lu); CS2_I(lude); CS2_I(mfd); CS2_I(mfu); CS2_I(q); CS2_I(qi);
  CS2_I(ql); CS2_I(qsat); CS2_I(supsat); CS2_I(t);
#undef CS2_I
  o.tnd_q = static_cast<const R*>(f.in_tnd_cml_q);
  o.tnd_qi = static_cast<const R*>(f.in_tnd_cml_qi);
  o.tnd_ql = static_cast<const R*>(f.in_tnd_cml_ql);
  o.tnd_t = static_cast<const R*>(f.in_tnd_cml_t);
#define CS2_O(n) o.n = static_cast<R*>(f.out_##n)
  CS2_O(clc); CS2_O(covptot); CS2_O(fhpsl); CS2_O(fhpsn); CS2_O(fplsl); CS2_O(fplsn);
#undef CS2_O
  o.o_tnd_q = static_cast<R*>(f.out_tnd_q);
  o.o_tnd_qi = static_cast<R*>(f.out_tnd_qi);
  o.o_tnd_ql = static_cast<R*>(f.out_tnd_ql);
  o.o_tnd_t = static_cast<R*>(f.out_tnd_t);
  return o;
}

template <class R>
struct ADSeeds {
  R *tnd_t, *tnd_q, *tnd_ql, *tnd_qi, *clc, *covptot, *fhpsl, *fhpsn, *fplsl, *fplsn;
};
template <class R>
struct ADOut {
  R *aph, *ap, *q, *qsat, *t, *ql, *qi, *lude, *lu, *mfu, *mfd, *supsat, *tnd_t, *tnd_q, *tnd_ql, *tnd_qi;
};

// Tropopause candidate of a column: index j of the LAST window level wlev[j-1] with
// t[k] > t[k+1] (0 = none -> trpaus = 0.1); nonlinear/_stencils/cloudsc2.py:106-111.
template <class R>
CS2_HD int tropopause_candidate(const DevParams<R>& p, const LevelTables<R>& tab, const R* CS2_RESTRICT t,
                                const R* CS2_RESTRICT tnd_t, int64_t S, int64_t i) {
  int jsel = 0;
  int kprev = -2;
  R tnext = R(0);
  for (int j = 0; j < tab.nw; ++j) {
    const int k = tab.wlev[j];
    const R tk = (k == kprev + 1) ? tnext : (t[k * S + i] + p.dt * tnd_t[k * S + i]);
    tnext = t[(k + 1) * S + i] + p.dt * tnd_t[(k + 1) * S + i];
    kprev = k;
    if (tk > tnext) jsel = j + 1;
  }
  return jsel;
}

template <class R>
CS2_HD void load_level(const NLFields<R>& f, int64_t S, int64_t i, int k, R aph0, LevelIn<R>& in) {
  // 32-bit element offsets ((nlev+1) * ncol_stride < 2^32 is checked by the launcher): one IMAD.WIDE per address
  const uint32_t o = uint32_t(k) * uint32_t(S) + uint32_t(i);
  const uint32_t S32 = uint32_t(S);
  in.ap = f.ap[o];
  in.aph0 = aph0;
  in.aph1 = f.aph[o + S32];
  in.lu1 = f.lu[o + S32];
  in.lude = f.lude[o];
  in.mfd = f.mfd[o];
  in.mfu = f.mfu[o];
  in.q = f.q[o];
  in.qi = f.qi[o];
  in.ql = f.ql[o];
  in.qsat = f.qsat[o];
  in.supsat = f.supsat[o];
  in.t = f.t[o];
  in.tnd_q = f.tnd_q[o];
  in.tnd_qi = f.tnd_qi[o];
  in.tnd_ql = f.tnd_ql[o];
  in.tnd_t = f.tnd_t[o];
}

// ---------------------------------------------------------------------------------------
// NL column (also the forward sweep of AD: `jsel_out` keeps the tropopause candidate)
// ---------------------------------------------------------------------------------------
// LIN: keep the linearisation-friendly form of the trajectory (AD forward sweep); false for plain NL
// cov_out (AD forward sweep with the evaporation branch): the overlap carry entering each level, [nlev][S]
template <class R, class C, bool LIN = false>
CS2_HD void column_nl(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f, int64_t S, int nlev,
                      int64_t i, bool ad_ref, int32_t* jsel_out, R* cov_out = nullptr) {
  const int jsel = tropopause_candidate(p, tab, f.t, f.tnd_t, S, i);
  if (jsel_out) jsel_out[i] = jsel;
  const int ncand = tab.nw + 1;

  Carry<R> c{R(0), R(0), R(0)};
  const R aph_s = f.aph[int64_t(nlev) * S + i];
  R aph0 = f.aph[i];
  // half level 0: enthalpy fluxes are zero (:391-394); NL leaves fplsl/fplsn[0] untouched,
  // the AD stencil (jsel_out != nullptr) writes them too (AD :466-470)
  f.fhpsl[i] = R(0);
  f.fhpsn[i] = R(0);
  if (jsel_out) {
    f.fplsl[i] = R(0);
    f.fplsn[i] = R(0);
  }
  for (int k = 0; k < nlev; ++k) {
    LevelIn<R> in;
    load_level(f, S, i, k, aph0, in);
    LevelOut<R> o;
    Traj<R> tr;
    Trans<R, 0> x;
    const uint32_t off = uint32_t(k) * uint32_t(S) + uint32_t(i);
    if (cov_out) cov_out[off] = c.covptot;
    level_fwd<R, C, LIN>(p, in, tab.scalm[k], tab.crh2[k * ncand + jsel], k < nlev - 1, aph_s, ad_ref, c, o, tr, x);
    f.clc[off] = o.clc;
    f.covptot[off] = o.covptot;
    f.o_tnd_q[off] = o.tnd_q;
    f.o_tnd_qi[off] = o.tnd_qi;
    f.o_tnd_ql[off] = o.tnd_ql;
    f.o_tnd_t[off] = o.tnd_t;
    // fluxes shifted one half level down (:395-399)
    const uint32_t offn = off + uint32_t(S);
    f.fplsl[offn] = c.rfl;
    f.fplsn[offn] = c.sfl;
    f.fhpsl[offn] = -c.rfl * p.RLVTT;
    f.fhpsn[offn] = -c.sfl * p.RLSTT;
    aph0 = in.aph1;
  }
}

// ---------------------------------------------------------------------------------------
// TL column: trajectory and perturbation together
// ---------------------------------------------------------------------------------------
// EVAP: LEVAPLS2 or LDRAIN1D (the precipitation-evaporation branch and its tangent, level_fwd_tl only)
template <class R, bool EVAP = false>
CS2_HD void column_tl(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f, const NLFields<R>& g,
                      int64_t S, int nlev, int64_t i) {
  using C = Cfg<EVAP, true>;
  const int jsel = tropopause_candidate(p, tab, f.t, f.tnd_t, S, i);
  const int ncand = tab.nw + 1;

  Carry<R> c{R(0), R(0), R(0)}, ci{R(0), R(0), R(0)};
  const R aph_s = f.aph[int64_t(nlev) * S + i], aph_s_i = g.aph[int64_t(nlev) * S + i];
  R aph0 = f.aph[i], aph0_i = g.aph[i];
  // half level 0 (TL :757-765)
  f.fplsl[i] = R(0); f.fplsn[i] = R(0); f.fhpsl[i] = R(0); f.fhpsn[i] = R(0);
  g.fplsl[i] = R(0); g.fplsn[i] = R(0); g.fhpsl[i] = R(0); g.fhpsn[i] = R(0);
  for (int k = 0; k < nlev; ++k) {
    LevelIn<R> in, d;
    load_level(f, S, i, k, aph0, in);
    load_level(g, S, i, k, aph0_i, d);
    LevelOut<R> o, oi;
#if defined(CS2_TL_SPLIT)  // the two-pass specification (level_fwd, then level_tl about its trajectory)
    Traj<R> tr;
    Trans<R, 0> x;
    level_fwd<R, C, true>(p, in, tab.scalm[k], tab.crh2[k * ncand + jsel], k < nlev - 1, aph_s, false, c, o, tr, x);
    level_tl<R>(p, in, d, tr, ci, oi);
    (void)aph_s_i;
#else
    level_fwd_tl<R, C>(p, in, d, tab.scalm[k], tab.crh2[k * ncand + jsel], k < nlev - 1, aph_s, aph_s_i, c, ci, o, oi);
#endif
    const uint32_t off = uint32_t(k) * uint32_t(S) + uint32_t(i);
    const uint32_t offn = off + uint32_t(S);
    f.clc[off] = o.clc;          g.clc[off] = oi.clc;
    f.covptot[off] = o.covptot;  g.covptot[off] = oi.covptot;
    f.o_tnd_q[off] = o.tnd_q;    g.o_tnd_q[off] = oi.tnd_q;
    f.o_tnd_qi[off] = o.tnd_qi;  g.o_tnd_qi[off] = oi.tnd_qi;
    f.o_tnd_ql[off] = o.tnd_ql;  g.o_tnd_ql[off] = oi.tnd_ql;
    f.o_tnd_t[off] = o.tnd_t;    g.o_tnd_t[off] = oi.tnd_t;
    f.fplsl[offn] = c.rfl;            g.fplsl[offn] = ci.rfl;
    f.fplsn[offn] = c.sfl;            g.fplsn[offn] = ci.sfl;
    f.fhpsl[offn] = -c.rfl * p.RLVTT; g.fhpsl[offn] = -ci.rfl * p.RLVTT;
    f.fhpsn[offn] = -c.sfl * p.RLSTT; g.fhpsn[offn] = -ci.sfl * p.RLSTT;
    aph0 = in.aph1;
    aph0_i = d.aph1;
  }
}

// ---------------------------------------------------------------------------------------
// AD backward column (recompute variant): the level-entry fluxes are read back from the
// trajectory outputs fplsl/fplsn that the forward sweep (column_nl) has just written, the rest
// of the level trajectory is recomputed from the inputs.
// ---------------------------------------------------------------------------------------
// EVAP: LEVAPLS2 or LDRAIN1D; cov_in = the overlap carry entering each level, written by the forward sweep
template <class R, bool EVAP = false>
CS2_HD void column_ad_bwd(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                          const ADSeeds<R>& s, const ADOut<R>& a, const int32_t* jsel_in, int64_t S, int nlev,
                          int64_t i, const R* cov_in = nullptr) {
  using C = Cfg<EVAP, true>;
  const bool ad_ref = !p.ad_tl_predicates;
  const int jsel = jsel_in[i];
  const int ncand = tab.nw + 1;
  const R aph_s = f.aph[int64_t(nlev) * S + i];

  R a_rfl = R(0), a_sfl = R(0);   // adjoint of the fluxes entering the level below
  R a_dp_below = R(0);            // a_dp of level k+1 (0 below the surface)
  R a_cov = R(0), a_aph_s = R(0); // evaporation branch: adjoint of the overlap carry, adjoint of the surface pressure
  R aph1 = aph_s;
  for (int k = nlev - 1; k >= 0; --k) {
    const uint32_t off = uint32_t(k) * uint32_t(S) + uint32_t(i);
    const uint32_t offn = off + uint32_t(S);
    LevelIn<R> in;
    const R aph0 = f.aph[off];
    load_level(f, S, i, k, aph0, in);
    in.aph1 = aph1;
    Carry<R> c;
    c.rfl = (k > 0) ? f.fplsl[off] : R(0);
    c.sfl = (k > 0) ? f.fplsn[off] : R(0);
    c.covptot = EVAP ? cov_in[off] : R(0);  // only feeds the evaporation branch
    LevelOut<R> o;
    Traj<R> tr;
    Trans<R, 0> x;
    level_fwd<R, C, true>(p, in, tab.scalm[k], tab.crh2[k * ncand + jsel], k < nlev - 1, aph_s, ad_ref, c, o, tr, x);

    // seeds: tendencies / cloud cover at k, fluxes at half level k+1 with the enthalpy-flux
    // seeds folded in (AD :479-484,500-501); all consumed seeds are zeroed like the reference.
    LevelOut<R> so;
    so.tnd_t = s.tnd_t[off];   s.tnd_t[off] = R(0);
    so.tnd_q = s.tnd_q[off];   s.tnd_q[off] = R(0);
    so.tnd_ql = s.tnd_ql[off]; s.tnd_ql[off] = R(0);
    so.tnd_qi = s.tnd_qi[off]; s.tnd_qi[off] = R(0);
    so.clc = s.clc[off];       s.clc[off] = R(0);
    so.covptot = EVAP ? s.covptot[off] : R(0);
    s.covptot[off] = R(0);
    R a_rfln = a_rfl + (s.fplsl[offn] - s.fhpsl[offn] * p.RLVTT);
    R a_sfln = a_sfl + (s.fplsn[offn] - s.fhpsn[offn] * p.RLSTT);
    s.fplsl[offn] = R(0); s.fhpsl[offn] = R(0);
    s.fplsn[offn] = R(0); s.fhpsn[offn] = R(0);

    LevelIn<R> ad;
    level_ad<R, C>(p, in, tr, so, ad_ref, a_rfln, a_sfln, ad, aph_s, &a_cov, &a_aph_s);
    a_rfl = a_rfln;
    a_sfl = a_sfln;

    a.t[off] = ad.t;           a.tnd_t[off] = ad.tnd_t;
    a.q[off] = ad.q;           a.tnd_q[off] = ad.tnd_q;
    a.ql[off] = ad.ql;         a.tnd_ql[off] = ad.tnd_ql;
    a.qi[off] = ad.qi;         a.tnd_qi[off] = ad.tnd_qi;
    a.supsat[off] = ad.supsat; a.qsat[off] = ad.qsat;
    a.ap[off] = ad.ap;         a.lude[off] = ad.lude;
    a.mfu[off] = ad.mfu;       a.mfd[off] = ad.mfd;
    // staggered fields (AD :969-986): aph_i[k+1] = a_dp(k) - a_dp(k+1); lu_i[k+1] = adjoint of lu[k+1]
    a.aph[offn] = ad.aph1 - a_dp_below;
    a.lu[offn] = ad.lu1;
    a_dp_below = ad.aph1;
    aph1 = aph0;
  }
  a.aph[i] = -a_dp_below;
  a.lu[i] = R(0);
  if (EVAP) a.aph[int64_t(nlev) * S + i] += a_aph_s;  // adjoint of the surface pressure (AD :974-975)
  s.fplsl[i] = R(0); s.fhpsl[i] = R(0); s.fplsn[i] = R(0); s.fhpsn[i] = R(0);
}

}  // namespace cs2
