// Host-callable column sweeps of the EXPERIMENTS (not part of the shipped library; built with -DCS2_EXPERIMENTS and by the
// host twin, which keeps their level functions under test against the oracle): the two-warp split (cs2_physics_split.cuh)
// and the software-pipelined sweep (cs2_physics_pipe.cuh).  Measurements: profiles/r1f_tl_fused_nl_split.md,
// profiles/r2b_nl_pipeline.md.
#pragma once

#include "../cs2_columns.cuh"
#include "cs2_physics_pipe.cuh"
#include "cs2_physics_split.cuh"

namespace cs2 {

// NL column evaluated through the two half-level functions of the split kernel (cs2_physics_split.cuh); host twin only
template <class R, class C>
CS2_HD void column_nl_split(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f, int64_t S, int nlev,
                            int64_t i) {
  const int jsel = tropopause_candidate(p, tab, f.t, f.tnd_t, S, i);
  const int ncand = tab.nw + 1;
  Carry<R> c{R(0), R(0), R(0)};
  R aph0 = f.aph[i];
  f.fhpsl[i] = R(0);
  f.fhpsn[i] = R(0);
  for (int k = 0; k < nlev; ++k) {
    LevelIn<R> in;
    load_level(f, S, i, k, aph0, in);
    Mid<R> m;
    LevelOut<R> o;
    level_nl_a<R, C>(p, in, tab.scalm[k], tab.crh2[k * ncand + jsel], k < nlev - 1, m);
    level_nl_b<R>(p, m, c, o);
    const uint32_t off = uint32_t(k) * uint32_t(S) + uint32_t(i);
    f.clc[off] = o.clc;
    f.covptot[off] = o.covptot;
    f.o_tnd_q[off] = o.tnd_q;
    f.o_tnd_qi[off] = o.tnd_qi;
    f.o_tnd_ql[off] = o.tnd_ql;
    f.o_tnd_t[off] = o.tnd_t;
    const uint32_t offn = off + uint32_t(S);
    f.fplsl[offn] = c.rfl;
    f.fplsn[offn] = c.sfl;
    f.fhpsl[offn] = -c.rfl * p.RLVTT;
    f.fhpsn[offn] = -c.sfl * p.RLSTT;
    aph0 = in.aph1;
  }
}

// NL column through the software-pipelined level function (cs2_physics_pipe.cuh), same iteration structure as the device
// sweep dev_column_nl_pipe; host twin only
template <class R>
CS2_HD void column_nl_pipe(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f, int64_t S, int nlev,
                           int64_t i, bool ad_ref, int32_t* jsel_out) {
  const int jsel = tropopause_candidate(p, tab, f.t, f.tnd_t, S, i);
  if (jsel_out) jsel_out[i] = jsel;
  const int ncand = tab.nw + 1;
  Carry<R> c{R(0), R(0), R(0)};
  R aph0 = f.aph[i];
  f.fhpsl[i] = R(0);
  f.fhpsn[i] = R(0);
  if (jsel_out) {
    f.fplsl[i] = R(0);
    f.fplsn[i] = R(0);
  }
  PipeMid<R> m = pipe_mid_idle(p);
  LevelIn<R> in;
  for (int k = -1; k < nlev; ++k) {
    const bool has_a = k + 1 < nlev;
    const int ka = has_a ? k + 1 : nlev - 1;
    if (has_a) load_level(f, S, i, ka, aph0, in);
    PipeMid<R> mn;
    PipeOutA<R> oa;
    PipeOutB<R> ob;
    pipe_step<R>(p, m, c, ad_ref, ob, in, tab.scalm[ka], tab.crh2[ka * ncand + jsel], ka < nlev - 1, mn, oa);
    if (k < 0) c.rfl = c.sfl = R(0);
    const uint32_t off = uint32_t(k + 1) * uint32_t(S) + uint32_t(i);
    if (k >= 0) {
      const uint32_t offb = off - uint32_t(S);
      f.o_tnd_q[offb] = ob.tnd_q;
      f.o_tnd_t[offb] = ob.tnd_t;
      f.o_tnd_qi[offb] = ob.tnd_qi;
      f.covptot[offb] = R(0);
      f.fplsl[off] = c.rfl;
      f.fplsn[off] = c.sfl;
      f.fhpsl[off] = -c.rfl * p.RLVTT;
      f.fhpsn[off] = -c.sfl * p.RLSTT;
    }
    if (has_a) {
      f.clc[off] = oa.clc;
      f.o_tnd_ql[off] = oa.tnd_ql;
    }
    m = mn;
    aph0 = in.aph1;
  }
}

}  // namespace cs2
