// Device sweep of the software-pipelined NL experiment (see cs2_physics_pipe.cuh, profiles/r2b_nl_pipeline.md).
#pragma once

#include "../cs2_device_columns.cuh"
#include "cs2_physics_pipe.cuh"

namespace cs2 {

// ---------------------------------------------------------------------------------------
// NL (default flags), software-pipelined: iteration k finishes level k (half B, needs the fluxes from above) and starts
// level k+1 (half A) in one basic block -- cs2_physics_pipe.cuh.  Iteration -1 is the prologue (A of level 0, B idle),
// the last iteration repeats A of the bottom level without storing it, so there is a single copy of the level code.
// Stores per iteration: tnd_q, tnd_t, tnd_qi, covptot and the four fluxes of level k; clc and tnd_ql of level k+1.
// ---------------------------------------------------------------------------------------
template <class R, int BLOCK>
__device__ __forceinline__ void dev_column_nl_pipe(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                                   const Streams<R, I_NL>& in_s, Ring<R, I_NL, BLOCK>& ring, uint32_t S,
                                                   int nlev, uint32_t i, bool valid, bool ad_ref, int32_t* jsel_out) {
  ring_issue(ring, in_s, i);  // level 0 is in flight while the tropopause scan runs
  const int jsel = tropopause_candidate(p, tab, f.t, f.tnd_t, int64_t(S), int64_t(i));
  if (jsel_out && valid) jsel_out[i] = jsel;
  const int ncand = tab.nw + 1;
  Carry<R> c{R(0), R(0), R(0)};
  R aph0 = f.aph[i];
  if (valid) {
    f.fhpsl[i] = R(0);
    f.fhpsn[i] = R(0);
    if (jsel_out) {  // the AD stencil also writes the level-0 precipitation fluxes (AD :466-470)
      f.fplsl[i] = R(0);
      f.fplsn[i] = R(0);
    }
  }
  PipeMid<R> m = pipe_mid_idle(p);
  LevelIn<R> in;
  // level-only table values of the level whose half A runs next, loaded one iteration ahead (an L2 hit on the chain otherwise)
  const R* crh2_col = tab.crh2 + jsel;
  R scalm_a = tab.scalm[0], crh2_a = crh2_col[0];
  for (int k = -1; k < nlev; ++k) {
    const bool has_a = k + 1 < nlev;
    const int ka = has_a ? k + 1 : nlev - 1;
    const int kn = (ka + 1 < nlev) ? ka + 1 : nlev - 1;
    if (has_a) {
      cp_async_wait_all();
      ring_read_level(ring, 0, aph0, in);
      if (k + 2 < nlev) ring_issue(ring, in_s, uint32_t(k + 2) * S + i);
    }
    const R scalm_n = tab.scalm[kn], crh2_n = crh2_col[kn * ncand];
    PipeMid<R> mn;
    PipeOutA<R> oa;
    PipeOutB<R> ob;
    pipe_step<R>(p, m, c, ad_ref, ob, in, scalm_a, crh2_a, ka < nlev - 1, mn, oa);
    scalm_a = scalm_n;
    crh2_a = crh2_n;
    if (k < 0) c.rfl = c.sfl = R(0);
    if (valid) {
      const uint32_t off = uint32_t(k + 1) * S + i;  // level k+1 (full-level index) == half level k+1
      if (k >= 0) {
        const uint32_t offb = off - S;
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
    }
    m = mn;
    aph0 = in.aph1;
  }
}

}  // namespace cs2
