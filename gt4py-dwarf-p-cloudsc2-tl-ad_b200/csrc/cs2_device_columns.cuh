// Device-only column sweeps with a cp.async prefetch buffer.
//
// Same level physics as cs2_columns.cuh (level_fwd / level_tl / level_ad); what differs is the
// data movement.  A thread owns a column and, while it computes level k (several thousand cycles of
// dependent FP64 work), the 16 / 32 / 27 inputs of the next level are already in flight: they are
// copied global -> shared with cp.async (LDGSTS) into a slot private to the thread
// (`v[field][tid]`, conflict-free, no __syncthreads needed: a thread only ever reads what it copied
// itself, after cp.async.wait_group 0).  One stage is enough: at the top of a level the thread moves
// its slots into registers and only then issues the copies of the next level into the same slots
// (shared-memory reads and the much later asynchronous writes of one thread stay in program order).
// This hides the HBM latency without spending a
// single register on the prefetch, which matters because the register budget decides whether all
// 65 536 columns of the headline case are resident in one wave.
#pragma once

#include "cs2_columns.cuh"

namespace cs2 {

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// N input streams, all indexed with the same element offset `k * S + i` (fields that are read at
// level k+1 are passed with their base pointer advanced by S on the host side).
template <class R, int N>
struct Streams {
  const R* p[N];
};

template <class R, int N, int BLOCK>
struct Ring {
  R v[N][BLOCK];
};

template <class R, int N, int BLOCK>
__device__ __forceinline__ void ring_issue(Ring<R, N, BLOCK>& ring, const Streams<R, N>& in, uint32_t off) {
#pragma unroll
  for (int f = 0; f < N; ++f) cp_async<sizeof(R)>(&ring.v[f][threadIdx.x], in.p[f] + off);
  cp_async_commit();
}

// order of the NL input streams
enum { I_AP, I_APH1, I_LU1, I_LUDE, I_MFD, I_MFU, I_Q, I_QI, I_QL, I_QSAT, I_SUPSAT, I_T, I_TQ, I_TQI, I_TQL, I_TT, I_NL };

template <class R>
inline Streams<R, I_NL> nl_streams(const NLFields<R>& f, int64_t S) {
  Streams<R, I_NL> s;
  s.p[I_AP] = f.ap; s.p[I_APH1] = f.aph + S; s.p[I_LU1] = f.lu + S; s.p[I_LUDE] = f.lude; s.p[I_MFD] = f.mfd;
  s.p[I_MFU] = f.mfu; s.p[I_Q] = f.q; s.p[I_QI] = f.qi; s.p[I_QL] = f.ql; s.p[I_QSAT] = f.qsat;
  s.p[I_SUPSAT] = f.supsat; s.p[I_T] = f.t; s.p[I_TQ] = f.tnd_q; s.p[I_TQI] = f.tnd_qi; s.p[I_TQL] = f.tnd_ql;
  s.p[I_TT] = f.tnd_t;
  return s;
}

template <class R, int N, int BLOCK>
__device__ __forceinline__ void ring_read_level(const Ring<R, N, BLOCK>& ring, int base, R aph0, LevelIn<R>& in) {
  const int t = threadIdx.x;
  in.ap = ring.v[base + I_AP][t];
  in.aph0 = aph0;
  in.aph1 = ring.v[base + I_APH1][t];
  in.lu1 = ring.v[base + I_LU1][t];
  in.lude = ring.v[base + I_LUDE][t];
  in.mfd = ring.v[base + I_MFD][t];
  in.mfu = ring.v[base + I_MFU][t];
  in.q = ring.v[base + I_Q][t];
  in.qi = ring.v[base + I_QI][t];
  in.ql = ring.v[base + I_QL][t];
  in.qsat = ring.v[base + I_QSAT][t];
  in.supsat = ring.v[base + I_SUPSAT][t];
  in.t = ring.v[base + I_T][t];
  in.tnd_q = ring.v[base + I_TQ][t];
  in.tnd_qi = ring.v[base + I_TQI][t];
  in.tnd_ql = ring.v[base + I_TQL][t];
  in.tnd_t = ring.v[base + I_TT][t];
}

// ---------------------------------------------------------------------------------------
// NL (and AD forward when jsel_out != nullptr)
// ---------------------------------------------------------------------------------------
// CKPT: also record the level's transcendental results into `ck` ([CK_N][nlev][S], CS2_AD_CHECKPOINT).
template <class R, class C, int BLOCK, bool CKPT, bool LIN>
__device__ __forceinline__ void dev_column_nl(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                              const Streams<R, I_NL>& in_s, Ring<R, I_NL, BLOCK>& ring, uint32_t S,
                                              int nlev, uint32_t i, bool valid, bool ad_ref, int32_t* jsel_out, R* ck,
                                              R* cov_out = nullptr) {
  ring_issue(ring, in_s, i);  // level 0 is in flight while the tropopause scan runs
  const int jsel = tropopause_candidate(p, tab, f.t, f.tnd_t, int64_t(S), int64_t(i));
  if (jsel_out && valid) jsel_out[i] = jsel;
  const int ncand = tab.nw + 1;

  Carry<R> c{R(0), R(0), R(0)};
  const R aph_s = f.aph[uint32_t(nlev) * S + i];
  R aph0 = f.aph[i];
  if (valid) {
    f.fhpsl[i] = R(0);
    f.fhpsn[i] = R(0);
    if (jsel_out) {  // the AD stencil also writes the level-0 precipitation fluxes (AD :466-470)
      f.fplsl[i] = R(0);
      f.fplsn[i] = R(0);
    }
  }
  for (int k = 0; k < nlev; ++k) {
    const uint32_t off = uint32_t(k) * S + i;
    cp_async_wait_all();
    LevelIn<R> in;
    ring_read_level(ring, 0, aph0, in);
    if (k + 1 < nlev) ring_issue(ring, in_s, off + S);
    if (C::EVAP && LIN && cov_out && valid) cov_out[off] = c.covptot;  // AD forward sweep with the evaporation branch only
    LevelOut<R> o;
    Traj<R> tr;
    Trans<R, CKPT ? 1 : 0> x;
    if (CKPT) x.defaults();
    level_fwd<R, C, LIN>(p, in, tab.scalm[k], tab.crh2[k * ncand + jsel], k < nlev - 1, aph_s, ad_ref, c, o, tr, x);
    if (CKPT && valid) {
      const size_t plane = size_t(nlev) * size_t(S);  // CK_N * plane can exceed 2^32 elements: 64-bit index
#pragma unroll
      for (int n = 0; n < CK_N; ++n) ck[size_t(n) * plane + off] = x.v[n];
    }
    if (valid) {
      f.clc[off] = o.clc;
      f.covptot[off] = o.covptot;
      f.o_tnd_q[off] = o.tnd_q;
      f.o_tnd_qi[off] = o.tnd_qi;
      f.o_tnd_ql[off] = o.tnd_ql;
      f.o_tnd_t[off] = o.tnd_t;
      const uint32_t offn = off + S;
      f.fplsl[offn] = c.rfl;
      f.fplsn[offn] = c.sfl;
      f.fhpsl[offn] = -c.rfl * p.RLVTT;
      f.fhpsn[offn] = -c.sfl * p.RLSTT;
    }
    aph0 = in.aph1;
  }
}

// ---------------------------------------------------------------------------------------
// Fused perturbed NL: NL of the state x + fac * x_i without materialising it (what the Taylor test's
// PerturbedState -> Cloudsc2NL pair computes; reference tangent_linear/validation.py:167-176).
// Streams [0, 16) = x, [16, 32) = x_i.  The combination is the same single FMA the stand-alone
// perturbed_state kernel performs, so the result is bit-identical to the unfused pair.
// ---------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ R axpy(R fac, R xi, R x) {
  return fac * xi + x;  // contracted to one FMA, like perturbed_state_kernel
}

template <class R>
__device__ __forceinline__ int tropopause_candidate_pert(const DevParams<R>& p, const LevelTables<R>& tab, const R* t,
                                                         const R* tnd_t, const R* t_i, const R* tnd_t_i, R fac, uint32_t S,
                                                         uint32_t i) {
  int jsel = 0;
  int kprev = -2;
  R tnext = R(0);
  for (int j = 0; j < tab.nw; ++j) {
    const int k = tab.wlev[j];
    const uint32_t o = uint32_t(k) * S + i;
    const R tk = (k == kprev + 1) ? tnext : (axpy(fac, t_i[o], t[o]) + p.dt * axpy(fac, tnd_t_i[o], tnd_t[o]));
    tnext = axpy(fac, t_i[o + S], t[o + S]) + p.dt * axpy(fac, tnd_t_i[o + S], tnd_t[o + S]);
    kprev = k;
    if (tk > tnext) jsel = j + 1;
  }
  return jsel;
}

template <class R, class C, int BLOCK>
__device__ __forceinline__ void dev_column_nl_pert(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                                   const NLFields<R>& g, R fac, const Streams<R, 2 * I_NL>& in_s,
                                                   Ring<R, 2 * I_NL, BLOCK>& ring, uint32_t S, int nlev, uint32_t i,
                                                   bool valid) {
  ring_issue(ring, in_s, i);
  const int jsel = tropopause_candidate_pert(p, tab, f.t, f.tnd_t, g.t, g.tnd_t, fac, S, i);
  const int ncand = tab.nw + 1;
  Carry<R> c{R(0), R(0), R(0)};
  const R aph_s = axpy(fac, g.aph[uint32_t(nlev) * S + i], f.aph[uint32_t(nlev) * S + i]);
  R aph0 = axpy(fac, g.aph[i], f.aph[i]);
  if (valid) {
    f.fhpsl[i] = R(0);
    f.fhpsn[i] = R(0);
  }
  for (int k = 0; k < nlev; ++k) {
    const uint32_t off = uint32_t(k) * S + i;
    cp_async_wait_all();
    LevelIn<R> in, d;
    ring_read_level(ring, 0, aph0, in);
    ring_read_level(ring, I_NL, R(0), d);
    if (k + 1 < nlev) ring_issue(ring, in_s, off + S);
    in.ap = axpy(fac, d.ap, in.ap);             in.aph1 = axpy(fac, d.aph1, in.aph1);
    in.lu1 = axpy(fac, d.lu1, in.lu1);          in.lude = axpy(fac, d.lude, in.lude);
    in.mfd = axpy(fac, d.mfd, in.mfd);          in.mfu = axpy(fac, d.mfu, in.mfu);
    in.q = axpy(fac, d.q, in.q);                in.qi = axpy(fac, d.qi, in.qi);
    in.ql = axpy(fac, d.ql, in.ql);             in.qsat = axpy(fac, d.qsat, in.qsat);
    in.supsat = axpy(fac, d.supsat, in.supsat); in.t = axpy(fac, d.t, in.t);
    in.tnd_q = axpy(fac, d.tnd_q, in.tnd_q);    in.tnd_qi = axpy(fac, d.tnd_qi, in.tnd_qi);
    in.tnd_ql = axpy(fac, d.tnd_ql, in.tnd_ql); in.tnd_t = axpy(fac, d.tnd_t, in.tnd_t);
    LevelOut<R> o;
    Traj<R> tr;
    Trans<R, 0> x;
    level_fwd<R, C, false>(p, in, tab.scalm[k], tab.crh2[k * ncand + jsel], k < nlev - 1, aph_s, false, c, o, tr, x);
    if (valid) {
      f.clc[off] = o.clc;
      f.covptot[off] = o.covptot;
      f.o_tnd_q[off] = o.tnd_q;
      f.o_tnd_qi[off] = o.tnd_qi;
      f.o_tnd_ql[off] = o.tnd_ql;
      f.o_tnd_t[off] = o.tnd_t;
      const uint32_t offn = off + S;
      f.fplsl[offn] = c.rfl;
      f.fplsn[offn] = c.sfl;
      f.fhpsl[offn] = -c.rfl * p.RLVTT;
      f.fhpsn[offn] = -c.sfl * p.RLSTT;
    }
    aph0 = in.aph1;
  }
}

// ---------------------------------------------------------------------------------------
// One factor of the Taylor test in one sweep (cs2_taylor_nl_sums): NL of the state x + f2 * (f1 * x) -- the
// StateIncrement(f1) / PerturbedState(f2) / Cloudsc2NL chain of tangent_linear/validation.py:158-176, formed with the same
// roundings (product rounded on its own, then one FMA) -- compared on the fly with the unperturbed NL outputs F_nl:
// SUM_k (F_p - F_nl) per output field and column is accumulated in fp64 in per-thread shared-memory slots, so F_p is
// never written and never re-read.  Streams [0, 16) = x, [16, 26) = F_nl in the order of T_* below.
// ---------------------------------------------------------------------------------------
enum { T_TT, T_TQ, T_TQL, T_TQI, T_CLC, T_FHPSL, T_FHPSN, T_FPLSL, T_FPLSN, T_COVPTOT, T_N };

template <class R>
inline Streams<R, I_NL + T_N> taylor_streams(const NLFields<R>& f, int64_t S) {
  Streams<R, I_NL + T_N> s;
  const Streams<R, I_NL> a = nl_streams(f, S);
  for (int n = 0; n < I_NL; ++n) s.p[n] = a.p[n];
  s.p[I_NL + T_TT] = f.o_tnd_t; s.p[I_NL + T_TQ] = f.o_tnd_q; s.p[I_NL + T_TQL] = f.o_tnd_ql; s.p[I_NL + T_TQI] = f.o_tnd_qi;
  s.p[I_NL + T_CLC] = f.clc; s.p[I_NL + T_COVPTOT] = f.covptot;
  s.p[I_NL + T_FHPSL] = f.fhpsl + S; s.p[I_NL + T_FHPSN] = f.fhpsn + S;  // half-level fields: level k+1
  s.p[I_NL + T_FPLSL] = f.fplsl + S; s.p[I_NL + T_FPLSN] = f.fplsn + S;
  return s;
}

template <class R>
__device__ __forceinline__ R pert2(R f1, R f2, R x) {
  return axpy(f2, mul_rn(f1, x), x);  // x + f2 * round(f1 * x): state_increment then perturbed_state
}

template <class R, class C, int BLOCK>
__device__ __forceinline__ void dev_column_nl_taylor(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                                     R f1, R f2, bool ignore_supsat,
                                                     const Streams<R, I_NL + T_N>& in_s, Ring<R, I_NL + T_N, BLOCK>& ring,
                                                     double (*acc)[BLOCK], uint32_t S, int nlev, uint32_t i, bool valid) {
  ring_issue(ring, in_s, i);
  const int t = threadIdx.x;
#pragma unroll
  for (int n = 0; n < T_N; ++n) acc[n][t] = 0.0;
  // tropopause candidate of the perturbed temperature profile (nonlinear/_stencils/cloudsc2.py:106-111)
  int jsel = 0;
  {
    int kprev = -2;
    R tnext = R(0);
    for (int j = 0; j < tab.nw; ++j) {
      const int k = tab.wlev[j];
      const uint32_t o = uint32_t(k) * S + i;
      const R tk = (k == kprev + 1) ? tnext : (pert2(f1, f2, f.t[o]) + p.dt * pert2(f1, f2, f.tnd_t[o]));
      tnext = pert2(f1, f2, f.t[o + S]) + p.dt * pert2(f1, f2, f.tnd_t[o + S]);
      kprev = k;
      if (tk > tnext) jsel = j + 1;
    }
  }
  const int ncand = tab.nw + 1;
  Carry<R> c{R(0), R(0), R(0)};
  const R aph_s = pert2(f1, f2, f.aph[uint32_t(nlev) * S + i]);
  R aph0 = pert2(f1, f2, f.aph[i]);
  for (int k = 0; k < nlev; ++k) {
    const uint32_t off = uint32_t(k) * S + i;
    cp_async_wait_all();
    LevelIn<R> in;
    ring_read_level(ring, 0, aph0, in);
    R ref[T_N];
#pragma unroll
    for (int n = 0; n < T_N; ++n) ref[n] = ring.v[I_NL + n][t];
    if (k + 1 < nlev) ring_issue(ring, in_s, off + S);
    in.ap = pert2(f1, f2, in.ap);         in.aph1 = pert2(f1, f2, in.aph1);     in.lu1 = pert2(f1, f2, in.lu1);
    in.lude = pert2(f1, f2, in.lude);     in.mfd = pert2(f1, f2, in.mfd);       in.mfu = pert2(f1, f2, in.mfu);
    in.q = pert2(f1, f2, in.q);           in.qi = pert2(f1, f2, in.qi);         in.ql = pert2(f1, f2, in.ql);
    in.qsat = pert2(f1, f2, in.qsat);     in.t = pert2(f1, f2, in.t);           in.tnd_q = pert2(f1, f2, in.tnd_q);
    in.tnd_qi = pert2(f1, f2, in.tnd_qi); in.tnd_ql = pert2(f1, f2, in.tnd_ql); in.tnd_t = pert2(f1, f2, in.tnd_t);
    in.supsat = ignore_supsat ? axpy(f2, R(0), in.supsat) : pert2(f1, f2, in.supsat);
    LevelOut<R> o;
    Traj<R> tr;
    Trans<R, 0> x;
    level_fwd<R, C, false>(p, in, tab.scalm[k], tab.crh2[k * ncand + jsel], k < nlev - 1, aph_s, false, c, o, tr, x);
    acc[T_TT][t] += double(o.tnd_t) - double(ref[T_TT]);
    acc[T_TQ][t] += double(o.tnd_q) - double(ref[T_TQ]);
    acc[T_TQL][t] += double(o.tnd_ql) - double(ref[T_TQL]);
    acc[T_TQI][t] += double(o.tnd_qi) - double(ref[T_TQI]);
    acc[T_CLC][t] += double(o.clc) - double(ref[T_CLC]);
    acc[T_COVPTOT][t] += double(o.covptot) - double(ref[T_COVPTOT]);
    acc[T_FPLSL][t] += double(c.rfl) - double(ref[T_FPLSL]);
    acc[T_FPLSN][t] += double(c.sfl) - double(ref[T_FPLSN]);
    acc[T_FHPSL][t] += double(-c.rfl * p.RLVTT) - double(ref[T_FHPSL]);
    acc[T_FHPSN][t] += double(-c.sfl * p.RLSTT) - double(ref[T_FHPSN]);
    aph0 = in.aph1;
  }
  if (!valid) {
#pragma unroll
    for (int n = 0; n < T_N; ++n) acc[n][t] = 0.0;  // shadow threads of a ragged last CTA contribute nothing
  }
}

// ---------------------------------------------------------------------------------------
// TL: streams [0, 16) = trajectory inputs, [16, 32) = perturbation inputs
// ---------------------------------------------------------------------------------------
template <class R>
inline Streams<R, 2 * I_NL> tl_streams(const NLFields<R>& f, const NLFields<R>& g, int64_t S) {
  Streams<R, 2 * I_NL> s;
  const Streams<R, I_NL> a = nl_streams(f, S), b = nl_streams(g, S);
  for (int n = 0; n < I_NL; ++n) {
    s.p[n] = a.p[n];
    s.p[I_NL + n] = b.p[n];
  }
  return s;
}

// INC: fused "state_increment" + "cloudsc2_tl" (cs2_tl_increment): the perturbation of every input is fac * input
// (common/_stencils/state_increment.py:60-80; supsat_i = 0 with IGNORE_SUPSAT), formed in registers -- only the 16
// trajectory inputs are streamed.
// NORM (with INC): also the symmetry test's first inner product, norm1[i] = SUM_k SUM_fields (TL output)^2 of the column
// (adjoint/validation.py:167-181), accumulated in fp64 from the values as stored.
template <class R, int BLOCK, bool INC = false, bool EVAP = false, bool NORM = false>
__device__ __forceinline__ void dev_column_tl(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                              const NLFields<R>& g, const Streams<R, (INC ? 1 : 2) * I_NL>& in_s,
                                              Ring<R, (INC ? 1 : 2) * I_NL, BLOCK>& ring, uint32_t S, int nlev, uint32_t i,
                                              bool valid, R fac = R(0), bool ignore_supsat = false,
                                              double* norm1 = nullptr) {
  using C = Cfg<EVAP, true>;
  double n1 = 0.0;
  ring_issue(ring, in_s, i);
  const int jsel = tropopause_candidate(p, tab, f.t, f.tnd_t, int64_t(S), int64_t(i));
  const int ncand = tab.nw + 1;

  Carry<R> c{R(0), R(0), R(0)}, ci{R(0), R(0), R(0)};
  const R aph_s = f.aph[uint32_t(nlev) * S + i];
  const R aph_s_i = EVAP ? (INC ? mul_rn(fac, aph_s) : g.aph[uint32_t(nlev) * S + i]) : R(0);
  R aph0 = f.aph[i], aph0_i = INC ? mul_rn(fac, aph0) : g.aph[i];
  if (valid) {  // half level 0 (TL :757-765)
    f.fplsl[i] = R(0); f.fplsn[i] = R(0); f.fhpsl[i] = R(0); f.fhpsn[i] = R(0);
    g.fplsl[i] = R(0); g.fplsn[i] = R(0); g.fhpsl[i] = R(0); g.fhpsn[i] = R(0);
  }
  for (int k = 0; k < nlev; ++k) {
    const uint32_t off = uint32_t(k) * S + i;
    cp_async_wait_all();
    LevelIn<R> in, d;
    ring_read_level(ring, 0, aph0, in);
    if constexpr (INC) {
      // products rounded on their own (never contracted into a following FMA): bit-identical to state_increment's output
      d.ap = mul_rn(fac, in.ap);         d.aph0 = aph0_i;                     d.aph1 = mul_rn(fac, in.aph1);
      d.lu1 = mul_rn(fac, in.lu1);       d.lude = mul_rn(fac, in.lude);       d.mfd = mul_rn(fac, in.mfd);
      d.mfu = mul_rn(fac, in.mfu);       d.q = mul_rn(fac, in.q);             d.qi = mul_rn(fac, in.qi);
      d.ql = mul_rn(fac, in.ql);         d.qsat = mul_rn(fac, in.qsat);
      d.supsat = ignore_supsat ? R(0) : mul_rn(fac, in.supsat);
      d.t = mul_rn(fac, in.t);           d.tnd_q = mul_rn(fac, in.tnd_q);     d.tnd_qi = mul_rn(fac, in.tnd_qi);
      d.tnd_ql = mul_rn(fac, in.tnd_ql); d.tnd_t = mul_rn(fac, in.tnd_t);
    } else {
      ring_read_level(ring, I_NL, aph0_i, d);
    }
    if (k + 1 < nlev) ring_issue(ring, in_s, off + S);
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
    if (valid) {
      const uint32_t offn = off + S;
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
    }
    if constexpr (NORM) {
      auto sq = [](R v) { return double(v) * double(v); };
      n1 += sq(oi.tnd_t) + sq(oi.tnd_q) + sq(oi.tnd_ql) + sq(oi.tnd_qi) + sq(oi.clc) + sq(oi.covptot) + sq(ci.rfl) +
            sq(ci.sfl) + sq(-ci.rfl * p.RLVTT) + sq(-ci.sfl * p.RLSTT);
    }
    aph0 = in.aph1;
    aph0_i = d.aph1;
  }
  if constexpr (NORM) {
    if (valid) norm1[i] = n1;
  }
}

// ---------------------------------------------------------------------------------------
// AD backward.  Streams: [0,16) NL inputs with aph read at level k (not k+1), then the level-entry
// fluxes fplsl/fplsn[k] written by the forward sweep, the 5 full-level seeds at k and the 4 flux
// seeds at half level k+1.  The reference consumes (zeroes) its seeds; here each thread resets the seeds of the
// level it consumed CS2_AD_ZERO_LAG iterations earlier from inside the level loop (plain stores off the dependent
// chain) and the last few levels after the loop -- see the end of dev_column_ad_bwd for the measurements
// (1.19 ms against 1.29 ms with 10 cudaMemsetAsync after the kernel, which CS2_AD_SEED_MEMSET=1 restores for A/B;
// zeroing a slot right after its own load serialises in L2 and was 4x slower, profiles/r1c_ad_bwd.md).
// ---------------------------------------------------------------------------------------
#ifndef CS2_AD_ZERO_LAG
#define CS2_AD_ZERO_LAG 4  // the backward sweep resets the seeds of the level it consumed this many iterations ago
#endif
enum { B_FPLSL = I_NL, B_FPLSN, B_S_TT, B_S_TQ, B_S_TQL, B_S_TQI, B_S_CLC, B_S_FPLSL, B_S_FHPSL, B_S_FPLSN, B_S_FHPSN, B_N,
       B_CK = B_N, B_NCK = B_N + CK_N };

// NS = B_N (recompute) or B_NCK (checkpoint: CK_N more streams with the recorded transcendentals)
// EVAP (recompute mode only): two more streams after the NS regular ones -- the overlap carry entering each level (written
// by the forward sweep) and the out_covptot seed
template <class R, int NS, bool EVAP = false>
inline Streams<R, NS + (EVAP ? 2 : 0)> ad_streams(const NLFields<R>& f, const ADSeeds<R>& s, int64_t S, int nlev, const R* ck,
                                                   const R* cov = nullptr) {
  Streams<R, NS + (EVAP ? 2 : 0)> o;
  if constexpr (EVAP) {
    o.p[NS] = cov;
    o.p[NS + 1] = s.covptot;
  }
  if constexpr (NS > B_N) {
    for (int n = 0; n < CK_N; ++n) o.p[B_N + n] = ck + size_t(n) * size_t(nlev) * size_t(S);
  }
  const Streams<R, I_NL> a = nl_streams(f, S);
  for (int n = 0; n < I_NL; ++n) o.p[n] = a.p[n];
  o.p[I_APH1] = f.aph;  // backward sweep: the new value per level is aph[k]; aph[k+1] is carried
  o.p[B_FPLSL] = f.fplsl; o.p[B_FPLSN] = f.fplsn;
  o.p[B_S_TT] = s.tnd_t; o.p[B_S_TQ] = s.tnd_q; o.p[B_S_TQL] = s.tnd_ql; o.p[B_S_TQI] = s.tnd_qi; o.p[B_S_CLC] = s.clc;
  o.p[B_S_FPLSL] = s.fplsl + S; o.p[B_S_FHPSL] = s.fhpsl + S; o.p[B_S_FPLSN] = s.fplsn + S; o.p[B_S_FHPSN] = s.fhpsn + S;
  return o;
}

// State a backward sweep carries from level k+1 to level k (and, in the level-chunked kernel, from the warp that did
// the chunk below to the warp that does the next one).
template <class R>
struct AdCarry {
  R a_rfl, a_sfl;   // adjoint of the fluxes entering the level below
  R a_dp_below;     // a_dp of level k+1
  R a_cov, a_aph_s; // evaporation branch: adjoint of the overlap carry / of the surface pressure
  R aph1;           // aph[k+1]
  double n2;        // NORM: running second inner product
};

// Levels k_hi ... k_lo (descending) of the backward sweep of one column; the inputs of level k_hi must already have been
// issued into the ring (ring_issue at offset k_hi * S + i).
template <class R, int BLOCK, int NS, bool EVAP = false, bool NORM = false>
__device__ __forceinline__ void dev_ad_bwd_span(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                                const ADOut<R>& a, const Streams<R, NS + (EVAP ? 2 : 0)>& in_s,
                                                Ring<R, NS + (EVAP ? 2 : 0), BLOCK>& ring, int jsel, R aph_s, uint32_t S,
                                                int nlev, uint32_t i, bool valid, int k_hi, int k_lo, AdCarry<R>& cy, R fac,
                                                bool ignore_supsat, R (*keep)[BLOCK], const ADSeeds<R>* zero_seeds) {
  constexpr bool CKPT = NS > B_N;
  using C = Cfg<EVAP, true>;
  const bool ad_ref = !p.ad_tl_predicates;
  const int ncand = tab.nw + 1;
  const int t = threadIdx.x;
  for (int k = k_hi; k >= k_lo; --k) {
    const uint32_t off = uint32_t(k) * S + i;
    cp_async_wait_all();
    LevelIn<R> in;
    ring_read_level(ring, 0, R(0), in);
    in.aph0 = in.aph1;  // stream I_APH1 carries aph[k] in this sweep
    in.aph1 = cy.aph1;
    if constexpr (NORM) {
      // the inner product needs every input again after level_ad: park them in shared memory instead of keeping 16 more
      // doubles live through the level (the backward kernel has no registers to spare)
      keep[0][t] = in.t; keep[1][t] = in.q; keep[2][t] = in.ql; keep[3][t] = in.qi; keep[4][t] = in.qsat; keep[5][t] = in.ap;
      keep[6][t] = in.lude; keep[7][t] = in.mfu; keep[8][t] = in.mfd; keep[9][t] = in.tnd_t; keep[10][t] = in.tnd_q;
      keep[11][t] = in.tnd_ql; keep[12][t] = in.tnd_qi; keep[13][t] = in.aph1; keep[14][t] = in.lu1; keep[15][t] = in.supsat;
    }
    Carry<R> c;
    c.rfl = ring.v[B_FPLSL][t];
    c.sfl = ring.v[B_FPLSN][t];
    c.covptot = EVAP ? ring.v[NS][t] : R(0);  // only feeds the evaporation branch
    LevelOut<R> so;
    so.tnd_t = ring.v[B_S_TT][t];
    so.tnd_q = ring.v[B_S_TQ][t];
    so.tnd_ql = ring.v[B_S_TQL][t];
    so.tnd_qi = ring.v[B_S_TQI][t];
    so.clc = ring.v[B_S_CLC][t];
    so.covptot = EVAP ? ring.v[NS + 1][t] : R(0);
    // flux seeds at half level k+1 with the enthalpy-flux seeds folded in (AD :479-484,500-501)
    R a_rfln = cy.a_rfl + (ring.v[B_S_FPLSL][t] - ring.v[B_S_FHPSL][t] * p.RLVTT);
    R a_sfln = cy.a_sfl + (ring.v[B_S_FPLSN][t] - ring.v[B_S_FHPSN][t] * p.RLSTT);

    Trans<R, CKPT ? 2 : 0> x;
    if constexpr (CKPT) {
#pragma unroll
      for (int n = 0; n < CK_N; ++n) x.v[n] = ring.v[B_N + n][t];
    }
    if (k > k_lo) ring_issue(ring, in_s, off - S);
    if (zero_seeds && valid && k + CS2_AD_ZERO_LAG < nlev) {  // seeds of a level consumed CS2_AD_ZERO_LAG iterations ago
      const ADSeeds<R>& z = *zero_seeds;
      const uint32_t zo = off + uint32_t(CS2_AD_ZERO_LAG) * S;
      z.tnd_t[zo] = R(0); z.tnd_q[zo] = R(0); z.tnd_ql[zo] = R(0); z.tnd_qi[zo] = R(0); z.clc[zo] = R(0);
      z.covptot[zo] = R(0); z.fhpsl[zo + S] = R(0); z.fhpsn[zo + S] = R(0); z.fplsl[zo + S] = R(0); z.fplsn[zo + S] = R(0);
    }

    LevelOut<R> o;
    Traj<R> tr;
    level_fwd<R, C, true>(p, in, tab.scalm[k], tab.crh2[k * ncand + jsel], k < nlev - 1, aph_s, ad_ref, c, o, tr, x);
    LevelIn<R> ad;
    level_ad<R, C>(p, in, tr, so, ad_ref, a_rfln, a_sfln, ad, aph_s, &cy.a_cov, &cy.a_aph_s);
    cy.a_rfl = a_rfln;
    cy.a_sfl = a_sfln;

    if (valid) {
      const uint32_t offn = off + S;
      a.t[off] = ad.t;           a.tnd_t[off] = ad.tnd_t;
      a.q[off] = ad.q;           a.tnd_q[off] = ad.tnd_q;
      a.ql[off] = ad.ql;         a.tnd_ql[off] = ad.tnd_ql;
      a.qi[off] = ad.qi;         a.tnd_qi[off] = ad.tnd_qi;
      a.supsat[off] = ad.supsat; a.qsat[off] = ad.qsat;
      a.ap[off] = ad.ap;         a.lude[off] = ad.lude;
      a.mfu[off] = ad.mfu;       a.mfd[off] = ad.mfd;
      // staggered fields (AD :969-986): aph_i[k+1] = a_dp(k) - a_dp(k+1); lu_i[k+1] = adjoint of lu[k+1]
      a.aph[offn] = ad.aph1 - cy.a_dp_below;
      a.lu[offn] = ad.lu1;
    }
    if constexpr (NORM) {
      asm volatile("" ::: "memory");
      auto pr = [fac](R x, R adj) { return double(mul_rn(fac, x)) * double(adj); };
      cy.n2 += pr(keep[0][t], ad.t) + pr(keep[1][t], ad.q) + pr(keep[2][t], ad.ql) + pr(keep[3][t], ad.qi) +
               pr(keep[4][t], ad.qsat) + pr(keep[5][t], ad.ap) + pr(keep[6][t], ad.lude) + pr(keep[7][t], ad.mfu) +
               pr(keep[8][t], ad.mfd) + pr(keep[9][t], ad.tnd_t) + pr(keep[10][t], ad.tnd_q) + pr(keep[11][t], ad.tnd_ql) +
               pr(keep[12][t], ad.tnd_qi) + pr(keep[13][t], R(ad.aph1 - cy.a_dp_below)) + pr(keep[14][t], ad.lu1);
      if (!ignore_supsat) cy.n2 += pr(keep[15][t], ad.supsat);
    }
    cy.a_dp_below = ad.aph1;
    cy.aph1 = in.aph0;
  }
}

// What the thread that reaches the top of a column does after its last level (k = 0).
template <class R, bool EVAP, bool NORM>
__device__ __forceinline__ void dev_ad_bwd_finish(const ADOut<R>& a, uint32_t S, int nlev, uint32_t i, const AdCarry<R>& cy,
                                                  R fac, double* norm2, const ADSeeds<R>* zero_seeds) {
  a.aph[i] = -cy.a_dp_below;
  a.lu[i] = R(0);
  if (EVAP) a.aph[uint32_t(nlev) * S + i] += cy.a_aph_s;  // adjoint of the surface pressure (AD :974-975); own earlier store
  if constexpr (NORM) {
    norm2[i] = cy.n2 + double(mul_rn(fac, cy.aph1)) * double(R(-cy.a_dp_below));  // half level 0 (aph1 holds aph[0] here)
  }
  // The reference consumes its seeds (adjoint/_stencils/cloudsc2.py:482-484,506-542,650,714,920,972-984).  A column's
  // seeds are read by its own thread(s) only, so the sweep resets them itself: inside the level loop the seeds of the level
  // consumed CS2_AD_ZERO_LAG iterations earlier (plain stores off the dependent chain, spread over the sweep), here the
  // last few levels.  Measured at 65 536 columns: 1.29 ms with 10 cudaMemsetAsync after the kernel, 1.25 ms with one
  // trailing burst of stores, 1.19 ms with the lagged in-loop stores (any lag from 1 to 16); zeroing a slot right after
  // its own load, as the very first version did, serialises in L2 and was 4x slower (profiles/r1c_ad_bwd.md).
  if (zero_seeds) {
    const ADSeeds<R>& z = *zero_seeds;
    const int ktail = nlev < CS2_AD_ZERO_LAG ? nlev : CS2_AD_ZERO_LAG;  // the levels the in-loop reset has not reached
    for (int k = 0; k < ktail; ++k) {
      const uint32_t off = uint32_t(k) * S + i;
      z.tnd_t[off] = R(0); z.tnd_q[off] = R(0); z.tnd_ql[off] = R(0); z.tnd_qi[off] = R(0); z.clc[off] = R(0);
      z.covptot[off] = R(0);
    }
    for (int k = 0; k <= ktail; ++k) {  // half-level seeds: one more level
      const uint32_t off = uint32_t(k) * S + i;
      z.fhpsl[off] = R(0); z.fhpsn[off] = R(0); z.fplsl[off] = R(0); z.fplsn[off] = R(0);
    }
  }
}

// NORM: also the symmetry test's second inner product, norm2[i] = SUM_k SUM_fields (fac * input) * (adjoint output) of the
// column (adjoint/validation.py:183-215; the increments are StateIncrement's: products rounded on their own, supsat_i = 0
// with ignore_supsat), accumulated in fp64 from the values as stored.
template <class R, int BLOCK, int NS, bool EVAP = false, bool NORM = false>
__device__ __forceinline__ void dev_column_ad_bwd(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                                  const ADOut<R>& a, const Streams<R, NS + (EVAP ? 2 : 0)>& in_s,
                                                  Ring<R, NS + (EVAP ? 2 : 0), BLOCK>& ring, const int32_t* jsel_in,
                                                  uint32_t S, int nlev, uint32_t i, bool valid, R fac = R(0),
                                                  bool ignore_supsat = false, double* norm2 = nullptr,
                                                  R (*keep)[BLOCK] = nullptr, const ADSeeds<R>* zero_seeds = nullptr) {
  ring_issue(ring, in_s, uint32_t(nlev - 1) * S + i);
  const int jsel = jsel_in[i];
  const R aph_s = f.aph[uint32_t(nlev) * S + i];
  AdCarry<R> cy{R(0), R(0), R(0), R(0), R(0), aph_s, 0.0};
  dev_ad_bwd_span<R, BLOCK, NS, EVAP, NORM>(p, tab, f, a, in_s, ring, jsel, aph_s, S, nlev, i, valid, nlev - 1, 0, cy, fac,
                                            ignore_supsat, keep, zero_seeds);
  if (valid) dev_ad_bwd_finish<R, EVAP, NORM>(a, S, nlev, i, cy, fac, norm2, zero_seeds);
}

}  // namespace cs2
