// NL column sweep with the level split between two warps (level physics: cs2_physics_split.cuh).
//
// A CTA owns COLS columns and has 2 * COLS threads: warps [0, COLS/32) are "A" warps, warps [COLS/32, 2*COLS/32)
// are "B" warps; A-warp w and B-warp w own the same 32 columns, lane for lane.  Per level the A thread of a column
// loads the 16 inputs (cp.async prefetch as in cs2_device_columns.cuh), evaluates level_nl_a and writes the
// hand-over values into a two-stage shared-memory buffer; the B thread evaluates level_nl_b with the carried
// fluxes and stores the tendencies and fluxes.  The two warps of a pair are coupled by four mbarriers (full /
// empty per stage, arrival count 1: lane 0 arrives after __syncwarp), nothing is CTA-wide.  A therefore runs up
// to two levels ahead of B, and the dependent FP64 chain a warp walks per level is about half as long as in the
// one-thread-per-column kernel, with twice as many warps resident.
#pragma once

#include "cs2_bulk_columns.cuh"  // mbarrier helpers
#include "cs2_physics_split.cuh"

namespace cs2 {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}

constexpr int kSplitMaxLev = 192;  // the launcher falls back to the one-thread-per-column kernel above this

template <class R, int NM, int COLS>
struct SplitShared {
  R in[I_NL + 1][COLS];  // cp.async slots of the A threads: the 16 level inputs + crh2[k][candidate of the column]
  R mid[2][NM][COLS];
  R scalm[kSplitMaxLev];
  alignas(8) uint64_t full[COLS / 32][2];
  alignas(8) uint64_t empty[COLS / 32][2];
};

template <class R, int COLS>
struct MidView {
  const R (*v)[COLS];
  int slot;
  __device__ __forceinline__ R operator[](int n) const { return v[n][slot]; }
};

template <class R, class C, int NM, int COLS>
__device__ __forceinline__ void dev_column_nl_split(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                                    const Streams<R, I_NL>& in_s, SplitShared<R, NM, COLS>& sh, uint32_t S,
                                                    int nlev, uint32_t ncol) {
  const int tid = threadIdx.x;
  const bool is_a = tid < COLS;
  const int slot = is_a ? tid : tid - COLS;
  const int pair = slot >> 5, lane = slot & 31;
  const uint32_t gi = blockIdx.x * uint32_t(COLS) + uint32_t(slot);
  const bool valid = gi < ncol;
  const uint32_t i = valid ? gi : ncol - 1;  // out-of-range threads shadow the last column and store nothing
  if (is_a && lane == 0) {
    mbar_init(&sh.full[pair][0], 1);
    mbar_init(&sh.full[pair][1], 1);
    mbar_init(&sh.empty[pair][0], 1);
    mbar_init(&sh.empty[pair][1], 1);
    mbar_fence_init();
  }
  if (!is_a)
    for (int k = slot; k < nlev; k += COLS) sh.scalm[k] = tab.scalm[k];
  __syncthreads();

  if (is_a) {
    // ---- A: carry-independent half; no global stores, no carried state except aph[k]
#pragma unroll
    for (int n = 0; n < I_NL; ++n) cp_async<sizeof(R)>(&sh.in[n][slot], in_s.p[n] + i);
    const int ncand = tab.nw + 1;
    const R* crh2p = tab.crh2 + tropopause_candidate(p, tab, f.t, f.tnd_t, int64_t(S), int64_t(i));
    cp_async<sizeof(R)>(&sh.in[I_NL][slot], crh2p);
    cp_async_commit();
    R aph0 = f.aph[i];
    uint32_t off = i;
    for (int k = 0; k < nlev; ++k) {
      cp_async_wait_all();
      LevelIn<R> in;
      in.ap = sh.in[I_AP][slot];
      in.aph0 = aph0;
      in.aph1 = sh.in[I_APH1][slot];
      in.lu1 = sh.in[I_LU1][slot];
      in.lude = sh.in[I_LUDE][slot];
      in.mfd = sh.in[I_MFD][slot];
      in.mfu = sh.in[I_MFU][slot];
      in.q = sh.in[I_Q][slot];
      in.qi = sh.in[I_QI][slot];
      in.ql = sh.in[I_QL][slot];
      in.qsat = sh.in[I_QSAT][slot];
      in.supsat = sh.in[I_SUPSAT][slot];
      in.t = sh.in[I_T][slot];
      in.tnd_q = sh.in[I_TQ][slot];
      in.tnd_qi = sh.in[I_TQI][slot];
      in.tnd_ql = sh.in[I_TQL][slot];
      in.tnd_t = sh.in[I_TT][slot];
      const R crh2 = sh.in[I_NL][slot];
      if (k + 1 < nlev) {
        off += S;
        crh2p += ncand;
#pragma unroll
        for (int n = 0; n < I_NL; ++n) cp_async<sizeof(R)>(&sh.in[n][slot], in_s.p[n] + off);
        cp_async<sizeof(R)>(&sh.in[I_NL][slot], crh2p);
        cp_async_commit();
      }
      Mid<R> m;
      level_nl_a<R, C>(p, in, sh.scalm[k], crh2, k < nlev - 1, m);
      const int s = k & 1;
      mbar_wait(&sh.empty[pair][s], unsigned(((k >> 1) & 1) ^ 1));
#pragma unroll
      for (int n = 0; n < NM; ++n) sh.mid[s][n][slot] = m.v[n];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.full[pair][s]);
      aph0 = in.aph1;
    }
  } else {
    // ---- B: carries the precipitation fluxes down the column and owns every store
    Carry<R> c{R(0), R(0), R(0)};
    if (valid) {
      f.fhpsl[i] = R(0);
      f.fhpsn[i] = R(0);
    }
    for (int k = 0; k < nlev; ++k) {
      const uint32_t off = uint32_t(k) * S + i;
      const int s = k & 1;
      mbar_wait(&sh.full[pair][s], unsigned((k >> 1) & 1));
      const MidView<R, COLS> m{sh.mid[s], slot};
      LevelOut<R> o;
      level_nl_b<R>(p, m, c, o);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.empty[pair][s]);
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
    }
  }
}

}  // namespace cs2
