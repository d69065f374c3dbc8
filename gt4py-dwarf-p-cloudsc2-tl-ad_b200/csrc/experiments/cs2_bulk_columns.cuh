// NL column sweep fed by the TMA engine: per level ONE elected thread issues 16 bulk copies
// (cp.async.bulk, SASS UBLKCP), one per input field, each moving the CTA's contiguous 64-column segment of
// that level (512 B in fp64) into shared memory and signalling an mbarrier with the byte count.  Compared with
// the per-thread cp.async version (cs2_device_columns.cuh) this removes 16 LDGSTS + 16 address IMADs per thread
// and level from the instruction stream at the price of coupling the CTA's two warps through one
// __syncthreads per level (the buffer a bulk copy overwrites must have been read by every thread).
#pragma once

#include "../cs2_device_columns.cuh"

namespace cs2 {

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}

template <class R, int N, int BLOCK>
struct BulkRing {
  alignas(128) R v[2][N][BLOCK];
  alignas(8) uint64_t bar[2];
};

template <class R, int N, int BLOCK>
__device__ __forceinline__ void bulk_issue(BulkRing<R, N, BLOCK>& ring, const Streams<R, N>& in, int stage, uint32_t off0,
                                           unsigned seg_bytes) {
  mbar_expect_tx(&ring.bar[stage], unsigned(N) * seg_bytes);
#pragma unroll
  for (int f = 0; f < N; ++f) bulk_copy_g2s(&ring.v[stage][f][0], in.p[f] + off0, seg_bytes, &ring.bar[stage]);
}

template <class R, class C, int BLOCK>
__device__ __forceinline__ void dev_column_nl_bulk(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                                   const Streams<R, I_NL>& in_s, BulkRing<R, I_NL, BLOCK>& ring,
                                                   uint32_t S, int nlev, uint32_t i0, uint32_t ncol) {
  const int tid = threadIdx.x;
  const uint32_t gi = i0 + tid;
  const bool valid = gi < ncol;
  const uint32_t i = valid ? gi : ncol - 1;           // out-of-range threads shadow the last column
  const int slot = int(i - i0);
  const unsigned seg_bytes = unsigned(min(uint32_t(BLOCK), S - i0)) * sizeof(R);  // stays inside the level row
  if (tid == 0) {
    mbar_init(&ring.bar[0], 1);
    mbar_init(&ring.bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) bulk_issue(ring, in_s, 0, i0, seg_bytes);

  const int jsel = tropopause_candidate(p, tab, f.t, f.tnd_t, int64_t(S), int64_t(i));
  const int ncand = tab.nw + 1;
  Carry<R> c{R(0), R(0), R(0)};
  const R aph_s = f.aph[uint32_t(nlev) * S + i];
  R aph0 = f.aph[i];
  if (valid) {
    f.fhpsl[i] = R(0);
    f.fhpsn[i] = R(0);
  }
  for (int k = 0; k < nlev; ++k) {
    const int st = k & 1;
    const uint32_t off = uint32_t(k) * S + i;
    mbar_wait(&ring.bar[st], unsigned(k >> 1) & 1u);
    LevelIn<R> in;
    in.ap = ring.v[st][I_AP][slot];         in.aph0 = aph0;                        in.aph1 = ring.v[st][I_APH1][slot];
    in.lu1 = ring.v[st][I_LU1][slot];       in.lude = ring.v[st][I_LUDE][slot];    in.mfd = ring.v[st][I_MFD][slot];
    in.mfu = ring.v[st][I_MFU][slot];       in.q = ring.v[st][I_Q][slot];          in.qi = ring.v[st][I_QI][slot];
    in.ql = ring.v[st][I_QL][slot];         in.qsat = ring.v[st][I_QSAT][slot];    in.supsat = ring.v[st][I_SUPSAT][slot];
    in.t = ring.v[st][I_T][slot];           in.tnd_q = ring.v[st][I_TQ][slot];     in.tnd_qi = ring.v[st][I_TQI][slot];
    in.tnd_ql = ring.v[st][I_TQL][slot];    in.tnd_t = ring.v[st][I_TT][slot];
    // every thread has now read stage st^1 (level k-1) and stage st (level k): stage st^1 may be refilled
    if (BLOCK == 32)
      __syncwarp();
    else
      __syncthreads();
    if (tid == 0 && k + 1 < nlev) bulk_issue(ring, in_s, st ^ 1, uint32_t(k + 1) * S + i0, seg_bytes);
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

}  // namespace cs2
