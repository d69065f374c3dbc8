// Level-chunked AD backward sweep with persistent warps -- EXPERIMENT, measured and not adopted (profiles/r2n_ad_chunked.md).
#pragma once

#include "../cs2_device_columns.cuh"

namespace cs2 {

// ---------------------------------------------------------------------------------------
// AD backward, level-chunked with persistent warps (default flags, recompute mode).  The backward sweep needs 240 registers:
// 8 warps per SM, 1 184 warp slots on the chip.  65 536 columns are 2 048 warps = 1.73 rounds of whole-column sweeps, and the
// second round, 73 % full, costs almost a full round (profiles/r2e).  Here the work item is (32 columns, a chunk of levels): a
// warp takes items from a ticket counter in the order "bottom chunk of every column group, then the next chunk up, ...", picks
// up the three adjoint carries the warp that did the chunk below left in global memory, sweeps its levels and hands over in
// turn.  An item's predecessor is 2 048 tickets older while only 1 184 warps run, so it has finished long before and nobody
// waits; the chip stays full until the last ticket.
// Hand-over without flags or fences: the carry slots of every chunk boundary are pre-set to an all-ones NaN pattern by the
// launcher (cudaMemsetAsync 0xFF) and every lane simply re-reads its own three slots until none holds the pattern (one round
// trip, issued together with the first level's input loads); arithmetic never produces that pattern (results are canonical
// NaNs), and a bounded spin traps instead of hanging if it ever did.  The ticket of the NEXT item is taken at the start of
// the current one, so its latency is hidden as well.
//   ticket = one unsigned counter;  carry = [nchunk - 1][3][S] (a_rfl, a_sfl, a_dp_below per chunk boundary)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_unset(double v) { return __double_as_longlong(v) == -1LL; }
__device__ __forceinline__ bool is_unset(float v) { return __float_as_int(v) == -1; }

template <class R>
__device__ __forceinline__ R ld_handover(const R* p) {
  R v = __ldcv(p);
  for (unsigned spin = 0; is_unset(v); ++spin) {
    if (spin > (1u << 22)) __trap();  // ~1 s: the predecessor chunk never arrived (cannot happen with in-order tickets)
    __nanosleep(100);
    v = __ldcv(p);
  }
  return v;
}

template <class R, int BLOCK>
__device__ __forceinline__ void dev_ad_bwd_chunked(const DevParams<R>& p, const LevelTables<R>& tab, const NLFields<R>& f,
                                                   const ADOut<R>& a, const Streams<R, B_N>& in_s, Ring<R, B_N, BLOCK>& ring,
                                                   const int32_t* jsel_in, uint32_t S, int nlev, uint32_t ncol, int nchunk,
                                                   int chunk_levels, unsigned* ticket, R* carry, const ADSeeds<R>* zero_seeds) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned ngroups = (ncol + 31u) / 32u;
  const unsigned nitems = ngroups * unsigned(nchunk);
  unsigned j = 0;
  if (lane == 0) j = atomicAdd(ticket, 1u);
  j = __shfl_sync(0xffffffffu, j, 0);
  while (j < nitems) {
    unsigned jnext = 0;
    if (lane == 0) jnext = atomicAdd(ticket, 1u);  // used after this item: the round trip overlaps the sweep
    const unsigned c = j / ngroups, g = j - c * ngroups;
    const uint32_t gi = g * 32u + lane;
    const bool valid = gi < ncol;
    const uint32_t i = valid ? gi : ncol - 1;
    const int k_hi = nlev - 1 - int(c) * chunk_levels;
    const int k_lo = (k_hi - chunk_levels + 1 > 0) ? k_hi - chunk_levels + 1 : 0;
    ring_issue(ring, in_s, uint32_t(k_hi) * S + i);  // inputs of the first level in flight while the hand-over is read
    const int jsel = jsel_in[i];
    const R aph_s = f.aph[uint32_t(nlev) * S + i];
    AdCarry<R> cy{R(0), R(0), R(0), R(0), R(0), aph_s, 0.0};
    if (c > 0) {
      const R* slot = carry + size_t(c - 1) * 3 * size_t(S) + gi;
      cy.aph1 = f.aph[uint32_t(k_hi + 1) * S + i];
      cy.a_rfl = ld_handover(slot);
      cy.a_sfl = ld_handover(slot + S);
      cy.a_dp_below = ld_handover(slot + 2 * size_t(S));
    }
    dev_ad_bwd_span<R, BLOCK, B_N, false, false>(p, tab, f, a, in_s, ring, jsel, aph_s, S, nlev, i, valid, k_hi, k_lo, cy,
                                                 R(0), false, nullptr, zero_seeds);
    if (k_lo > 0) {
      R* slot = carry + size_t(c) * 3 * size_t(S) + gi;
      __stcg(slot, cy.a_rfl);
      __stcg(slot + S, cy.a_sfl);
      __stcg(slot + 2 * size_t(S), cy.a_dp_below);
    } else if (valid) {
      dev_ad_bwd_finish<R, false, false>(a, S, nlev, i, cy, R(0), nullptr, zero_seeds);
    }
    j = __shfl_sync(0xffffffffu, jnext, 0);
  }
}

}  // namespace cs2
