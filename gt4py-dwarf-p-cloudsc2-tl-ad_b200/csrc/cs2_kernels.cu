// libcloudsc2_b200.so -- CUDA kernels (sm_100a) and the C ABI declared in include/cloudsc2_b200.h.
//
// Kernel shapes (see DESIGN.md for the roofline of each):
//   * pointwise kernels (saturation, state_increment, perturbed_state): one element per thread
//     per field, grid (ceil(ncol/256), levels): pure HBM streaming.
//   * column kernels (NL, TL, AD forward, AD backward): one thread per column, 128-thread CTAs,
//     sequential walk over the 137 levels with the flux / overlap carries in registers; a warp
//     reads/writes 32 consecutive columns of one level per request (256 B in fp64).
//   * reductions (Taylor sums, symmetry inner products): fp64 accumulation, warp shuffles,
//     deterministic two-stage block reduction.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "cs2_device_columns.cuh"
#ifdef CS2_EXPERIMENTS  // measured-and-rejected kernel variants (profiles/README.md); not in the shipped library
#include "experiments/cs2_ad_chunked.cuh"
#include "experiments/cs2_bulk_columns.cuh"
#include "experiments/cs2_pipe_columns.cuh"
#include "experiments/cs2_split_columns.cuh"
#endif

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return CS2_OK;
  return fail(CS2_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

int check_dims(const cs2_dims* d) {
  if (!d) return fail(CS2_ERR_NULL_POINTER, "dims is NULL");
  if (d->dtype != CS2_F64 && d->dtype != CS2_F32) return fail(CS2_ERR_BAD_DIMS, "dtype must be CS2_F64 or CS2_F32");
  if (d->ncol < 0 || d->nlev < 1 || d->nlev > 4096) return fail(CS2_ERR_BAD_DIMS, "ncol < 0 or nlev outside [1, 4096]");
  if (d->ncol_stride < d->ncol) return fail(CS2_ERR_BAD_DIMS, "ncol_stride < ncol");
  if (d->ncol_stride % 32 != 0) return fail(CS2_ERR_MISALIGNED, "ncol_stride must be a multiple of 32 elements");
  if (uint64_t(d->ncol_stride) * uint64_t(d->nlev + 1) >= (uint64_t(1) << 32))
    return fail(CS2_ERR_BAD_DIMS, "ncol_stride * (nlev + 1) must be < 2^32 elements per call (kernels use 32-bit element offsets)");
  return CS2_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_ptrs(const void* const* ptrs, int n, const char* what) {
  for (int i = 0; i < n; ++i) {
    if (!ptrs[i]) return fail(CS2_ERR_NULL_POINTER, std::string(what) + ": field pointer " + std::to_string(i) + " is NULL");
    if (!aligned16(ptrs[i]))
      return fail(CS2_ERR_MISALIGNED, std::string(what) + ": field pointer " + std::to_string(i) + " is not 16-byte aligned");
  }
  return CS2_OK;
}

int check_nl_fields(const cs2_nl_fields* f, const char* what) {
  if (!f) return fail(CS2_ERR_NULL_POINTER, std::string(what) + " is NULL");
  static_assert(sizeof(cs2_nl_fields) == 26 * sizeof(void*), "cs2_nl_fields layout");
  return check_ptrs(reinterpret_cast<const void* const*>(f), 26, what);
}

constexpr int kColumnBlock = 64;  // 65 536 columns = 1024 CTAs: one wave at 7 CTAs/SM (<= 144 registers)
constexpr int kPointBlock = 256;
#ifndef CS2_WIDE_BLOCK
#define CS2_WIDE_BLOCK 64
#endif
constexpr int kWideBlock = CS2_WIDE_BLOCK;  // CTA size of the register-hungry TL / AD-backward kernels

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
// Four consecutive columns of one level per thread: four independent FP64 chains (the kernel is bound by the FP64 pipe
// and its latency: 1-2 exponentials and 3 reciprocals per 24 bytes), two 16-byte loads per input in flight.
template <class R>
__global__ void __launch_bounds__(kPointBlock)
saturation_kernel(const __grid_constant__ cs2::DevParams<R> p, int lphylin, const R* __restrict__ ap,
                  const R* __restrict__ t, R* __restrict__ qsat, int64_t ncol, int64_t S) {
  const int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= ncol) return;
  const int64_t off = int64_t(blockIdx.y) * S + i;
  if (i + 3 < ncol) {
    using V = typename cs2::Vec2<R>::type;
    const V a0 = *reinterpret_cast<const V*>(ap + off), a1 = *reinterpret_cast<const V*>(ap + off + 2);
    const V t0 = *reinterpret_cast<const V*>(t + off), t1 = *reinterpret_cast<const V*>(t + off + 2);
    const R av[4] = {a0.x, a0.y, a1.x, a1.y}, tv[4] = {t0.x, t0.y, t1.x, t1.y};
    R q[4];
    cs2::saturation_points<R, 4>(p, lphylin != 0, av, tv, q);
    V q0, q1;
    q0.x = q[0]; q0.y = q[1]; q1.x = q[2]; q1.y = q[3];
    *reinterpret_cast<V*>(qsat + off) = q0;
    *reinterpret_cast<V*>(qsat + off + 2) = q1;
  } else {
    for (int64_t j = 0; j < 4 && i + j < ncol; ++j)
      qsat[off + j] = cs2::saturation_point<R>(p, lphylin != 0, ap[off + j], t[off + j]);
  }
}

template <class R>
struct StatePtrs {
  const R* in[CS2_NSTATE];
  const R* in_i[CS2_NSTATE];
  R* out[CS2_NSTATE];
};

template <class R>
__global__ void __launch_bounds__(kPointBlock)
state_increment_kernel(const __grid_constant__ StatePtrs<R> f, R fac, int ignore_supsat, int64_t ncol, int64_t S) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= ncol) return;
  const int64_t off = int64_t(blockIdx.y) * S + i;
  R v[CS2_NSTATE];
#pragma unroll
  for (int n = 0; n < CS2_NSTATE; ++n) v[n] = f.in[n][off];
#pragma unroll
  for (int n = 0; n < CS2_NSTATE; ++n) f.out[n][off] = fac * v[n];
  if (ignore_supsat) f.out[CS2_NSTATE - 1][off] = R(0);
}

template <class R>
__global__ void __launch_bounds__(kPointBlock)
perturbed_state_kernel(const __grid_constant__ StatePtrs<R> f, R fac, int64_t ncol, int64_t S) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= ncol) return;
  const int64_t off = int64_t(blockIdx.y) * S + i;
  R v[CS2_NSTATE], w[CS2_NSTATE];
#pragma unroll
  for (int n = 0; n < CS2_NSTATE; ++n) {
    v[n] = f.in[n][off];
    w[n] = f.in_i[n][off];
  }
#pragma unroll
  for (int n = 0; n < CS2_NSTATE; ++n) f.out[n][off] = v[n] + fac * w[n];
}

// CKPT: record the transcendentals (AD forward, checkpoint mode); LIN: AD forward sweep (trajectory evaluated in the
// same form as the TL / AD-backward kernels evaluate it)
#ifndef CS2_NLP_MAXNREG
#define CS2_NLP_MAXNREG 128  // the fused perturbed-NL / Taylor-factor sweeps
#endif
#ifndef CS2_NL_MAXNREG
#define CS2_NL_MAXNREG 96  // spill-free in every instantiation; 20 warps per SM when the grid has them (1 M columns: -5 %)
#endif
template <class R, class C, bool CKPT, bool LIN>
__global__ void __maxnreg__(CS2_NL_MAXNREG)
nl_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
          const __grid_constant__ cs2::NLFields<R> f, const __grid_constant__ cs2::Streams<R, cs2::I_NL> in_s,
          int64_t ncol, int64_t S, int nlev, int ad_ref, int32_t* jsel_out, R* ck, R* cov_out) {
  __shared__ cs2::Ring<R, cs2::I_NL, kColumnBlock> ring;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < ncol;
  if (!valid) i = ncol - 1;  // out-of-range threads shadow the last column and store nothing
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::dev_column_nl<R, C, kColumnBlock, CKPT, LIN>(p, tab, f, in_s, ring, uint32_t(S), nlev, uint32_t(i), valid, ad_ref != 0,
                                               jsel_out, ck, cov_out);
}

#ifdef CS2_EXPERIMENTS
// default-flag NL (and the AD forward sweep in recompute mode), software-pipelined across two levels
template <class R>
#ifdef CS2_PIPE_MAXNREG
__global__ void __maxnreg__(CS2_PIPE_MAXNREG)
#else
__global__ void __launch_bounds__(kColumnBlock, 7)
#endif
nl_pipe_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
               const __grid_constant__ cs2::NLFields<R> f, const __grid_constant__ cs2::Streams<R, cs2::I_NL> in_s,
               int64_t ncol, int64_t S, int nlev, int ad_ref, int32_t* jsel_out) {
  __shared__ cs2::Ring<R, cs2::I_NL, kColumnBlock> ring;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < ncol;
  if (!valid) i = ncol - 1;
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::dev_column_nl_pipe<R, kColumnBlock>(p, tab, f, in_s, ring, uint32_t(S), nlev, uint32_t(i), valid, ad_ref != 0, jsel_out);
}

#ifndef CS2_BULK_BLOCK
#define CS2_BULK_BLOCK 64
#endif
constexpr int kBulkBlock = CS2_BULK_BLOCK;
template <class R, class C>
__global__ void __launch_bounds__(kBulkBlock, 448 / kBulkBlock)
nl_bulk_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
               const __grid_constant__ cs2::NLFields<R> f, const __grid_constant__ cs2::Streams<R, cs2::I_NL> in_s,
               int64_t ncol, int64_t S, int nlev) {
  __shared__ cs2::BulkRing<R, cs2::I_NL, kBulkBlock> ring;
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::dev_column_nl_bulk<R, C, kBulkBlock>(p, tab, f, in_s, ring, uint32_t(S), nlev, uint32_t(blockIdx.x) * kBulkBlock,
                                            uint32_t(ncol));
}

#ifndef CS2_SPLIT_CTAS
#define CS2_SPLIT_CTAS 7
#endif
constexpr int kSplitCols = 64;  // columns per CTA of the split NL kernel (2 * kSplitCols threads)
template <class R, class C, int NM>
__global__ void __launch_bounds__(2 * kSplitCols, CS2_SPLIT_CTAS)
nl_split_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
                const __grid_constant__ cs2::NLFields<R> f, const __grid_constant__ cs2::Streams<R, cs2::I_NL> in_s,
                int64_t ncol, int64_t S, int nlev) {
  __shared__ cs2::SplitShared<R, NM, kSplitCols> sh;
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::dev_column_nl_split<R, C, NM, kSplitCols>(p, tab, f, in_s, sh, uint32_t(S), nlev, uint32_t(ncol));
}

#endif  // CS2_EXPERIMENTS

template <class R, class C>
__global__ void __maxnreg__(CS2_NLP_MAXNREG)
nlp_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
           const __grid_constant__ cs2::NLFields<R> f, const __grid_constant__ cs2::NLFields<R> g, R fac,
           const __grid_constant__ cs2::Streams<R, 2 * cs2::I_NL> in_s, int64_t ncol, int64_t S, int nlev) {
  __shared__ cs2::Ring<R, 2 * cs2::I_NL, kColumnBlock> ring;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < ncol;
  if (!valid) i = ncol - 1;
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::dev_column_nl_pert<R, C, kColumnBlock>(p, tab, f, g, fac, in_s, ring, uint32_t(S), nlev, uint32_t(i), valid);
}

__device__ __forceinline__ double warp_sum(double v);

// One Taylor-test factor in one sweep: NL of x + f2 * (f1 * x) and SUM(F_p - F_nl) per output field (cs2_taylor_nl_sums).
// Per-CTA partial sums go to `partial[field][blockIdx.x]`; taylor_final_kernel adds them up in a fixed order.
template <class R, class C>
__global__ void __maxnreg__(CS2_NLP_MAXNREG)
taylor_nl_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
                 const __grid_constant__ cs2::NLFields<R> f, R f1, R f2, int ignore_supsat,
                 const __grid_constant__ cs2::Streams<R, cs2::I_NL + cs2::T_N> in_s, int64_t ncol, int64_t S, int nlev,
                 double2* __restrict__ partial) {
  __shared__ cs2::Ring<R, cs2::I_NL + cs2::T_N, kColumnBlock> ring;
  __shared__ double acc[cs2::T_N][kColumnBlock];
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < ncol;
  if (!valid) i = ncol - 1;
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::dev_column_nl_taylor<R, C, kColumnBlock>(p, tab, f, f1, f2, ignore_supsat != 0, in_s, ring, acc, uint32_t(S), nlev,
                                                uint32_t(i), valid);
  __syncthreads();
  // fixed-order reduction over the CTA's columns: warp w sums field w, w + 2, ... (kColumnBlock = 64: two warps)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int n = w; n < cs2::T_N; n += kColumnBlock / 32) {
    double v = 0.0;
    for (int c0 = lane; c0 < kColumnBlock; c0 += 32) v += acc[n][c0];
    v = warp_sum(v);
    if (lane == 0) partial[int64_t(n) * gridDim.x + blockIdx.x] = make_double2(v, 0.0);
  }
}

// register caps of the two register-hungry kernels (measured on B200, profiles/README.md): spilling costs far
// more than the occupancy it buys, the caps below are the largest spill-free values that changed the
// compiler's allocation for the better
#ifndef CS2_TL_MAXNREG
#define CS2_TL_MAXNREG 128
#endif
#ifndef CS2_TL_EVAP_MAXNREG
#define CS2_TL_EVAP_MAXNREG 168  // the fp64 evaporation branch (non-default flags) does not fit 128 registers without spills
#endif
#ifndef CS2_AD_MAXNREG_F32
#define CS2_AD_MAXNREG_F32 128  // fp32 values take one register: the backward sweep fits 128 = 16 warps per SM, every column resident
#endif
#ifndef CS2_AD_MAXNREG
#define CS2_AD_MAXNREG 240
#endif
template <class R, bool EVAP>
__global__ void __maxnreg__((EVAP && sizeof(R) == 8) ? CS2_TL_EVAP_MAXNREG : CS2_TL_MAXNREG)
tl_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
          const __grid_constant__ cs2::NLFields<R> f, const __grid_constant__ cs2::NLFields<R> g,
          const __grid_constant__ cs2::Streams<R, 2 * cs2::I_NL> in_s, int64_t ncol, int64_t S, int nlev) {
  __shared__ cs2::Ring<R, 2 * cs2::I_NL, kWideBlock> ring;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < ncol;
  if (!valid) i = ncol - 1;
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::dev_column_tl<R, kWideBlock, false, EVAP>(p, tab, f, g, in_s, ring, uint32_t(S), nlev, uint32_t(i), valid);
}

// fused state_increment + TL (cs2_tl_increment)
template <class R, bool EVAP, bool NORM>
__global__ void __maxnreg__((EVAP && sizeof(R) == 8) ? CS2_TL_EVAP_MAXNREG : CS2_TL_MAXNREG)
tl_inc_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
              const __grid_constant__ cs2::NLFields<R> f, const __grid_constant__ cs2::NLFields<R> g,
              const __grid_constant__ cs2::Streams<R, cs2::I_NL> in_s, int64_t ncol, int64_t S, int nlev, R fac,
              int ignore_supsat, double* __restrict__ norm1) {
  __shared__ cs2::Ring<R, cs2::I_NL, kWideBlock> ring;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < ncol;
  if (!valid) i = ncol - 1;
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::dev_column_tl<R, kWideBlock, true, EVAP, NORM>(p, tab, f, g, in_s, ring, uint32_t(S), nlev, uint32_t(i), valid, fac,
                                                      ignore_supsat != 0, norm1);
}

template <class R, int NS, bool EVAP, bool NORM = false>
__global__ void __maxnreg__(sizeof(R) == 4 ? CS2_AD_MAXNREG_F32 : (EVAP ? 255 : CS2_AD_MAXNREG))
ad_bwd_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
              const __grid_constant__ cs2::NLFields<R> f, const __grid_constant__ cs2::ADOut<R> a,
              const __grid_constant__ cs2::Streams<R, NS + (EVAP ? 2 : 0)> in_s, const int32_t* __restrict__ jsel,
              int64_t ncol, int64_t S, int nlev, R fac, int ignore_supsat, double* __restrict__ norm2,
              const __grid_constant__ cs2::ADSeeds<R> seeds, int zero_seeds) {
  __shared__ cs2::Ring<R, NS + (EVAP ? 2 : 0), kWideBlock> ring;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < ncol;
  if (!valid) i = ncol - 1;
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  __shared__ R keep[NORM ? 16 : 1][kWideBlock];  // inputs of the current level, for the fused inner product
  cs2::dev_column_ad_bwd<R, kWideBlock, NS, EVAP, NORM>(p, tab, f, a, in_s, ring, jsel, uint32_t(S), nlev, uint32_t(i), valid, fac,
                                                        ignore_supsat != 0, norm2, keep, zero_seeds ? &seeds : nullptr);
}

#ifdef CS2_EXPERIMENTS
// level-chunked backward sweep with persistent warps (default flags, recompute mode): see dev_ad_bwd_chunked
template <class R>
__global__ void __maxnreg__(CS2_AD_MAXNREG)
ad_bwd_chunked_kernel(const __grid_constant__ cs2::DevParams<R> p, const void* __restrict__ tables,
                      const __grid_constant__ cs2::NLFields<R> f, const __grid_constant__ cs2::ADOut<R> a,
                      const __grid_constant__ cs2::Streams<R, cs2::B_N> in_s, const int32_t* __restrict__ jsel, int64_t ncol,
                      int64_t S, int nlev, const __grid_constant__ cs2::ADSeeds<R> seeds, int zero_seeds, int nchunk,
                      int chunk_levels, unsigned* ticket, R* carry) {
  __shared__ cs2::Ring<R, cs2::B_N, kWideBlock> ring;
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::dev_ad_bwd_chunked<R, kWideBlock>(p, tab, f, a, in_s, ring, jsel, uint32_t(S), nlev, uint32_t(ncol), nchunk, chunk_levels,
                                         ticket, carry, zero_seeds ? &seeds : nullptr);
}

#endif  // CS2_EXPERIMENTS

// ---- FP64 pipe micro-benchmark (the roofline's second axis: MEASURED_PEAKS.json has no FP64 entry) -----------
__global__ void __launch_bounds__(256) dfma_rate_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[size_t(blockIdx.x) * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// Same with all three operands in REGISTERS that differ per chain (what the column kernels' DFMAs look like: products and
// sums of per-column values), and with the mix of the physics: a dependent pair DMUL -> DFMA per chain.  MODE 1: x = fma(x, y, z);
// MODE 2: x = fma(x * y, z, w).
template <int MODE>
__global__ void __launch_bounds__(256) dfma_rate_regs_kernel(double* out, int iters, const double* __restrict__ seed) {
  const int t = threadIdx.x;
  double x[8], y[8], z[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    x[n] = seed[(t + n) & 255];
    y[n] = 0.999999 + 1e-9 * seed[(t + 3 * n + 1) & 255];
    z[n] = 1e-6 * seed[(t + 5 * n + 2) & 255];
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (MODE == 1) x[n] = fma(x[n], y[n], z[n]);
      else x[n] = fma(x[n] * y[n], y[(n + 1) & 7], z[n]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int n = 0; n < 8; ++n) s += x[n];
  out[size_t(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

// ---- reductions -----------------------------------------------------------------------
constexpr int kMaxRedFields = 32;
struct RedPtrs {
  const void* a[kMaxRedFields];
  const void* b[kMaxRedFields];
  const void* c[kMaxRedFields];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// grid (nblk, nfields): partial[f][blk] = { sum(a - b), sum(c) } over this block's elements
template <class R>
__global__ void __launch_bounds__(256)
taylor_partial_kernel(const __grid_constant__ RedPtrs f, int64_t ncol, int64_t S, int nlevp1, double2* partial) {
  const int fld = blockIdx.y;
  const R* a = static_cast<const R*>(f.a[fld]);
  const R* b = static_cast<const R*>(f.b[fld]);
  const R* c = static_cast<const R*>(f.c[fld]);
  double s0 = 0.0, s1 = 0.0;
  // blocks stride over (level, 256-column chunk) tiles, four tiles in flight per thread (32-bit tile arithmetic: the
  // launcher's dims check bounds ncol_stride * (nlev + 1) below 2^32)
  const unsigned chunks = unsigned((ncol + blockDim.x - 1) / blockDim.x);
  const unsigned tiles = chunks * unsigned(nlevp1);
  for (unsigned tile = blockIdx.x; tile < tiles; tile += 4u * gridDim.x) {
    double va[4], vb[4], vc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned tt = tile + unsigned(u) * gridDim.x;
      const unsigned k = tt / chunks;
      const int64_t i = int64_t(tt - k * chunks) * blockDim.x + threadIdx.x;
      const bool ok = tt < tiles && i < ncol;
      const int64_t off = int64_t(k) * S + i;
      va[u] = (ok && a) ? double(a[off]) : 0.0;
      vb[u] = (ok && b) ? double(b[off]) : 0.0;
      vc[u] = (ok && c) ? double(c[off]) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s0 += va[u] - vb[u];
      s1 += vc[u];
    }
  }
  __shared__ double sh0[8], sh1[8];
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh0[w] = s0; sh1[w] = s1; }
  __syncthreads();
  if (w == 0) {
    s0 = lane < 8 ? sh0[lane] : 0.0;
    s1 = lane < 8 ? sh1[lane] : 0.0;
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) partial[int64_t(fld) * gridDim.x + blockIdx.x] = make_double2(s0, s1);
  }
}

// one warp per field: sums[2f], sums[2f+1] += fixed-order sum of the partials
__global__ void taylor_final_kernel(const double2* partial, int nblk, double* sums) {
  const int fld = blockIdx.x;
  double s0 = 0.0, s1 = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 32) {
    const double2 v = partial[int64_t(fld) * nblk + b];
    s0 += v.x;
    s1 += v.y;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (threadIdx.x == 0) {
    sums[2 * fld] += s0;
    sums[2 * fld + 1] += s1;
  }
}

// norm[i] = SUM_k SUM_f a_f[k, i] * b_f[k, i].  A CTA of 8 warps owns 32 columns: warp w sums the field pairs w, w + 8, ...
// of those columns over all levels (a warp reads 32 consecutive columns of one level of one field: 256 bytes, coalesced),
// the eight partial sums of a column are added in a fixed order through shared memory (deterministic).  65 536 columns are
// 2 048 CTAs = 16 384 warps: the memory-level parallelism comes from full occupancy, not from the compiler batching the
// loads of one thread (one thread per column with all fields in its loop ran at 17 % of the HBM peak, profiles/r2k).
// A pair with a_f == b_f (norm1 = <TL x, TL x>) is read once.
constexpr int kNormWarps = 8;
template <class R>
__global__ void __launch_bounds__(32 * kNormWarps)
symmetry_norm_kernel(const __grid_constant__ RedPtrs f, int nfields, int64_t ncol, int64_t S, int nlevp1,
                     double* __restrict__ norm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = int64_t(blockIdx.x) * 32 + lane;
  double acc0 = 0.0, acc1 = 0.0;
  if (i < ncol) {
    for (int n = w; n < nfields; n += kNormWarps) {
      const R* a = static_cast<const R*>(f.a[n]) + i;
      const R* b = static_cast<const R*>(f.b[n]) + i;
      if (a == b) {
        int k = 0;
        for (; k + 1 < nlevp1; k += 2) {
          const double x0 = double(a[int64_t(k) * S]), x1 = double(a[int64_t(k + 1) * S]);
          acc0 = fma(x0, x0, acc0);
          acc1 = fma(x1, x1, acc1);
        }
        if (k < nlevp1) {
          const double x0 = double(a[int64_t(k) * S]);
          acc0 = fma(x0, x0, acc0);
        }
      } else {
        int k = 0;
        for (; k + 1 < nlevp1; k += 2) {
          const double x0 = double(a[int64_t(k) * S]), x1 = double(a[int64_t(k + 1) * S]);
          const double y0 = double(b[int64_t(k) * S]), y1 = double(b[int64_t(k + 1) * S]);
          acc0 = fma(x0, y0, acc0);
          acc1 = fma(x1, y1, acc1);
        }
        if (k < nlevp1) acc0 = fma(double(a[int64_t(k) * S]), double(b[int64_t(k) * S]), acc0);
      }
    }
  }
  __shared__ double sh[kNormWarps][32];
  sh[w][lane] = acc0 + acc1;
  __syncthreads();
  if (w == 0 && i < ncol) {
    double s = sh[0][lane];
#pragma unroll
    for (int j = 1; j < kNormWarps; ++j) s += sh[j][lane];
    norm[i] = s;
  }
}

// Symmetry-test residual (adjoint/validation.py:157-160): norm3[i] = |n1 - n2| / eps if n2 == 0 else |n1 - n2| / (eps * n2),
// and its maximum over the columns.  Stage 1: per-block maxima (NaN propagates, like numpy's max); stage 2: one warp.
__global__ void __launch_bounds__(256)
symmetry_residual_kernel(const double* __restrict__ n1, const double* __restrict__ n2, int64_t ncol, double eps,
                         double* __restrict__ norm3, double* __restrict__ partial) {
  double m = -INFINITY;
  bool nan = false;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < ncol; i += int64_t(gridDim.x) * blockDim.x) {
    const double a = n1[i], b = n2[i];
    const double d = fabs(a - b);
    const double r = (b == 0.0) ? d / eps : d / (eps * b);
    if (norm3) norm3[i] = r;
    nan = nan || (r != r);
    m = fmax(m, r);
  }
  __shared__ double sh[8];
  __shared__ int shn[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  nan = __any_sync(0xffffffffu, nan);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[w] = m; shn[w] = nan; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 1; j < 8; ++j) { m = fmax(m, sh[j]); nan = nan || shn[j]; }
    partial[blockIdx.x] = nan ? NAN : m;
  }
}

__global__ void symmetry_residual_final_kernel(const double* __restrict__ partial, int nblk, double* __restrict__ out) {
  double m = -INFINITY;
  bool nan = false;
  for (int b = threadIdx.x; b < nblk; b += 32) {
    const double v = partial[b];
    nan = nan || (v != v);
    m = fmax(m, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  nan = __any_sync(0xffffffffu, nan);
  if (threadIdx.x == 0) out[0] = nan ? NAN : m;
}

// ---------------------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------------------
template <class R>
void build_tables(const cs2_params& P, int nlev, const R* eta, void* out) {
  // window levels: 0.1 < eta[k] < 0.4 for k in [0, nlev-2]  (nonlinear/_stencils/cloudsc2.py:109-110)
  int32_t nw = 0;
  for (int k = 0; k + 1 < nlev; ++k)
    if (eta[k] > R(0.1) && eta[k] < R(0.4)) ++nw;
  char* b = static_cast<char*>(out);
  reinterpret_cast<int32_t*>(b)[0] = nlev;
  reinterpret_cast<int32_t*>(b)[1] = nw;
  reinterpret_cast<int32_t*>(b)[2] = 0;
  reinterpret_cast<int32_t*>(b)[3] = 0;
  size_t off = 16;
  R* scalm = reinterpret_cast<R*>(b + off);
  off += cs2::tables_align16(size_t(nlev) * sizeof(R));
  R* crh2 = reinterpret_cast<R*>(b + off);
  off += cs2::tables_align16(size_t(nlev) * size_t(nw + 1) * sizeof(R));
  int32_t* wlev = reinterpret_cast<int32_t*>(b + off);
  {
    int j = 0;
    for (int k = 0; k + 1 < nlev; ++k)
      if (eta[k] > R(0.1) && eta[k] < R(0.4)) wlev[j++] = k;
  }
  for (int k = 0; k < nlev; ++k) {
    // :127  scalm = ZSCAL * max(eta - 0.2, ZEPS1) ** 0.2
    const R x = eta[k] - R(0.2);
    const R m = x > R(P.ZEPS1) ? x : R(P.ZEPS1);
    scalm[k] = R(P.ZSCAL) * R(std::pow(m, R(0.2)));
    for (int j = 0; j <= nw; ++j) {
      // :165-186 for trpaus = candidate j
      const R trp = (j == 0) ? R(0.1) : eta[wlev[j - 1]];
      const R e = eta[k];
      const R u = (trp - R(0.25)) / R(0.15);
      const R mn = (trp - R(0.25)) < R(0) ? (trp - R(0.25)) : R(0);
      const R rh2 = R(0.35) + R(0.14) * (u * u) + R(0.04) * mn / R(0.15);
      const R rh1 = R(1), rh3 = R(1);
      R c;
      if (e < trp) {
        c = rh3;
      } else {
        const R deta2 = R(0.3);
        const R bound1 = trp + deta2;
        if (e < bound1) {
          c = rh3 + (rh2 - rh3) * (e - trp) / deta2;
        } else {
          const R deta1 = R(0.09) + R(0.16) * (R(0.4) - trp) / R(0.3);
          const R bound2 = R(1) - deta1;
          if (e < bound2)
            c = rh2;
          else
            c = rh1 + (rh2 - rh1) * R(std::sqrt((R(1) - e) / deta1));
        }
      }
      crh2[size_t(k) * size_t(nw + 1) + j] = c;
    }
  }
}

template <class R>
int launch_saturation(const cs2_dims* d, const cs2_params* P, const void* ap, const void* t, void* qsat, cudaStream_t st) {
  if (d->ncol == 0) return CS2_OK;
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, 1.0);
  dim3 grid((unsigned)((d->ncol + 4 * kPointBlock - 1) / (4 * kPointBlock)), (unsigned)d->nlev);
  saturation_kernel<R><<<grid, kPointBlock, 0, st>>>(p, P->LPHYLIN, static_cast<const R*>(ap), static_cast<const R*>(t),
                                                    static_cast<R*>(qsat), d->ncol, d->ncol_stride);
  return check_cuda(cudaGetLastError(), "saturation launch");
}

template <class R>
int launch_state(const cs2_dims* d, double f, int mode, int ignore_supsat, const void* const* in,
                 const void* const* in_i, void* const* out, cudaStream_t st) {
  if (d->ncol == 0) return CS2_OK;
  StatePtrs<R> ptrs;
  for (int n = 0; n < CS2_NSTATE; ++n) {
    ptrs.in[n] = static_cast<const R*>(in[n]);
    ptrs.in_i[n] = in_i ? static_cast<const R*>(in_i[n]) : nullptr;
    ptrs.out[n] = static_cast<R*>(out[n]);
  }
  dim3 grid((unsigned)((d->ncol + kPointBlock - 1) / kPointBlock), (unsigned)(d->nlev + 1));
  if (mode == 0)
    state_increment_kernel<R><<<grid, kPointBlock, 0, st>>>(ptrs, R(f), ignore_supsat, d->ncol, d->ncol_stride);
  else
    perturbed_state_kernel<R><<<grid, kPointBlock, 0, st>>>(ptrs, R(f), d->ncol, d->ncol_stride);
  return check_cuda(cudaGetLastError(), mode == 0 ? "state_increment launch" : "perturbed_state launch");
}

template <class R>
int launch_nl(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f,
              bool ad_ref, int32_t* jsel_out, R* ck, cudaStream_t st, R* cov_out = nullptr) {
  if (d->ncol == 0) return CS2_OK;
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, dt);
  const cs2::NLFields<R> nf = cs2::make_nl_fields<R>(*f);
  const cs2::Streams<R, cs2::I_NL> ns = cs2::nl_streams<R>(nf, d->ncol_stride);
  const unsigned grid = (unsigned)((d->ncol + kColumnBlock - 1) / kColumnBlock);
  const bool evap = P->LEVAPLS2 || P->LDRAIN1D;
  const bool tetens = P->LPHYLIN || P->LDRAIN1D;
#define CS2_LAUNCH_NL(E, T)                                                                                         \
  nl_kernel<R, cs2::Cfg<E, T>, false, false><<<grid, kColumnBlock, 0, st>>>(p, tables, nf, ns, d->ncol, d->ncol_stride, \
                                                                            d->nlev, ad_ref ? 1 : 0, jsel_out, nullptr, nullptr)
#ifdef CS2_EXPERIMENTS
  static const bool use_pipe = std::getenv("CS2_NL_PIPE") != nullptr;  // experiment switches (profiles/README.md)
  static const bool use_bulk = std::getenv("CS2_NL_BULK") != nullptr;
  static const bool use_split = std::getenv("CS2_NL_SPLIT") != nullptr;
  if (use_pipe && !evap && tetens && !ck && !cov_out && P->RVTMP2 == 0.0) {  // two levels in flight per thread
    nl_pipe_kernel<R><<<grid, kColumnBlock, 0, st>>>(p, tables, nf, ns, d->ncol, d->ncol_stride, d->nlev, ad_ref ? 1 : 0, jsel_out);
    return check_cuda(cudaGetLastError(), "cloudsc2_nl (pipelined) launch");
  }
  if (use_split && !jsel_out && !evap && d->nlev <= cs2::kSplitMaxLev) {
    const unsigned sgrid = (unsigned)((d->ncol + kSplitCols - 1) / kSplitCols);
#define CS2_LAUNCH_SPLIT(T, NM)                                                                                  \
  nl_split_kernel<R, cs2::Cfg<false, T>, NM><<<sgrid, 2 * kSplitCols, 0, st>>>(p, tables, nf, ns, d->ncol, d->ncol_stride, \
                                                                               d->nlev)
    if (P->RVTMP2 == 0.0) {
      if (tetens) CS2_LAUNCH_SPLIT(true, cs2::M_NZ); else CS2_LAUNCH_SPLIT(false, cs2::M_NZ);
    } else {
      if (tetens) CS2_LAUNCH_SPLIT(true, cs2::M_N); else CS2_LAUNCH_SPLIT(false, cs2::M_N);
    }
#undef CS2_LAUNCH_SPLIT
    return check_cuda(cudaGetLastError(), "cloudsc2_nl (split) launch");
  }
  if (use_bulk && !jsel_out && !evap && tetens) {
    nl_bulk_kernel<R, cs2::Cfg<false, true>><<<(unsigned)((d->ncol + kBulkBlock - 1) / kBulkBlock), kBulkBlock, 0, st>>>(
        p, tables, nf, ns, d->ncol, d->ncol_stride, d->nlev);
    return check_cuda(cudaGetLastError(), "cloudsc2_nl (bulk) launch");
  }
#endif
  if (cov_out)  // AD forward sweep with the evaporation branch (recompute mode): also stores the overlap carry
    nl_kernel<R, cs2::Cfg<true, true>, false, true><<<grid, kColumnBlock, 0, st>>>(p, tables, nf, ns, d->ncol, d->ncol_stride,
                                                                                   d->nlev, ad_ref ? 1 : 0, jsel_out, nullptr, cov_out);
  else if (ck)  // AD forward sweep with checkpointing of the transcendentals (evaporation off, Tetens path)
    nl_kernel<R, cs2::Cfg<false, true>, true, true><<<grid, kColumnBlock, 0, st>>>(p, tables, nf, ns, d->ncol, d->ncol_stride,
                                                                                   d->nlev, ad_ref ? 1 : 0, jsel_out, ck, nullptr);
  else if (jsel_out)  // AD forward sweep, recompute mode
    nl_kernel<R, cs2::Cfg<false, true>, false, true><<<grid, kColumnBlock, 0, st>>>(p, tables, nf, ns, d->ncol, d->ncol_stride,
                                                                                    d->nlev, ad_ref ? 1 : 0, jsel_out, nullptr, nullptr);
  else if (evap && tetens) CS2_LAUNCH_NL(true, true);
  else if (evap) CS2_LAUNCH_NL(true, false);
  else if (tetens) CS2_LAUNCH_NL(false, true);
  else CS2_LAUNCH_NL(false, false);
#undef CS2_LAUNCH_NL
  return check_cuda(cudaGetLastError(), "cloudsc2_nl launch");
}

template <class R>
int launch_nl_pert(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f,
                   const cs2_nl_fields* fi, double factor, cudaStream_t st) {
  if (d->ncol == 0) return CS2_OK;
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, dt);
  const cs2::NLFields<R> nf = cs2::make_nl_fields<R>(*f), ng = cs2::make_nl_fields<R>(*fi);
  const cs2::Streams<R, 2 * cs2::I_NL> ns = cs2::tl_streams<R>(nf, ng, d->ncol_stride);
  const unsigned grid = (unsigned)((d->ncol + kColumnBlock - 1) / kColumnBlock);
  const bool evap = P->LEVAPLS2 || P->LDRAIN1D;
  const bool tetens = P->LPHYLIN || P->LDRAIN1D;
#define CS2_LAUNCH_NLP(E, T) \
  nlp_kernel<R, cs2::Cfg<E, T>><<<grid, kColumnBlock, 0, st>>>(p, tables, nf, ng, R(factor), ns, d->ncol, d->ncol_stride, d->nlev)
  if (evap && tetens) CS2_LAUNCH_NLP(true, true);
  else if (evap) CS2_LAUNCH_NLP(true, false);
  else if (tetens) CS2_LAUNCH_NLP(false, true);
  else CS2_LAUNCH_NLP(false, false);
#undef CS2_LAUNCH_NLP
  return check_cuda(cudaGetLastError(), "cloudsc2_nl_perturbed launch");
}

template <class R>
int launch_tl(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* traj,
              const cs2_nl_fields* pert, cudaStream_t st) {
  const unsigned grid = (unsigned)((d->ncol + kWideBlock - 1) / kWideBlock);
  const auto f = cs2::make_nl_fields<R>(*traj), g = cs2::make_nl_fields<R>(*pert);
  const auto p = cs2::make_dev_params<R>(*P, dt);
  const auto ns = cs2::tl_streams<R>(f, g, d->ncol_stride);
  if (P->LEVAPLS2 || P->LDRAIN1D)
    tl_kernel<R, true><<<grid, kWideBlock, 0, st>>>(p, tables, f, g, ns, d->ncol, d->ncol_stride, d->nlev);
  else
    tl_kernel<R, false><<<grid, kWideBlock, 0, st>>>(p, tables, f, g, ns, d->ncol, d->ncol_stride, d->nlev);
  return check_cuda(cudaGetLastError(), "cloudsc2_tl launch");
}

template <class R>
int launch_tl_inc(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* traj,
                  const cs2_nl_fields* pert_out, double factor, int32_t ignore_supsat, double* norm1, cudaStream_t st) {
  const unsigned grid = (unsigned)((d->ncol + kWideBlock - 1) / kWideBlock);
  const auto f = cs2::make_nl_fields<R>(*traj), g = cs2::make_nl_fields<R>(*pert_out);
  const auto p = cs2::make_dev_params<R>(*P, dt);
  const auto ns = cs2::nl_streams<R>(f, d->ncol_stride);
#define CS2_LAUNCH_TLI(E, N) \
  tl_inc_kernel<R, E, N><<<grid, kWideBlock, 0, st>>>(p, tables, f, g, ns, d->ncol, d->ncol_stride, d->nlev, R(factor), ignore_supsat, norm1)
  const bool evap = P->LEVAPLS2 || P->LDRAIN1D;
  if (evap && norm1) CS2_LAUNCH_TLI(true, true);
  else if (evap) CS2_LAUNCH_TLI(true, false);
  else if (norm1) CS2_LAUNCH_TLI(false, true);
  else CS2_LAUNCH_TLI(false, false);
#undef CS2_LAUNCH_TLI
  return check_cuda(cudaGetLastError(), "cloudsc2_tl_increment launch");
}

template <class R>
int launch_taylor_nl(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f,
                     double factor1, int32_t ignore_supsat, double factor2, double* sums, double2* partial,
                     cudaStream_t st) {
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, dt);
  const cs2::NLFields<R> nf = cs2::make_nl_fields<R>(*f);
  const auto ns = cs2::taylor_streams<R>(nf, d->ncol_stride);
  const unsigned grid = (unsigned)((d->ncol + kColumnBlock - 1) / kColumnBlock);
  const bool evap = P->LEVAPLS2 || P->LDRAIN1D;
  const bool tetens = P->LPHYLIN || P->LDRAIN1D;
#define CS2_LAUNCH_TNL(E, T)                                                                                              \
  taylor_nl_kernel<R, cs2::Cfg<E, T>><<<grid, kColumnBlock, 0, st>>>(p, tables, nf, R(factor1), R(factor2), ignore_supsat, ns, \
                                                                     d->ncol, d->ncol_stride, d->nlev, partial)
  if (evap && tetens) CS2_LAUNCH_TNL(true, true);
  else if (evap) CS2_LAUNCH_TNL(true, false);
  else if (tetens) CS2_LAUNCH_TNL(false, true);
  else CS2_LAUNCH_TNL(false, false);
#undef CS2_LAUNCH_TNL
  if (int rc = check_cuda(cudaGetLastError(), "taylor_nl launch")) return rc;
  taylor_final_kernel<<<cs2::T_N, 32, 0, st>>>(partial, (int)grid, sums);
  return check_cuda(cudaGetLastError(), "taylor final launch");
}

cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

}  // namespace

namespace {
// AD workspace: [jsel: int32 per column] [checkpoint planes | overlap-carry plane]; with -DCS2_EXPERIMENTS also [chunk ticket
// counter] [chunk hand-over slots: 3 values per column and chunk boundary, up to kMaxAdChunks - 1 boundaries]
#ifdef CS2_EXPERIMENTS
constexpr int kMaxAdChunks = 8;
#endif
struct AdWorkspace {
  size_t off_extra, off_sync, sync_bytes, off_carry, total;
};
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
inline AdWorkspace ad_workspace(const cs2_dims* d, const cs2_params* P, int mode) {
  AdWorkspace w;
  const size_t es = d->dtype == CS2_F32 ? 4 : 8;
  const size_t plane = size_t(d->nlev) * size_t(d->ncol_stride) * es;
  w.off_extra = align256(size_t(d->ncol_stride) * sizeof(int32_t));
  size_t extra = 0;
  if (P && (P->LEVAPLS2 || P->LDRAIN1D)) extra = plane;        // evaporation branch: overlap carry per (level, column)
  else if (mode == CS2_AD_CHECKPOINT) extra = size_t(cs2::CK_N) * plane;  // CK_N transcendental results per (level, column)
  w.off_sync = align256(w.off_extra + extra);
#ifdef CS2_EXPERIMENTS
  w.sync_bytes = 256;
  w.off_carry = w.off_sync + w.sync_bytes;
  w.total = align256(w.off_carry + size_t(kMaxAdChunks - 1) * 3 * size_t(d->ncol_stride) * es);
#else
  w.sync_bytes = 0;
  w.off_carry = w.off_sync;
  w.total = w.off_sync;
#endif
  return w;
}

#ifdef CS2_EXPERIMENTS
// How many level chunks the backward sweep is cut into (0: whole-column sweeps).  Whole columns need ceil(warps / slots)
// rounds, chunked ones ceil(warps * C / slots) / C: worth it when a mostly empty last round would cost a full one.
template <class R>
int ad_bwd_chunks(const cs2_dims* d, int* slots_cta_out) {
  static const int forced = []() { const char* e = std::getenv("CS2_AD_CHUNKS"); return e ? std::atoi(e) : -1; }();
  int dev = 0, nsm = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ad_bwd_chunked_kernel<R>, kWideBlock, 0) != cudaSuccess ||
      nsm < 1 || per_sm < 1) {
    (void)cudaGetLastError();
    return 0;
  }
  const int64_t slots = int64_t(nsm) * per_sm;  // CTAs resident at once
  *slots_cta_out = int(slots);
  if (forced >= 0) return (forced <= 1 || forced > kMaxAdChunks || d->nlev < 2 * forced) ? 0 : forced;
  return 0;  // measured: no gain (the sweep is FP64-issue-bound, not round-bound), so never chosen automatically
  const int64_t ctas = (d->ncol + kWideBlock - 1) / kWideBlock;
  const double rounds = double((ctas + slots - 1) / slots);
  int best = 0;
  double best_rounds = 0.93 * rounds;  // must save at least 7 % to pay for the hand-overs
  for (int c : {4, 8}) {
    if (d->nlev < 4 * c) continue;
    const double r = double((ctas * c + slots - 1) / slots) / c;
    if (r < best_rounds) {
      best_rounds = r;
      best = c;
    }
  }
  return best;
}

#endif  // CS2_EXPERIMENTS

template <class R>
int launch_ad(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* traj,
              const cs2_ad_seeds* seeds, const cs2_ad_outputs* adj, void* ws, int mode, cudaStream_t st,
              double factor = 0.0, int32_t ignore_supsat = 0, double* norm2 = nullptr) {
  int32_t* jsel = static_cast<int32_t*>(ws);
  const AdWorkspace wsl = ad_workspace(d, P, mode);
  R* const after_jsel = reinterpret_cast<R*>(static_cast<char*>(ws) + wsl.off_extra);
  // evaporation branch (LEVAPLS2 / LDRAIN1D, non-default): always the recompute sweep, plus the overlap carry per level
  const bool evap = P->LEVAPLS2 || P->LDRAIN1D;
  R* ck = (!evap && mode == CS2_AD_CHECKPOINT) ? after_jsel : nullptr;
  R* cov = evap ? after_jsel : nullptr;
  const bool ad_ref = !P->AD_TL_PREDICATES;
  if (int rc = launch_nl<R>(d, P, dt, tables, traj, ad_ref, jsel, ck, st, cov)) return rc;
  if (d->ncol == 0) return CS2_OK;
  cs2::ADSeeds<R> s;
  s.tnd_t = static_cast<R*>(seeds->in_tnd_t_i); s.tnd_q = static_cast<R*>(seeds->in_tnd_q_i);
  s.tnd_ql = static_cast<R*>(seeds->in_tnd_ql_i); s.tnd_qi = static_cast<R*>(seeds->in_tnd_qi_i);
  s.clc = static_cast<R*>(seeds->in_clc_i); s.covptot = static_cast<R*>(seeds->in_covptot_i);
  s.fhpsl = static_cast<R*>(seeds->in_fhpsl_i); s.fhpsn = static_cast<R*>(seeds->in_fhpsn_i);
  s.fplsl = static_cast<R*>(seeds->in_fplsl_i); s.fplsn = static_cast<R*>(seeds->in_fplsn_i);
  cs2::ADOut<R> a;
  a.aph = static_cast<R*>(adj->out_aph_i); a.ap = static_cast<R*>(adj->out_ap_i); a.q = static_cast<R*>(adj->out_q_i);
  a.qsat = static_cast<R*>(adj->out_qsat_i); a.t = static_cast<R*>(adj->out_t_i); a.ql = static_cast<R*>(adj->out_ql_i);
  a.qi = static_cast<R*>(adj->out_qi_i); a.lude = static_cast<R*>(adj->out_lude_i); a.lu = static_cast<R*>(adj->out_lu_i);
  a.mfu = static_cast<R*>(adj->out_mfu_i); a.mfd = static_cast<R*>(adj->out_mfd_i);
  a.supsat = static_cast<R*>(adj->out_supsat_i); a.tnd_t = static_cast<R*>(adj->out_tnd_cml_t_i);
  a.tnd_q = static_cast<R*>(adj->out_tnd_cml_q_i); a.tnd_ql = static_cast<R*>(adj->out_tnd_cml_ql_i);
  a.tnd_qi = static_cast<R*>(adj->out_tnd_cml_qi_i);
  const unsigned grid = (unsigned)((d->ncol + kWideBlock - 1) / kWideBlock);
  const cs2::NLFields<R> nf = cs2::make_nl_fields<R>(*traj);
  // seeds are reset by the backward kernel itself (each thread after its own column); CS2_AD_SEED_MEMSET=1 restores
  // the separate cudaMemsetAsync calls (A/B switch for profiles/)
  static const bool use_memset = std::getenv("CS2_AD_SEED_MEMSET") != nullptr;
  const int zero_in_kernel = use_memset ? 0 : 1;
  const auto dp = cs2::make_dev_params<R>(*P, dt);
#define CS2_LAUNCH_BWD(NS, E, N, CK, COV)                                                                             \
  ad_bwd_kernel<R, NS, E, N><<<grid, kWideBlock, 0, st>>>(dp, tables, nf, a,                                             \
                                                          cs2::ad_streams<R, NS, E>(nf, s, d->ncol_stride, d->nlev, CK, COV), \
                                                          jsel, d->ncol, d->ncol_stride, d->nlev, R(factor), ignore_supsat,  \
                                                          norm2, s, zero_in_kernel)
#ifdef CS2_EXPERIMENTS  // level-chunked sweep with persistent warps: CS2_AD_CHUNKS=C (profiles/r2n_ad_chunked.md)
  int slots_cta = 0;
  int nchunk = (!evap && !ck && !norm2) ? ad_bwd_chunks<R>(d, &slots_cta) : 0;
  if (nchunk > 1) {  // level-chunked sweep with persistent warps (see dev_ad_bwd_chunked)
    unsigned* ticket = reinterpret_cast<unsigned*>(static_cast<char*>(ws) + wsl.off_sync);
    R* carry = reinterpret_cast<R*>(static_cast<char*>(ws) + wsl.off_carry);
    const int chunk_levels = (d->nlev + nchunk - 1) / nchunk;
    nchunk = (d->nlev + chunk_levels - 1) / chunk_levels;  // no empty last chunk
    if (int rc = check_cuda(cudaMemsetAsync(ticket, 0, wsl.sync_bytes, st), "cloudsc2_ad chunk ticket reset")) return rc;
    // hand-over slots: all-ones = "not written yet"
    if (int rc = check_cuda(cudaMemsetAsync(carry, 0xFF, size_t(nchunk - 1) * 3 * size_t(d->ncol_stride) * sizeof(R), st),
                            "cloudsc2_ad chunk hand-over reset"))
      return rc;
    const int64_t want = ((d->ncol + 31) / 32 + 1) / 2;  // never more CTAs than pairs of column groups
    const unsigned pgrid = (unsigned)(want < slots_cta ? want : slots_cta);
    ad_bwd_chunked_kernel<R><<<pgrid, kWideBlock, 0, st>>>(dp, tables, nf, a, cs2::ad_streams<R, cs2::B_N, false>(nf, s, d->ncol_stride, d->nlev, nullptr, nullptr),
                                                          jsel, d->ncol, d->ncol_stride, d->nlev, s, zero_in_kernel, nchunk, chunk_levels, ticket, carry);
  } else
#endif
  if (evap) CS2_LAUNCH_BWD(cs2::B_N, true, false, nullptr, cov);
  else if (ck && norm2) CS2_LAUNCH_BWD(cs2::B_NCK, false, true, ck, nullptr);
  else if (norm2) CS2_LAUNCH_BWD(cs2::B_N, false, true, nullptr, nullptr);
  else if (ck) CS2_LAUNCH_BWD(cs2::B_NCK, false, false, ck, nullptr);
  else CS2_LAUNCH_BWD(cs2::B_N, false, false, nullptr, nullptr);
#undef CS2_LAUNCH_BWD
  if (int rc = check_cuda(cudaGetLastError(), "cloudsc2_ad backward launch")) return rc;
  if (zero_in_kernel) return CS2_OK;
  // the reference stencil consumes its seeds (adjoint/_stencils/cloudsc2.py:482-484,506-542,650,714,920,972-984)
  const size_t full = size_t(d->nlev) * size_t(d->ncol_stride) * sizeof(R);
  const size_t half = size_t(d->nlev + 1) * size_t(d->ncol_stride) * sizeof(R);
  R* const full_seeds[6] = {s.tnd_t, s.tnd_q, s.tnd_ql, s.tnd_qi, s.clc, s.covptot};
  R* const half_seeds[4] = {s.fhpsl, s.fhpsn, s.fplsl, s.fplsn};
  for (R* ptr : full_seeds)
    if (int rc = check_cuda(cudaMemsetAsync(ptr, 0, full, st), "cloudsc2_ad seed reset")) return rc;
  for (R* ptr : half_seeds)
    if (int rc = check_cuda(cudaMemsetAsync(ptr, 0, half, st), "cloudsc2_ad seed reset")) return rc;
  return CS2_OK;
}
}  // namespace


// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int cs2_abi_version(void) { return CS2_ABI_VERSION; }

const char* cs2_last_error(void) { return g_last_error.c_str(); }

int cs2_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
    (void)cudaGetLastError();
    return 0;
  }
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDeviceCount");
  return n;
}

size_t cs2_level_tables_bytes(int32_t nlev, int32_t dtype) {
  if (nlev < 1) return 0;
  const size_t es = dtype == CS2_F32 ? 4 : 8;
  return 16 + cs2::tables_align16(size_t(nlev) * es) + cs2::tables_align16(size_t(nlev) * size_t(nlev) * es) +
         cs2::tables_align16(size_t(nlev) * 4);
}

int cs2_level_tables_build(const cs2_params* params, int32_t nlev, int32_t dtype, const void* eta_host,
                           void* tables_host, size_t tables_bytes) {
  if (!params || !eta_host || !tables_host) return fail(CS2_ERR_NULL_POINTER, "level_tables_build: NULL argument");
  if (nlev < 1 || nlev > 4096 || (dtype != CS2_F64 && dtype != CS2_F32))
    return fail(CS2_ERR_BAD_DIMS, "level_tables_build: bad nlev or dtype");
  if (tables_bytes < cs2_level_tables_bytes(nlev, dtype))
    return fail(CS2_ERR_WORKSPACE, "level_tables_build: tables buffer too small");
  std::memset(tables_host, 0, tables_bytes);
  if (dtype == CS2_F64)
    build_tables<double>(*params, nlev, static_cast<const double*>(eta_host), tables_host);
  else
    build_tables<float>(*params, nlev, static_cast<const float*>(eta_host), tables_host);
  return CS2_OK;
}

int cs2_dfma_rate(double* scratch_dev, int32_t blocks, int32_t iters, void* stream) {
  if (!scratch_dev) return fail(CS2_ERR_NULL_POINTER, "dfma_rate: scratch is NULL");
  if (blocks < 1 || iters < 1) return fail(CS2_ERR_BAD_DIMS, "dfma_rate: blocks and iters must be positive");
  dfma_rate_kernel<<<blocks, 256, 0, as_stream(stream)>>>(scratch_dev, iters, 0.999999, 1e-6);
  return check_cuda(cudaGetLastError(), "dfma_rate launch");
}

int cs2_dfma_rate_regs(double* scratch_dev, int32_t blocks, int32_t iters, int32_t mode, void* stream) {
  if (!scratch_dev) return fail(CS2_ERR_NULL_POINTER, "dfma_rate_regs: scratch is NULL");
  if (blocks < 2 || iters < 1 || (mode != 1 && mode != 2)) return fail(CS2_ERR_BAD_DIMS, "dfma_rate_regs: bad blocks / iters / mode");
  // the first 256 doubles of the scratch buffer are the operand seeds (the caller fills them), results go behind them
  if (mode == 1) dfma_rate_regs_kernel<1><<<blocks - 1, 256, 0, as_stream(stream)>>>(scratch_dev + 256, iters, scratch_dev);
  else dfma_rate_regs_kernel<2><<<blocks - 1, 256, 0, as_stream(stream)>>>(scratch_dev + 256, iters, scratch_dev);
  return check_cuda(cudaGetLastError(), "dfma_rate_regs launch");
}

int cs2_saturation(const cs2_dims* dims, const cs2_params* params, const void* in_ap, const void* in_t,
                   void* out_qsat, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!params) return fail(CS2_ERR_NULL_POINTER, "saturation: params is NULL");
  const void* ptrs[3] = {in_ap, in_t, out_qsat};
  if (int rc = check_ptrs(ptrs, 3, "saturation")) return rc;
  return dims->dtype == CS2_F64 ? launch_saturation<double>(dims, params, in_ap, in_t, out_qsat, as_stream(stream))
                                : launch_saturation<float>(dims, params, in_ap, in_t, out_qsat, as_stream(stream));
}

int cs2_state_increment(const cs2_dims* dims, double f, int32_t ignore_supsat, const void* const in[CS2_NSTATE],
                        void* const out_i[CS2_NSTATE], void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!in || !out_i) return fail(CS2_ERR_NULL_POINTER, "state_increment: NULL field array");
  if (int rc = check_ptrs(in, CS2_NSTATE, "state_increment in")) return rc;
  if (int rc = check_ptrs(const_cast<const void* const*>(out_i), CS2_NSTATE, "state_increment out")) return rc;
  return dims->dtype == CS2_F64 ? launch_state<double>(dims, f, 0, ignore_supsat, in, nullptr, out_i, as_stream(stream))
                                : launch_state<float>(dims, f, 0, ignore_supsat, in, nullptr, out_i, as_stream(stream));
}

int cs2_perturbed_state(const cs2_dims* dims, double f, const void* const in[CS2_NSTATE],
                        const void* const in_i[CS2_NSTATE], void* const out[CS2_NSTATE], void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!in || !in_i || !out) return fail(CS2_ERR_NULL_POINTER, "perturbed_state: NULL field array");
  if (int rc = check_ptrs(in, CS2_NSTATE, "perturbed_state in")) return rc;
  if (int rc = check_ptrs(in_i, CS2_NSTATE, "perturbed_state in_i")) return rc;
  if (int rc = check_ptrs(const_cast<const void* const*>(out), CS2_NSTATE, "perturbed_state out")) return rc;
  return dims->dtype == CS2_F64 ? launch_state<double>(dims, f, 1, 0, in, in_i, out, as_stream(stream))
                                : launch_state<float>(dims, f, 1, 0, in, in_i, out, as_stream(stream));
}

int cs2_nl(const cs2_dims* dims, const cs2_params* params, double dt, const void* level_tables_dev,
           const cs2_nl_fields* f, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!params || !level_tables_dev) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_nl: params or level tables NULL");
  if (int rc = check_nl_fields(f, "cloudsc2_nl fields")) return rc;
  return dims->dtype == CS2_F64
             ? launch_nl<double>(dims, params, dt, level_tables_dev, f, false, nullptr, nullptr, as_stream(stream))
             : launch_nl<float>(dims, params, dt, level_tables_dev, f, false, nullptr, nullptr, as_stream(stream));
}

int cs2_nl_perturbed(const cs2_dims* dims, const cs2_params* params, double dt, const void* level_tables_dev,
                     const cs2_nl_fields* f, const cs2_nl_fields* in_i, double factor, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!params || !level_tables_dev) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_nl_perturbed: params or level tables NULL");
  if (int rc = check_nl_fields(f, "cloudsc2_nl_perturbed fields")) return rc;
  if (!in_i) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_nl_perturbed: increment fields NULL");
  if (int rc = check_ptrs(reinterpret_cast<const void* const*>(in_i), 16, "cloudsc2_nl_perturbed increments")) return rc;
  cs2_nl_fields gi = *in_i;  // only the 16 in_* members of `in_i` are read; outputs alias the base outputs (unused)
  gi.out_clc = f->out_clc; gi.out_covptot = f->out_covptot; gi.out_fhpsl = f->out_fhpsl; gi.out_fhpsn = f->out_fhpsn;
  gi.out_fplsl = f->out_fplsl; gi.out_fplsn = f->out_fplsn; gi.out_tnd_q = f->out_tnd_q; gi.out_tnd_qi = f->out_tnd_qi;
  gi.out_tnd_ql = f->out_tnd_ql; gi.out_tnd_t = f->out_tnd_t;
  return dims->dtype == CS2_F64
             ? launch_nl_pert<double>(dims, params, dt, level_tables_dev, f, &gi, factor, as_stream(stream))
             : launch_nl_pert<float>(dims, params, dt, level_tables_dev, f, &gi, factor, as_stream(stream));
}

int cs2_tl(const cs2_dims* dims, const cs2_params* params, double dt, const void* level_tables_dev,
           const cs2_nl_fields* traj, const cs2_nl_fields* pert, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!params || !level_tables_dev) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_tl: params or level tables NULL");
  if (int rc = check_nl_fields(traj, "cloudsc2_tl trajectory fields")) return rc;
  if (int rc = check_nl_fields(pert, "cloudsc2_tl perturbation fields")) return rc;
  if (dims->ncol == 0) return CS2_OK;
  return dims->dtype == CS2_F64 ? launch_tl<double>(dims, params, dt, level_tables_dev, traj, pert, as_stream(stream))
                                : launch_tl<float>(dims, params, dt, level_tables_dev, traj, pert, as_stream(stream));
}

int cs2_tl_increment(const cs2_dims* dims, const cs2_params* params, double dt, const void* level_tables_dev,
                     const cs2_nl_fields* traj, const cs2_nl_fields* pert_out, double factor, int32_t ignore_supsat,
                     double* norm1_dev, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!params || !level_tables_dev) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_tl_increment: params or level tables NULL");
  if (int rc = check_nl_fields(traj, "cloudsc2_tl_increment trajectory fields")) return rc;
  if (!pert_out) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_tl_increment: perturbation outputs NULL");
  if (int rc = check_ptrs(reinterpret_cast<const void* const*>(pert_out) + 16, 10, "cloudsc2_tl_increment perturbation outputs"))
    return rc;
  if (dims->ncol == 0) return CS2_OK;
  return dims->dtype == CS2_F64
             ? launch_tl_inc<double>(dims, params, dt, level_tables_dev, traj, pert_out, factor, ignore_supsat, norm1_dev,
                                     as_stream(stream))
             : launch_tl_inc<float>(dims, params, dt, level_tables_dev, traj, pert_out, factor, ignore_supsat, norm1_dev,
                                    as_stream(stream));
}

size_t cs2_ad_workspace_bytes(const cs2_dims* dims, const cs2_params* params, int32_t mode) {
  if (!dims || dims->ncol_stride <= 0) return 0;
  return ad_workspace(dims, params, mode).total;
}

int cs2_ad(const cs2_dims* dims, const cs2_params* params, double dt, const void* level_tables_dev,
           const cs2_nl_fields* traj, const cs2_ad_seeds* seeds, const cs2_ad_outputs* adj, void* workspace_dev,
           size_t workspace_bytes, int32_t mode, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!params || !level_tables_dev) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_ad: params or level tables NULL");
  if (int rc = check_nl_fields(traj, "cloudsc2_ad trajectory fields")) return rc;
  if (!seeds || !adj) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_ad: seeds or adjoint outputs NULL");
  static_assert(sizeof(cs2_ad_seeds) == 10 * sizeof(void*), "cs2_ad_seeds layout");
  static_assert(sizeof(cs2_ad_outputs) == 16 * sizeof(void*), "cs2_ad_outputs layout");
  if (int rc = check_ptrs(reinterpret_cast<const void* const*>(seeds), 10, "cloudsc2_ad seeds")) return rc;
  if (int rc = check_ptrs(reinterpret_cast<const void* const*>(adj), 16, "cloudsc2_ad adjoint outputs")) return rc;
  if (mode != CS2_AD_RECOMPUTE && mode != CS2_AD_CHECKPOINT) return fail(CS2_ERR_BAD_DIMS, "cloudsc2_ad: unknown mode");
  if (!workspace_dev || workspace_bytes < cs2_ad_workspace_bytes(dims, params, mode))
    return fail(CS2_ERR_WORKSPACE, "cloudsc2_ad: workspace missing or too small");
  return dims->dtype == CS2_F64
             ? launch_ad<double>(dims, params, dt, level_tables_dev, traj, seeds, adj, workspace_dev, mode, as_stream(stream))
             : launch_ad<float>(dims, params, dt, level_tables_dev, traj, seeds, adj, workspace_dev, mode, as_stream(stream));
}

int cs2_ad_norm2(const cs2_dims* dims, const cs2_params* params, double dt, const void* level_tables_dev,
           const cs2_nl_fields* traj, const cs2_ad_seeds* seeds, const cs2_ad_outputs* adj, void* workspace_dev,
           size_t workspace_bytes, int32_t mode, double factor, int32_t ignore_supsat, double* norm2_dev,
                 void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!params || !level_tables_dev) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_ad: params or level tables NULL");
  if (int rc = check_nl_fields(traj, "cloudsc2_ad trajectory fields")) return rc;
  if (!seeds || !adj) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_ad: seeds or adjoint outputs NULL");
  static_assert(sizeof(cs2_ad_seeds) == 10 * sizeof(void*), "cs2_ad_seeds layout");
  static_assert(sizeof(cs2_ad_outputs) == 16 * sizeof(void*), "cs2_ad_outputs layout");
  if (int rc = check_ptrs(reinterpret_cast<const void* const*>(seeds), 10, "cloudsc2_ad seeds")) return rc;
  if (int rc = check_ptrs(reinterpret_cast<const void* const*>(adj), 16, "cloudsc2_ad adjoint outputs")) return rc;
  if (!norm2_dev) return fail(CS2_ERR_NULL_POINTER, "cloudsc2_ad_norm2: norm2_dev is NULL");
  if (params->LEVAPLS2 || params->LDRAIN1D)
    return fail(CS2_ERR_UNSUPPORTED, "cloudsc2_ad_norm2: the fused inner product is not built for the evaporation branch "
                                     "(LEVAPLS2 / LDRAIN1D); use cs2_ad + cs2_symmetry_norms");
  if (mode != CS2_AD_RECOMPUTE && mode != CS2_AD_CHECKPOINT) return fail(CS2_ERR_BAD_DIMS, "cloudsc2_ad: unknown mode");
  if (!workspace_dev || workspace_bytes < cs2_ad_workspace_bytes(dims, params, mode))
    return fail(CS2_ERR_WORKSPACE, "cloudsc2_ad: workspace missing or too small");
  return dims->dtype == CS2_F64
             ? launch_ad<double>(dims, params, dt, level_tables_dev, traj, seeds, adj, workspace_dev, mode, as_stream(stream), factor,
                                  ignore_supsat, norm2_dev)
             : launch_ad<float>(dims, params, dt, level_tables_dev, traj, seeds, adj, workspace_dev, mode, as_stream(stream), factor,
                                  ignore_supsat, norm2_dev);
}

static int taylor_blocks(const cs2_dims* dims) {
  const int64_t n = dims->ncol * int64_t(dims->nlev + 1);
  int64_t nb = (n + 256 * 8 - 1) / (256 * 8);
  if (nb < 1) nb = 1;
  if (nb > 148 * 8) nb = 148 * 8;
  return int(nb);
}

size_t cs2_taylor_scratch_bytes(const cs2_dims* dims, int32_t nfields) {
  if (!dims || nfields < 1) return 0;
  return size_t(taylor_blocks(dims)) * size_t(nfields) * sizeof(double2);
}

int cs2_taylor_sums(const cs2_dims* dims, int32_t nfields, const void* const* a_dev, const void* const* b_dev,
                    const void* const* c_dev, double* sums_dev, void* scratch_dev, size_t scratch_bytes,
                    void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (nfields < 1 || nfields > 16) return fail(CS2_ERR_BAD_DIMS, "taylor_sums: nfields outside [1, 16]");
  if (!a_dev || !sums_dev || !scratch_dev) return fail(CS2_ERR_NULL_POINTER, "taylor_sums: NULL argument");
  if (scratch_bytes < cs2_taylor_scratch_bytes(dims, nfields)) return fail(CS2_ERR_WORKSPACE, "taylor_sums: scratch too small");
  RedPtrs f;
  for (int n = 0; n < kMaxRedFields; ++n) {
    f.a[n] = n < nfields ? a_dev[n] : nullptr;
    f.b[n] = (n < nfields && b_dev) ? b_dev[n] : nullptr;
    f.c[n] = (n < nfields && c_dev) ? c_dev[n] : nullptr;
  }
  const int nb = taylor_blocks(dims);
  dim3 grid((unsigned)nb, (unsigned)nfields);
  double2* partial = static_cast<double2*>(scratch_dev);
  if (dims->dtype == CS2_F64)
    taylor_partial_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(f, dims->ncol, dims->ncol_stride, dims->nlev + 1, partial);
  else
    taylor_partial_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(f, dims->ncol, dims->ncol_stride, dims->nlev + 1, partial);
  if (int rc = check_cuda(cudaGetLastError(), "taylor partial launch")) return rc;
  taylor_final_kernel<<<nfields, 32, 0, as_stream(stream)>>>(partial, nb, sums_dev);
  return check_cuda(cudaGetLastError(), "taylor final launch");
}

size_t cs2_taylor_nl_scratch_bytes(const cs2_dims* dims) {
  if (!dims || dims->ncol < 0) return 0;
  return size_t((dims->ncol + kColumnBlock - 1) / kColumnBlock) * size_t(cs2::T_N) * sizeof(double2);
}

int cs2_taylor_nl_sums(const cs2_dims* dims, const cs2_params* params, double dt, const void* level_tables_dev,
                       const cs2_nl_fields* f, double factor1, int32_t ignore_supsat, double factor2, double* sums_dev,
                       void* scratch_dev, size_t scratch_bytes, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (!params || !level_tables_dev || !sums_dev || !scratch_dev)
    return fail(CS2_ERR_NULL_POINTER, "taylor_nl_sums: NULL argument");
  if (int rc = check_nl_fields(f, "taylor_nl_sums fields")) return rc;
  if (scratch_bytes < cs2_taylor_nl_scratch_bytes(dims)) return fail(CS2_ERR_WORKSPACE, "taylor_nl_sums: scratch too small");
  if (dims->ncol == 0) return CS2_OK;
  return dims->dtype == CS2_F64
             ? launch_taylor_nl<double>(dims, params, dt, level_tables_dev, f, factor1, ignore_supsat, factor2, sums_dev,
                                        static_cast<double2*>(scratch_dev), as_stream(stream))
             : launch_taylor_nl<float>(dims, params, dt, level_tables_dev, f, factor1, ignore_supsat, factor2, sums_dev,
                                       static_cast<double2*>(scratch_dev), as_stream(stream));
}

int cs2_symmetry_norms(const cs2_dims* dims, int32_t nfields, const void* const* a_dev, const void* const* b_dev,
                       double* norm_dev, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (nfields < 1 || nfields > kMaxRedFields) return fail(CS2_ERR_BAD_DIMS, "symmetry_norms: nfields outside [1, 32]");
  if (!a_dev || !b_dev || !norm_dev) return fail(CS2_ERR_NULL_POINTER, "symmetry_norms: NULL argument");
  if (int rc = check_ptrs(a_dev, nfields, "symmetry_norms a")) return rc;
  if (int rc = check_ptrs(b_dev, nfields, "symmetry_norms b")) return rc;
  if (dims->ncol == 0) return CS2_OK;
  RedPtrs f;
  for (int n = 0; n < kMaxRedFields; ++n) {
    f.a[n] = n < nfields ? a_dev[n] : nullptr;
    f.b[n] = n < nfields ? b_dev[n] : nullptr;
    f.c[n] = nullptr;
  }
  const unsigned grid = (unsigned)((dims->ncol + 31) / 32);
  if (dims->dtype == CS2_F64)
    symmetry_norm_kernel<double><<<grid, 32 * kNormWarps, 0, as_stream(stream)>>>(f, nfields, dims->ncol, dims->ncol_stride, dims->nlev + 1, norm_dev);
  else
    symmetry_norm_kernel<float><<<grid, 32 * kNormWarps, 0, as_stream(stream)>>>(f, nfields, dims->ncol, dims->ncol_stride, dims->nlev + 1, norm_dev);
  return check_cuda(cudaGetLastError(), "symmetry norm launch");
}

static int residual_blocks(int64_t ncol) {
  int64_t nb = (ncol + 255) / 256;
  if (nb < 1) nb = 1;
  if (nb > 148 * 4) nb = 148 * 4;
  return int(nb);
}

size_t cs2_symmetry_residual_scratch_bytes(int64_t ncol) { return size_t(residual_blocks(ncol)) * sizeof(double); }

int cs2_symmetry_residual(int64_t ncol, const double* norm1_dev, const double* norm2_dev, double eps, double* norm3_dev,
                          double* max_dev, void* scratch_dev, size_t scratch_bytes, void* stream) {
  if (ncol < 0) return fail(CS2_ERR_BAD_DIMS, "symmetry_residual: ncol < 0");
  if (!max_dev || !scratch_dev || (ncol > 0 && (!norm1_dev || !norm2_dev)))
    return fail(CS2_ERR_NULL_POINTER, "symmetry_residual: NULL argument");
  if (!(eps > 0.0)) return fail(CS2_ERR_BAD_DIMS, "symmetry_residual: eps must be positive");
  if (scratch_bytes < cs2_symmetry_residual_scratch_bytes(ncol)) return fail(CS2_ERR_WORKSPACE, "symmetry_residual: scratch too small");
  const int nb = residual_blocks(ncol);
  symmetry_residual_kernel<<<nb, 256, 0, as_stream(stream)>>>(norm1_dev, norm2_dev, ncol, eps, norm3_dev,
                                                              static_cast<double*>(scratch_dev));
  if (int rc = check_cuda(cudaGetLastError(), "symmetry residual launch")) return rc;
  symmetry_residual_final_kernel<<<1, 32, 0, as_stream(stream)>>>(static_cast<const double*>(scratch_dev), nb, max_dev);
  return check_cuda(cudaGetLastError(), "symmetry residual final launch");
}

}  // extern "C"
