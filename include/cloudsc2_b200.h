/*
 * cloudsc2_b200 -- C ABI of the B200-native CLOUDSC2 NL / TL / AD column physics.
 *
 * The reference (cloudsc2_gt4py) has no FFI: its plug-in point is the GT4Py StencilObject
 * returned by `compile_stencil(name, externals)` and invoked from each component's
 * `array_call` with keyword field arguments.  Every entry point below replaces one such
 * stencil object (the reference interface it stands for is cited per function).  The Python
 * host side (`cloudsc2_b200.framework.stencil`) binds these symbols with ctypes and exposes
 * them under the reference's stencil names and keyword signatures.
 *
 * Conventions
 *   - Every field is ONE device allocation laid out `[nlev+1][ncol_stride]`, column index
 *     fastest (the `(K, IJ)` layout of the reference's HDF5 files); `ncol_stride` is a
 *     multiple of 32 elements and the base pointer is 256-byte aligned, so a warp reading 32
 *     consecutive columns of one level issues one fully coalesced request.  Full-level fields
 *     carry one padding level (index nlev) which the kernels never write, exactly like the
 *     reference's `(nx, 1, nz+1)` storages.
 *   - `dtype` selects the arithmetic type of ALL fields and of the computation:
 *     CS2_F64 (double) or CS2_F32 (float).  Scalars are passed as double and cast.
 *   - All pointers in the field structs are DEVICE pointers owned by the caller.  The library
 *     allocates no device memory.  Host pointers are named `*_host`.
 *   - Calls are asynchronous and ordered on `stream` (a cudaStream_t passed as void*; NULL =
 *     legacy default stream).  The library is stateless and re-entrant apart from the
 *     thread-local error string.
 *   - Return value: CS2_OK (0) or a negative CS2_ERR_* code; `cs2_last_error()` returns a
 *     human-readable message for the last failure on the calling thread.  No exception or
 *     abort crosses the boundary.  There is no CPU fallback: without a CUDA device every
 *     compute entry point fails with CS2_ERR_CUDA.
 */
#ifndef CLOUDSC2_B200_H
#define CLOUDSC2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS2_ABI_VERSION 1

enum { CS2_F64 = 0, CS2_F32 = 1 };

enum {
  CS2_OK = 0,
  CS2_ERR_BAD_DIMS = -1,     /* ncol/nlev/stride/dtype out of range                     */
  CS2_ERR_NULL_POINTER = -2, /* a required pointer is NULL                              */
  CS2_ERR_MISALIGNED = -3,   /* a field pointer is not 16-byte aligned or stride % 32   */
  CS2_ERR_UNSUPPORTED = -4,  /* an opt-in fused entry point is not built for these flags   */
  CS2_ERR_CUDA = -5,         /* CUDA runtime error (message carries cudaGetErrorString) */
  CS2_ERR_WORKSPACE = -6     /* workspace / table buffer too small                      */
};

/* Grid of one call: columns [0, ncol) x full levels [0, nlev); half-level fields have nlev+1
 * levels.  Mirrors ComputationalGrid(GridConfig(nx, ny=1, nz)) of the reference drivers
 * (drivers/run_nonlinear.py:57). */
typedef struct cs2_dims {
  int64_t ncol;        /* nx                                                   */
  int64_t ncol_stride; /* elements between consecutive levels (multiple of 32) */
  int32_t nlev;        /* nz (137 in all shipped data)                         */
  int32_t dtype;       /* CS2_F64 | CS2_F32                                    */
} cs2_dims;

/* The `externals` dict of the reference components, restricted to the names the stencils
 * actually import (nonlinear/_stencils/cloudsc2.py:61-91, tangent_linear/_stencils/
 * cloudsc2.py:91-122, the three cuadjtqs.py files, common/_stencils/fcttre.py, saturation.py:27). */
typedef struct cs2_params {
  /* YOETHF (iox.py:25-45) */
  double R2ES, R3IES, R3LES, R4IES, R4LES, R5ALSCP, R5ALVCP, R5IES, R5LES;
  double RALSDCP, RALVDCP, RTICE, RTICECU, RTWAT, RTWAT_RTICE_R, RTWAT_RTICECU_R, RVTMP2;
  /* YOMCST (iox.py:48-57) */
  double RCPD, RD, RETV, RG, RLMLT, RLSTT, RLVTT, RTT;
  /* YRECLDP / YREPHLI (iox.py:60-201) */
  double RCLCRIT, RKCONV, RLMIN, RPECONS, RLPTRC;
  /* component constants (nonlinear/microphysics.py:68-78, common/saturation.py:51) */
  double ZEPS1, ZEPS2, ZQMAX, ZSCAL, QMAX;
  /* flags */
  int32_t LPHYLIN, LDRAIN1D, LEVAPLS2, LREGCL, KFLAG, ICALL;
  /* AD branch predicates: 0 = literal reference (second freezing test on the pre-adjustment
   * temperature, adjoint/_stencils/cloudsc2.py:427,577; backward first freezing test on the
   * post-adjustment temperature, :729); 1 = the TL predicates on the same trajectory
   * (tangent_linear/_stencils/cloudsc2.py:510,677), i.e. the exact adjoint of the TL. */
  int32_t AD_TL_PREDICATES;
  int32_t reserved_;
} cs2_params;

/* ---------------------------------------------------------------------------------------
 * library / error handling
 * ------------------------------------------------------------------------------------- */
int cs2_abi_version(void);
const char* cs2_last_error(void);
/* Number of CUDA devices visible to the library (0 without a GPU); < 0 on runtime error. */
int cs2_device_count(void);

/* Measurement helper: `blocks` CTAs of 256 threads each run 8 independent chains of `iters` dependent
 * DFMAs (flops = blocks * 256 * 8 * iters * 2) and write one double per thread to `scratch_dev`
 * (>= blocks * 256 doubles).  Timed with CUDA events by bench.py to put the FP64-pipe peak of the
 * device next to the HBM peak (the roofline's second axis). */
int cs2_dfma_rate(double* scratch_dev, int32_t blocks, int32_t iters, void* stream);
/* Same measurement with every DFMA operand in a per-thread register (mode 1: x = fma(x, y, z); mode 2: the dependent pair
 * x = fma(x * y, y', z)), i.e. the operand pattern of the column kernels instead of two uniform operands.  The first 256 doubles
 * of `scratch_dev` are the operand seeds (caller-filled, O(1) values), (blocks - 1) * 256 results follow.
 * flops per launch: mode 1: (blocks - 1) * 256 * 8 * iters * 2;  mode 2: ... * 3. */
int cs2_dfma_rate_regs(double* scratch_dev, int32_t blocks, int32_t iters, int32_t mode, void* stream);

/* ---------------------------------------------------------------------------------------
 * Level tables.  Everything that depends on the level only -- `scalm = ZSCAL*max(eta-0.2,
 * ZEPS1)**0.2` (nonlinear/_stencils/cloudsc2.py:127) and the critical relative humidity
 * `crh2(eta[k], trpaus)` (:165-186) for each of the finitely many tropopause candidates
 * trpaus in {0.1} U {eta[j] : 0.1 < eta[j] < 0.4, j <= nlev-2} (:106-111) -- is evaluated
 * once on the host, in the field dtype, and uploaded by the caller.  `eta_host` is the
 * K-field produced by EtaLevels (common/diagnostics.py:42-45), length >= nlev, in `dtype`.
 * ------------------------------------------------------------------------------------- */
size_t cs2_level_tables_bytes(int32_t nlev, int32_t dtype);
/* Fills `tables_host` (>= cs2_level_tables_bytes) ; returns CS2_OK. */
int cs2_level_tables_build(const cs2_params* params, int32_t nlev, int32_t dtype,
                           const void* eta_host, void* tables_host, size_t tables_bytes);

/* ---------------------------------------------------------------------------------------
 * "saturation" stencil -- common/_stencils/saturation.py:23-42, called from
 * Saturation.array_call (common/saturation.py:67-76) over full levels only.
 * ------------------------------------------------------------------------------------- */
int cs2_saturation(const cs2_dims* dims, const cs2_params* params, const void* in_ap,
                   const void* in_t, void* out_qsat, void* stream);

/* ---------------------------------------------------------------------------------------
 * "state_increment" / "perturbed_state" stencils -- common/_stencils/state_increment.py:22-80
 * and perturbed_state.py:22-91, called from common/increment.py:93-132,206-261 over nlev+1
 * levels.  Field order of the 16-pointer arrays (CS2_STATE_*): aph, ap, q, qsat, t, ql, qi,
 * lude, lu, mfu, mfd, tnd_cml_t, tnd_cml_q, tnd_cml_ql, tnd_cml_qi, supsat.
 * ------------------------------------------------------------------------------------- */
#define CS2_NSTATE 16
int cs2_state_increment(const cs2_dims* dims, double f, int32_t ignore_supsat,
                        const void* const in[CS2_NSTATE], void* const out_i[CS2_NSTATE],
                        void* stream);
int cs2_perturbed_state(const cs2_dims* dims, double f, const void* const in[CS2_NSTATE],
                        const void* const in_i[CS2_NSTATE], void* const out[CS2_NSTATE],
                        void* stream);

/* ---------------------------------------------------------------------------------------
 * "cloudsc2_nl" stencil -- nonlinear/_stencils/cloudsc2.py:24-399, called from
 * Cloudsc2NL.array_call (nonlinear/microphysics.py:123-172).  Member names are the stencil's
 * argument names.  The five `tmp_*` IJ scratch fields of the reference are not needed
 * (the carries live in registers).
 * ------------------------------------------------------------------------------------- */
typedef struct cs2_nl_fields {
  const void *in_ap, *in_aph, *in_lu, *in_lude, *in_mfd, *in_mfu, *in_q, *in_qi, *in_ql;
  const void *in_qsat, *in_supsat, *in_t, *in_tnd_cml_q, *in_tnd_cml_qi, *in_tnd_cml_ql;
  const void *in_tnd_cml_t;
  void *out_clc, *out_covptot, *out_fhpsl, *out_fhpsn, *out_fplsl, *out_fplsn;
  void *out_tnd_q, *out_tnd_qi, *out_tnd_ql, *out_tnd_t;
} cs2_nl_fields;

int cs2_nl(const cs2_dims* dims, const cs2_params* params, double dt,
           const void* level_tables_dev, const cs2_nl_fields* f, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused "perturbed_state" + "cloudsc2_nl": NL of the state x + factor * x_i without materialising it --
 * the pair the Taylor test runs once per factor (tangent_linear/validation.py:167-176).  `f` carries
 * the base state x (in_*) and the NL outputs (out_*); of `in_i` only the 16 in_* members (the x_i
 * fields) are read.  Bit-identical to cs2_perturbed_state followed by cs2_nl (same single FMA per
 * input), at 42 instead of 74 field passes through HBM.
 * ------------------------------------------------------------------------------------- */
int cs2_nl_perturbed(const cs2_dims* dims, const cs2_params* params, double dt,
                     const void* level_tables_dev, const cs2_nl_fields* f, const cs2_nl_fields* in_i,
                     double factor, void* stream);

/* ---------------------------------------------------------------------------------------
 * "cloudsc2_tl" stencil -- tangent_linear/_stencils/cloudsc2.py:23-774, called from
 * Cloudsc2TL.array_call (tangent_linear/microphysics.py:162-242).  `traj` holds the NL
 * fields (trajectory inputs and outputs); `pert` holds the `_i` twin of each member.
 * ------------------------------------------------------------------------------------- */
int cs2_tl(const cs2_dims* dims, const cs2_params* params, double dt,
           const void* level_tables_dev, const cs2_nl_fields* traj, const cs2_nl_fields* pert,
           void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused "state_increment" + "cloudsc2_tl": TL with the perturbation of every input formed in registers as
 * factor * input (common/_stencils/state_increment.py:60-80; supsat_i = 0 when ignore_supsat != 0), which is
 * how both validation harnesses build their perturbation (tangent_linear/validation.py:158-162,
 * adjoint/validation.py:136-140).  `traj` as in cs2_tl; of `pert_out` only the 10 out_* members are used.
 * The perturbations are bit-identical to cs2_state_increment's (products rounded on their own); the results equal
 * cs2_state_increment followed by cs2_tl up to FMA contraction (field-scaled difference ~1e-14, within the 1e-12 parity tolerance), at 36 instead of
 * 84 field passes through HBM.
 * norm1_dev (may be NULL): if given, norm1_dev[i] = SUM over levels and the 10 perturbation outputs of (output)^2 for
 * column i in fp64 -- the first inner product of the symmetry test (adjoint/validation.py:167-181), from the sweep itself.
 * ------------------------------------------------------------------------------------- */
int cs2_tl_increment(const cs2_dims* dims, const cs2_params* params, double dt,
                     const void* level_tables_dev, const cs2_nl_fields* traj,
                     const cs2_nl_fields* pert_out, double factor, int32_t ignore_supsat,
                     double* norm1_dev, void* stream);

/* ---------------------------------------------------------------------------------------
 * "cloudsc2_ad" stencil -- adjoint/_stencils/cloudsc2.py:24-996, called from
 * Cloudsc2AD.array_call (adjoint/microphysics.py:159-238).
 *   traj : NL inputs + trajectory outputs (clc, covptot, fluxes, tendencies).
 *   seeds: adjoint seeds, CONSUMED -- the kernel zeroes them like the reference does
 *          (adjoint/_stencils/cloudsc2.py:482-484,506-542,650,714,920,972-984).
 *   adj  : adjoint outputs.
 * `workspace_dev` must hold cs2_ad_workspace_bytes(dims, params, mode) bytes.
 * mode: CS2_AD_RECOMPUTE -- the backward sweep recomputes each level's trajectory from the inputs and
 *       the level-entry fluxes (which are the trajectory outputs fplsl/fplsn); workspace = 4 B/column;
 *       CS2_AD_CHECKPOINT -- the forward sweep also stores the 9 transcendental results of each point
 *       (exp / 1+tanh values, 72 B/point in fp64) to the workspace and the backward sweep replays them
 *       instead of re-evaluating them; the cheap algebra is recomputed in both modes.
 *       Both give the same result up to FMA contraction; measured on B200: DESIGN.md section 3.
 *       With LEVAPLS2 / LDRAIN1D (precipitation-evaporation branch) the sweep is always the recompute one and the
 *       workspace holds one more plane: the overlap carry entering each level (cs2_ad_workspace_bytes accounts for it).
 * ------------------------------------------------------------------------------------- */
enum { CS2_AD_RECOMPUTE = 0, CS2_AD_CHECKPOINT = 1 };

typedef struct cs2_ad_seeds {
  void *in_tnd_t_i, *in_tnd_q_i, *in_tnd_ql_i, *in_tnd_qi_i, *in_clc_i, *in_covptot_i;
  void *in_fhpsl_i, *in_fhpsn_i, *in_fplsl_i, *in_fplsn_i;
} cs2_ad_seeds;

typedef struct cs2_ad_outputs {
  void *out_aph_i, *out_ap_i, *out_q_i, *out_qsat_i, *out_t_i, *out_ql_i, *out_qi_i;
  void *out_lude_i, *out_lu_i, *out_mfu_i, *out_mfd_i, *out_supsat_i;
  void *out_tnd_cml_t_i, *out_tnd_cml_q_i, *out_tnd_cml_ql_i, *out_tnd_cml_qi_i;
} cs2_ad_outputs;

size_t cs2_ad_workspace_bytes(const cs2_dims* dims, const cs2_params* params, int32_t mode);
int cs2_ad(const cs2_dims* dims, const cs2_params* params, double dt,
           const void* level_tables_dev, const cs2_nl_fields* traj, const cs2_ad_seeds* seeds,
           const cs2_ad_outputs* adj, void* workspace_dev, size_t workspace_bytes, int32_t mode,
           void* stream);

/* cs2_ad plus the second inner product of the symmetry test (adjoint/validation.py:183-215) from the backward sweep
 * itself: norm2_dev[i] = SUM over levels and the 16 state fields of (factor * input) * (adjoint output) for column i in
 * fp64, with StateIncrement's roundings (supsat_i = 0 when ignore_supsat != 0) -- the increments need not exist in memory.
 * Not built for LEVAPLS2 / LDRAIN1D (CS2_ERR_UNSUPPORTED: use cs2_ad + cs2_symmetry_norms there). */
int cs2_ad_norm2(const cs2_dims* dims, const cs2_params* params, double dt,
                 const void* level_tables_dev, const cs2_nl_fields* traj, const cs2_ad_seeds* seeds,
                 const cs2_ad_outputs* adj, void* workspace_dev, size_t workspace_bytes, int32_t mode,
                 double factor, int32_t ignore_supsat, double* norm2_dev, void* stream);

/* ---------------------------------------------------------------------------------------
 * One factor of the Taylor test in ONE sweep (tangent_linear/validation.py:158-176,252-261): the NL of the state
 * x + factor2 * (factor1 * x) -- StateIncrement(factor1, ignore_supsat) -> PerturbedState(factor2) -> Cloudsc2NL, with the
 * same roundings as the three stencils -- is evaluated level by level and compared on the fly with the unperturbed NL
 * outputs: sums_dev[2 * n] += SUM over levels and columns of (F_p - F_nl) for the 10 output fields n in the order
 * tnd_t, tnd_q, tnd_ql, tnd_qi, clc, fhpsl, fhpsn, fplsl, fplsn, covptot (the order of TaylorTest.get_norm).  `f` holds the
 * base state x (in_*) and the unperturbed NL outputs F_nl (out_*, READ here); neither the increment, nor the perturbed
 * state, nor F_p ever reach HBM: 26 field passes per factor instead of 62.  Deterministic (fixed-order partial sums in
 * `scratch_dev`, cs2_taylor_nl_scratch_bytes(dims) bytes).
 * ------------------------------------------------------------------------------------- */
size_t cs2_taylor_nl_scratch_bytes(const cs2_dims* dims);
int cs2_taylor_nl_sums(const cs2_dims* dims, const cs2_params* params, double dt,
                       const void* level_tables_dev, const cs2_nl_fields* f, double factor1,
                       int32_t ignore_supsat, double factor2, double* sums_dev, void* scratch_dev,
                       size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Validation reductions (device-side replacements of the host NumPy sums of the reference).
 *
 * cs2_taylor_sums -- TaylorTest.get_field_norm (tangent_linear/validation.py:252-261):
 *   for each of `nfields` fields, accumulates into sums_dev[2*f+0] += SUM(a_f - b_f) and
 *   sums_dev[2*f+1] += SUM(c_f) over all nlev+1 levels and ncol columns (fp64 accumulation,
 *   deterministic two-stage reduction; `b`/`c` entries may be NULL to skip that term).
 *   sums_dev: 2*nfields doubles, NOT zeroed by the call (so shards can be accumulated);
 *   scratch_dev: cs2_taylor_scratch_bytes(dims, nfields) bytes.
 * cs2_symmetry_norms -- SymmetryTest.get_norm1/get_norm2 (adjoint/validation.py:167-215):
 *   norm_dev[i] = SUM_k SUM_f a_f[k,i] * b_f[k,i] per column (fp64), up to 32 field pairs per call (norm2 has 16);
 *   a pair with a_f == b_f is read once.
 * cs2_symmetry_residual -- SymmetryTest.__call__ (adjoint/validation.py:157-165):
 *   norm3_dev[i] = |n1 - n2| / eps where n2 == 0, |n1 - n2| / (eps * n2) elsewhere (norm3_dev may be NULL), and
 *   max_dev[0] = max_i norm3[i] (NaN if any norm3 is NaN, -inf for ncol == 0) -- the one double a sharded run
 *   all-reduces (MAX).  scratch_dev: cs2_symmetry_residual_scratch_bytes(ncol) bytes.  Deterministic.
 * ------------------------------------------------------------------------------------- */
size_t cs2_taylor_scratch_bytes(const cs2_dims* dims, int32_t nfields);
int cs2_taylor_sums(const cs2_dims* dims, int32_t nfields, const void* const* a_dev,
                    const void* const* b_dev, const void* const* c_dev, double* sums_dev,
                    void* scratch_dev, size_t scratch_bytes, void* stream);
int cs2_symmetry_norms(const cs2_dims* dims, int32_t nfields, const void* const* a_dev,
                       const void* const* b_dev, double* norm_dev, void* stream);
size_t cs2_symmetry_residual_scratch_bytes(int64_t ncol);
int cs2_symmetry_residual(int64_t ncol, const double* norm1_dev, const double* norm2_dev, double eps,
                          double* norm3_dev, double* max_dev, void* scratch_dev, size_t scratch_bytes,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLOUDSC2_B200_H */
