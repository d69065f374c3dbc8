// Host "twin" of the CUDA column sweeps (TEST INFRASTRUCTURE / CPU-baseline arm only).
//
// Compiles the product's column code (csrc/cs2_columns.cuh -- the same templates the sm_100a
// kernels instantiate) with g++ and OpenMP, one loop iteration per column, on HOST arrays in
// the product layout `[nlev+1][ncol_stride]`.  Two uses, neither of them in the product path:
//   1. `tests/ -m "not gpu"`: lets the CPU-only suite check the kernel math (NL, TL, AD)
//      against the independent NumPy oracle (oracle/cloudsc2_numpy.py) without a GPU;
//   2. `bench.py`: the multi-threaded C++ CPU port timed beside the GPU numbers
//      (cpu_baseline kind "port").
// It is NOT the parity oracle (it shares source with the kernels); the oracle is the NumPy
// restatement of the reference stencils.
#include <cstdint>
#include <cstring>

#include "../../gt4py-dwarf-p-cloudsc2-tl-ad_b200/csrc/cs2_columns.cuh"
// the level functions of the kernel experiments stay under test here although the shipped library does not contain them
#include "../../gt4py-dwarf-p-cloudsc2-tl-ad_b200/csrc/experiments/cs2_experiment_columns.cuh"

namespace {
template <class R>
void run_nl(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f) {
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, dt);
  const cs2::NLFields<R> nf = cs2::make_nl_fields<R>(*f);
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  const bool evap = P->LEVAPLS2 || P->LDRAIN1D, tetens = P->LPHYLIN || P->LDRAIN1D;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < d->ncol; ++i) {
    if (evap && tetens) cs2::column_nl<R, cs2::Cfg<true, true>>(p, tab, nf, d->ncol_stride, d->nlev, i, false, nullptr);
    else if (evap) cs2::column_nl<R, cs2::Cfg<true, false>>(p, tab, nf, d->ncol_stride, d->nlev, i, false, nullptr);
    else if (tetens) cs2::column_nl<R, cs2::Cfg<false, true>>(p, tab, nf, d->ncol_stride, d->nlev, i, false, nullptr);
    else cs2::column_nl<R, cs2::Cfg<false, false>>(p, tab, nf, d->ncol_stride, d->nlev, i, false, nullptr);
  }
}

template <class R>
void run_nl_split(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f) {
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, dt);
  const cs2::NLFields<R> nf = cs2::make_nl_fields<R>(*f);
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  const bool tetens = P->LPHYLIN || P->LDRAIN1D;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < d->ncol; ++i) {
    if (tetens) cs2::column_nl_split<R, cs2::Cfg<false, true>>(p, tab, nf, d->ncol_stride, d->nlev, i);
    else cs2::column_nl_split<R, cs2::Cfg<false, false>>(p, tab, nf, d->ncol_stride, d->nlev, i);
  }
}

template <class R>
void run_nl_pipe(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f, bool ad_ref) {
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, dt);
  const cs2::NLFields<R> nf = cs2::make_nl_fields<R>(*f);
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < d->ncol; ++i) cs2::column_nl_pipe<R>(p, tab, nf, d->ncol_stride, d->nlev, i, ad_ref, nullptr);
}

template <class R>
void run_tl(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f,
            const cs2_nl_fields* g) {
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, dt);
  const cs2::NLFields<R> nf = cs2::make_nl_fields<R>(*f), ng = cs2::make_nl_fields<R>(*g);
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  const bool evap = P->LEVAPLS2 || P->LDRAIN1D;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < d->ncol; ++i) {
    if (evap) cs2::column_tl<R, true>(p, tab, nf, ng, d->ncol_stride, d->nlev, i);
    else cs2::column_tl<R, false>(p, tab, nf, ng, d->ncol_stride, d->nlev, i);
  }
}

template <class R>
void run_ad(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f,
            const cs2_ad_seeds* sd, const cs2_ad_outputs* ao, int32_t* jsel, void* cov_ws) {
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, dt);
  const cs2::NLFields<R> nf = cs2::make_nl_fields<R>(*f);
  const cs2::LevelTables<R> tab = cs2::view_tables<R>(tables);
  cs2::ADSeeds<R> s;
  s.tnd_t = (R*)sd->in_tnd_t_i; s.tnd_q = (R*)sd->in_tnd_q_i; s.tnd_ql = (R*)sd->in_tnd_ql_i;
  s.tnd_qi = (R*)sd->in_tnd_qi_i; s.clc = (R*)sd->in_clc_i; s.covptot = (R*)sd->in_covptot_i;
  s.fhpsl = (R*)sd->in_fhpsl_i; s.fhpsn = (R*)sd->in_fhpsn_i; s.fplsl = (R*)sd->in_fplsl_i;
  s.fplsn = (R*)sd->in_fplsn_i;
  cs2::ADOut<R> a;
  a.aph = (R*)ao->out_aph_i; a.ap = (R*)ao->out_ap_i; a.q = (R*)ao->out_q_i; a.qsat = (R*)ao->out_qsat_i;
  a.t = (R*)ao->out_t_i; a.ql = (R*)ao->out_ql_i; a.qi = (R*)ao->out_qi_i; a.lude = (R*)ao->out_lude_i;
  a.lu = (R*)ao->out_lu_i; a.mfu = (R*)ao->out_mfu_i; a.mfd = (R*)ao->out_mfd_i; a.supsat = (R*)ao->out_supsat_i;
  a.tnd_t = (R*)ao->out_tnd_cml_t_i; a.tnd_q = (R*)ao->out_tnd_cml_q_i; a.tnd_ql = (R*)ao->out_tnd_cml_ql_i;
  a.tnd_qi = (R*)ao->out_tnd_cml_qi_i;
  const bool ad_ref = !P->AD_TL_PREDICATES;
  const bool evap = P->LEVAPLS2 || P->LDRAIN1D;
  R* cov = static_cast<R*>(cov_ws);  // [nlev][ncol_stride], evaporation branch only
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < d->ncol; ++i) {
    if (evap) {
      cs2::column_nl<R, cs2::Cfg<true, true>, true>(p, tab, nf, d->ncol_stride, d->nlev, i, ad_ref, jsel, cov);
      cs2::column_ad_bwd<R, true>(p, tab, nf, s, a, jsel, d->ncol_stride, d->nlev, i, cov);
    } else {
      cs2::column_nl<R, cs2::Cfg<false, true>, true>(p, tab, nf, d->ncol_stride, d->nlev, i, ad_ref, jsel);
      cs2::column_ad_bwd<R, false>(p, tab, nf, s, a, jsel, d->ncol_stride, d->nlev, i);
    }
  }
}

template <class R>
void run_sat(const cs2_dims* d, const cs2_params* P, const void* ap, const void* t, void* qsat) {
  const cs2::DevParams<R> p = cs2::make_dev_params<R>(*P, 1.0);
  const R* a = (const R*)ap;
  const R* tt = (const R*)t;
  R* q = (R*)qsat;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < d->nlev; ++k) {
    int64_t i = 0;
    for (; i + 3 < d->ncol; i += 4) {  // four points at once, as a thread of saturation_kernel does
      const int64_t o = int64_t(k) * d->ncol_stride + i;
      const R av[4] = {a[o], a[o + 1], a[o + 2], a[o + 3]}, tv[4] = {tt[o], tt[o + 1], tt[o + 2], tt[o + 3]};
      R qv[4];
      cs2::saturation_points<R, 4>(p, P->LPHYLIN != 0, av, tv, qv);
      for (int j = 0; j < 4; ++j) q[o + j] = qv[j];
    }
    for (; i < d->ncol; ++i) {
      const int64_t o = int64_t(k) * d->ncol_stride + i;
      q[o] = cs2::saturation_point<R>(p, P->LPHYLIN != 0, a[o], tt[o]);
    }
  }
}
}  // namespace

extern "C" {
int twin_saturation(const cs2_dims* d, const cs2_params* P, const void* ap, const void* t, void* qsat) {
  if (d->dtype == CS2_F64) run_sat<double>(d, P, ap, t, qsat); else run_sat<float>(d, P, ap, t, qsat);
  return 0;
}
int twin_nl(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f) {
  if (d->dtype == CS2_F64) run_nl<double>(d, P, dt, tables, f); else run_nl<float>(d, P, dt, tables, f);
  return 0;
}
// the two half-level functions of the split NL kernel, evaluated back to back (evaporation branch off only)
int twin_nl_split(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f) {
  if (P->LEVAPLS2 || P->LDRAIN1D) return 1;
  if (d->dtype == CS2_F64) run_nl_split<double>(d, P, dt, tables, f); else run_nl_split<float>(d, P, dt, tables, f);
  return 0;
}
// the software-pipelined level function of the default-flag NL kernel (cs2_physics_pipe.cuh); default flags only
int twin_nl_pipe(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f, int ad_ref) {
  if (P->LEVAPLS2 || P->LDRAIN1D || !P->LPHYLIN || P->RVTMP2 != 0.0) return 1;
  if (d->dtype == CS2_F64) run_nl_pipe<double>(d, P, dt, tables, f, ad_ref != 0);
  else run_nl_pipe<float>(d, P, dt, tables, f, ad_ref != 0);
  return 0;
}
int twin_tl(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f,
            const cs2_nl_fields* g) {
  if (d->dtype == CS2_F64) run_tl<double>(d, P, dt, tables, f, g); else run_tl<float>(d, P, dt, tables, f, g);
  return 0;
}
// cov_ws: (nlev * ncol_stride) elements of scratch, only read/written with LEVAPLS2 / LDRAIN1D (may be NULL otherwise)
int twin_ad(const cs2_dims* d, const cs2_params* P, double dt, const void* tables, const cs2_nl_fields* f,
            const cs2_ad_seeds* s, const cs2_ad_outputs* a, int32_t* jsel, void* cov_ws) {
  if ((P->LEVAPLS2 || P->LDRAIN1D) && !cov_ws) return 1;
  if (d->dtype == CS2_F64) run_ad<double>(d, P, dt, tables, f, s, a, jsel, cov_ws);
  else run_ad<float>(d, P, dt, tables, f, s, a, jsel, cov_ws);
  return 0;
}
}
