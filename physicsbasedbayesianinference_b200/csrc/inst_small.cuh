// Instantiations of the small-D fused kernels for one dtype (included by inst_small_f32/f64.cu).
#pragma once

#include "host_defs.h"
#include "k_misc.cuh"
#include "k_small.cuh"

namespace ehmc {

template <typename T, int DT>
static DiagPot<T, DT> make_diag(const ehmc_potential* p) {
  DiagPot<T, DT> f;
  for (int d = 0; d < DT; ++d) f.k[d] = d < p->D ? (T)p->hp0[d] : T(0);
  return f;
}
template <typename T, int DT>
static DenseSmallPot<T, DT> make_dense_small(const ehmc_potential* p) {
  DenseSmallPot<T, DT> f;
  for (int i = 0; i < DT; ++i) {
    f.mu[i] = i < p->D ? (T)p->hp1[i] : T(0);
    for (int j = 0; j < DT; ++j) f.lam[i * DT + j] = (i < p->D && j < p->D) ? (T)p->hp0[(size_t)i * p->D + j] : T(0);
  }
  return f;
}
template <typename T, int DT>
static FunnelPot<T, DT> make_funnel(const ehmc_potential* p) {
  FunnelPot<T, DT> f;
  f.inv_s2 = (T)(1.0 / (p->scalars[1] * p->scalars[1]));
  f.half_dm1 = (T)(0.5 * (p->D - 1));
  return f;
}

template <typename T, int DT, class Pot>
static int launch_small_pot(ehmc_ctx* c, const IterArgs<T>& A, const Pot& pot, int integ, bool hmc, cudaStream_t st) {
  static_assert(K1_THREADS == K1_THREADS_HOST, "K1 block size");
  const unsigned grid = (unsigned)((A.P + K1_THREADS - 1) / K1_THREADS);
  const size_t sm = A.partials != nullptr ? sizeof(double) * (K1_THREADS / 32) * (2 * A.D + 3) : 0;
  if (hmc) {
    if (integ == INTEG_LEAPFROG)
      k_small<T, DT, Pot, INTEG_LEAPFROG, true><<<grid, K1_THREADS, sm, st>>>(A, pot);
    else
      k_small<T, DT, Pot, INTEG_STORMER, true><<<grid, K1_THREADS, sm, st>>>(A, pot);
  } else {
    if (integ == INTEG_LEAPFROG)
      k_small<T, DT, Pot, INTEG_LEAPFROG, false><<<grid, K1_THREADS, sm, st>>>(A, pot);
    else
      k_small<T, DT, Pot, INTEG_STORMER, false><<<grid, K1_THREADS, sm, st>>>(A, pot);
  }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return EHMC_OK;
}

template <typename T, int DT>
static int launch_small_dt(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, int integ, bool hmc,
                           cudaStream_t st) {
  switch (p->family) {
    case EHMC_FAMILY_DIAG_GAUSSIAN:
      return launch_small_pot<T, DT>(c, A, make_diag<T, DT>(p), integ, hmc, st);
    case EHMC_FAMILY_FUNNEL:
      return launch_small_pot<T, DT>(c, A, make_funnel<T, DT>(p), integ, hmc, st);
    case EHMC_FAMILY_DENSE_GAUSSIAN:
      if constexpr (DT <= 16) return launch_small_pot<T, DT>(c, A, make_dense_small<T, DT>(p), integ, hmc, st);
      break;
    default:
      break;
  }
  return fail(EHMC_ERR_UNSUPPORTED, "family %d has no small-D kernel for D = %d", p->family, p->D);
}

template <typename T>
int launch_small(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, int integ, bool hmc, cudaStream_t st) {
  const int D = p->D;
  if (D <= 2) return launch_small_dt<T, 2>(c, p, A, integ, hmc, st);
  if (D <= 4) return launch_small_dt<T, 4>(c, p, A, integ, hmc, st);
  if (D <= 8) return launch_small_dt<T, 8>(c, p, A, integ, hmc, st);
  if (D <= 10) return launch_small_dt<T, 10>(c, p, A, integ, hmc, st);
  if (D <= 16) return launch_small_dt<T, 16>(c, p, A, integ, hmc, st);
  if (D <= 32) return launch_small_dt<T, 32>(c, p, A, integ, hmc, st);
  return fail(EHMC_ERR_UNSUPPORTED, "small-D kernel: D = %d > 32", D);
}

template <typename T, int DT>
static int eval_small_dt(ehmc_ctx* c, const ehmc_potential* p, const T* q, long long q_ld, long long P, T* e, T* g,
                         long long g_ld, cudaStream_t st) {
  const unsigned grid = (unsigned)((P + 127) / 128);
  switch (p->family) {
    case EHMC_FAMILY_DIAG_GAUSSIAN:
      k_eval_small<T, DT><<<grid, 128, 0, st>>>(q, q_ld, P, p->D, e, g, g_ld, make_diag<T, DT>(p));
      break;
    case EHMC_FAMILY_FUNNEL:
      k_eval_small<T, DT><<<grid, 128, 0, st>>>(q, q_ld, P, p->D, e, g, g_ld, make_funnel<T, DT>(p));
      break;
    default:
      return fail(EHMC_ERR_UNSUPPORTED, "eval: family %d", p->family);
  }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return EHMC_OK;
}

template <typename T>
int eval_small(ehmc_ctx* c, const ehmc_potential* p, const T* q, long long q_ld, long long P, T* e, T* g,
               long long g_ld, cudaStream_t st) {
  const int D = p->D;
  if (D <= 2) return eval_small_dt<T, 2>(c, p, q, q_ld, P, e, g, g_ld, st);
  if (D <= 4) return eval_small_dt<T, 4>(c, p, q, q_ld, P, e, g, g_ld, st);
  if (D <= 8) return eval_small_dt<T, 8>(c, p, q, q_ld, P, e, g, g_ld, st);
  if (D <= 10) return eval_small_dt<T, 10>(c, p, q, q_ld, P, e, g, g_ld, st);
  if (D <= 16) return eval_small_dt<T, 16>(c, p, q, q_ld, P, e, g, g_ld, st);
  if (D <= 32) return eval_small_dt<T, 32>(c, p, q, q_ld, P, e, g, g_ld, st);
  return fail(EHMC_ERR_UNSUPPORTED, "eval: D = %d", D);
}

}  // namespace ehmc
