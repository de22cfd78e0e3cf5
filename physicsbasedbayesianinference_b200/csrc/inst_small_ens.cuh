// Instantiations of the fused adaptive ensemble run (k_small_ens) for one dtype (included by inst_small_ens_f32/f64.cu).
#pragma once

#include <algorithm>

#include "inst_small.cuh"
#include "k_small_ens.cuh"

namespace ehmc {

template <typename T, int DT, class Pot, bool EXACT>
static int ens_launch(ehmc_ctx* c, const IterArgs<T>& A, const Pot& pot, EnsRunArgs<T> R, cudaStream_t st) {
  auto kernel = k_small_ens<T, DT, Pot, INTEG_LEAPFROG, EXACT>;
  const size_t sm = sizeof(double) * (K1_THREADS + K1_THREADS / 32) * (2 * DT + 3);
  if (sm > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int occ = 0, coop = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, K1_THREADS, sm));
  CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
  if (!coop || occ < 1) return fail(EHMC_ERR_UNSUPPORTED, "fused ensemble run: cooperative launch unavailable");
  // one resident wave: compute CTAs plus the service CTAs
  const long long need = (A.P + K1_THREADS - 1) / K1_THREADS;
  const long long cap = (long long)occ * c->prop.multiProcessorCount;
  if (cap <= ENS_SERVICE_CTAS) return fail(EHMC_ERR_UNSUPPORTED, "fused ensemble run: device too small");
  const unsigned ncompute = (unsigned)std::max<long long>(1, std::min<long long>(need, cap - ENS_SERVICE_CTAS));
  const int NS = 2 * A.D + 3;
  // control block: hsched [nIter + 2] | published | ticket [2] | group tickets [2][ngroups] | rows [2][ncompute][NS] |
  // group rows [2][ngroups][NS] | gains [adaptIters]
  const size_t n_h = ((size_t)R.nIter + 2 + 15) / 16 * 16;  // (every section starts on a 128-byte line)
  const size_t n_pub = (size_t)ENS_PUB_COPIES * 16;
  const size_t n_tk = 32 + (size_t)ENS_PUB_COPIES * 16;  // ticket[0], ticket[1] on lines of their own | arrival counters
  const unsigned ngroups = (ncompute + ENS_GROUP - 1) / ENS_GROUP;
  const size_t n_gt = ((size_t)2 * ngroups / 2 + 16) / 16 * 16;  // group tickets (unsigned) in units of doubles
  const size_t n_rows = ((size_t)2 * ncompute * NS + 15) / 16 * 16;
  const size_t n_grows = ((size_t)2 * ngroups * NS + 15) / 16 * 16;
  const size_t bytes = sizeof(double) * (n_h + n_pub + n_tk + n_gt + n_rows + n_grows + (size_t)std::max(1, R.adaptIters)) + 128;
  TRY(c->ens_ctl.ensure(bytes));
  double* base = reinterpret_cast<double*>(((uintptr_t)c->ens_ctl.ptr + 127) & ~(uintptr_t)127);
  R.hsched = base;
  R.published = reinterpret_cast<long long*>(base + n_h);
  R.ticket = reinterpret_cast<unsigned*>(base + n_h + n_pub);  // ticket[par] at R.ticket[32 * par] (see the kernel)
  R.arrived = reinterpret_cast<unsigned long long*>(base + n_h + n_pub + 32);
  R.lockstep = c->ens_lockstep;
  R.gticket = reinterpret_cast<unsigned*>(base + n_h + n_pub + n_tk);
  R.rows = base + n_h + n_pub + n_tk + n_gt;
  R.grows = R.rows + n_rows;
  double* gains = R.grows + n_grows;
  CUDA_TRY(cudaMemsetAsync(base + n_h, 0, (n_pub + n_tk + n_gt) * sizeof(double), st));  // published = 0, all tickets = 0
  if (R.adaptIters > 0)  // R.gains arrives as a HOST array
    CUDA_TRY(cudaMemcpyAsync(gains, R.gains, sizeof(double) * R.adaptIters, cudaMemcpyHostToDevice, st));
  R.gains = gains;
  R.dbg = nullptr;
  R.dbg_iters = 0;
  if (c->ens_debug > 0) {
    TRY(c->ens_dbg_buf.ensure(sizeof(long long) * 8 * (size_t)c->ens_debug));
    CUDA_TRY(cudaMemsetAsync(c->ens_dbg_buf.ptr, 0, sizeof(long long) * 8 * (size_t)c->ens_debug, st));
    R.dbg = static_cast<long long*>(c->ens_dbg_buf.ptr);
    R.dbg_iters = c->ens_debug;
  }
  IterArgs<T> Ac = A;
  void* args[] = {(void*)&Ac, (void*)&pot, (void*)&R};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kernel, dim3(ncompute + ENS_SERVICE_CTAS), dim3(K1_THREADS), args, sm, st));
  c->launches++;
  return EHMC_OK;
}

template <typename T, int DT, class Pot>
static int ens_pot(ehmc_ctx* c, const IterArgs<T>& A, const Pot& pot, const EnsRunArgs<T>& R, cudaStream_t st) {
  return A.D == DT ? ens_launch<T, DT, Pot, true>(c, A, pot, R, st) : ens_launch<T, DT, Pot, false>(c, A, pot, R, st);
}

template <typename T, int DT>
static int ens_dt(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, const EnsRunArgs<T>& R, cudaStream_t st) {
  switch (p->family) {
    case EHMC_FAMILY_DIAG_GAUSSIAN: return ens_pot<T, DT>(c, A, make_diag<T, DT>(p), R, st);
    case EHMC_FAMILY_FUNNEL: return ens_pot<T, DT>(c, A, make_funnel<T, DT>(p), R, st);
    case EHMC_FAMILY_COIN_TOSS: return ens_pot<T, DT>(c, A, make_coin<T, DT>(p), R, st);
    case EHMC_FAMILY_DENSE_GAUSSIAN:
      if constexpr (DT <= 16) return ens_pot<T, DT>(c, A, make_dense_small<T, DT>(p), R, st);
      break;
    default: break;
  }
  return fail(EHMC_ERR_UNSUPPORTED, "family %d has no fused ensemble-run kernel for D = %d", p->family, p->D);
}

template <typename T>
int run_small_ens(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, const EnsRunArgs<T>& R, cudaStream_t st) {
  const int D = p->D;
  if (D <= 2) return ens_dt<T, 2>(c, p, A, R, st);
  if (D <= 4) return ens_dt<T, 4>(c, p, A, R, st);
  if (D <= 8) return ens_dt<T, 8>(c, p, A, R, st);
  if (D <= 10) return ens_dt<T, 10>(c, p, A, R, st);
  if (D <= 16) return ens_dt<T, 16>(c, p, A, R, st);
  if (D <= 32) return ens_dt<T, 32>(c, p, A, R, st);
  return fail(EHMC_ERR_UNSUPPORTED, "small-D kernel: D = %d > 32", D);
}

}  // namespace ehmc
