#include <algorithm>

#include "host_defs.h"
#include "k_logistic_tc.cuh"
#include "k_logistic_tcs.cuh"

namespace ehmc {

int logistic_grad_tc(ehmc_ctx* c, const ehmc_potential* p, const float* theta, long long t_ld, long long P, float* g,
                     long long g_ld, float* e, double* e64, cudaStream_t st) {
  LogisticTcArgs pa;
  pa.chunks = static_cast<const unsigned char*>(p->d6);
  pa.NC = p->lt_nc;
  pa.DP = p->lt_dp;
  pa.D = p->D;
  pa.n_pad = p->lt_npad;
  pa.chunk_bytes = p->lt_chunk_bytes;
  pa.inv_s2 = (float)(1.0 / (p->scalars[0] * p->scalars[0]));
  const size_t fixed = 2 * LT_M * 8 + (2 * LT_MAX_STAGES + 5) * 8 + 16;
  pa.stages = (int)std::min<size_t>(LT_MAX_STAGES, (227 * 1024 - fixed) / pa.chunk_bytes);
  if (pa.stages < 2) return fail(EHMC_ERR_UNSUPPORTED, "logistic tensor-core kernel: chunk of %u B does not fit twice", pa.chunk_bytes);
  const size_t sm = (size_t)pa.stages * pa.chunk_bytes + fixed;
  // wave quantisation: T tiles on S SMs take ceil(T / S) rounds; splitting every tile over two halves of the
  // data rows doubles the units of work (config 3: 512 tiles, 148 SMs: 4 rounds for 3.46 -> 7 for 6.92)
  const long long tiles = (P + LT_M - 1) / LT_M;
  const int sms = c->prop.multiProcessorCount;
  auto waste = [&](long long units) { return (double)((units + sms - 1) / sms) * sms / (double)units; };
  pa.split = (pa.NC >= 2 && waste(2 * tiles) + 0.03 < waste(tiles)) ? 2 : 1;
  if (pa.split == 2) {
    if (g != nullptr) CUDA_TRY(cudaMemset2DAsync(g, (size_t)g_ld * sizeof(float), 0, (size_t)P * sizeof(float), (size_t)p->D, st));
    if (e != nullptr) CUDA_TRY(cudaMemsetAsync(e, 0, (size_t)P * sizeof(float), st));
    if (e64 != nullptr) CUDA_TRY(cudaMemsetAsync(e64, 0, (size_t)P * sizeof(double), st));
  }
  const unsigned grid = (unsigned)(tiles * pa.split);
  if (e != nullptr || e64 != nullptr) {
    CUDA_TRY(cudaFuncSetAttribute(k_logistic_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_logistic_tc<true><<<grid, LT_THREADS, sm, st>>>(theta, t_ld, P, g, g_ld, e, e64, pa);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(k_logistic_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_logistic_tc<false><<<grid, LT_THREADS, sm, st>>>(theta, t_ld, P, g, g_ld, e, e64, pa);
  }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return EHMC_OK;
}

// float32-accurate variant: 3-pass fp16 split of both GEMMs (k_logistic_tcs.cuh)
int logistic_grad_tcs(ehmc_ctx* c, const ehmc_potential* p, const float* theta, long long t_ld, long long P, float* g,
                      long long g_ld, float* e, double* e64, cudaStream_t st) {
  LogisticTcsArgs pa;
  pa.entries = static_cast<const unsigned char*>(p->d6);
  pa.NC = p->lt_nc;
  pa.DP = p->lt_dp;
  pa.D = p->D;
  pa.n_pad = p->lt_npad;
  pa.entry_bytes = p->lt_chunk_bytes;
  pa.inv_s2 = (float)(1.0 / (p->scalars[0] * p->scalars[0]));
  pa.x_iscale = p->lts_x_iscale;
  const size_t fixed = (size_t)pa.DP * LT_M * 2 + (2 * LT_MAX_STAGES + 6) * 8 + 16;  // Theta_lo + barriers + TMEM slot
  pa.stages = (int)std::min<size_t>(LT_MAX_STAGES, (227 * 1024 - fixed) / pa.entry_bytes);
  if (pa.stages < LTS_MIN_STAGES)
    return fail(EHMC_ERR_UNSUPPORTED, "logistic tensor-core kernel (fp16 split): %d ring entries of %u B fit, %d needed",
                pa.stages, pa.entry_bytes, LTS_MIN_STAGES);
  const size_t sm = (size_t)pa.stages * pa.entry_bytes + fixed;
  const long long tiles = (P + LT_M - 1) / LT_M;
  const int sms = c->prop.multiProcessorCount;
  auto waste = [&](long long units) { return (double)((units + sms - 1) / sms) * sms / (double)units; };
  pa.split = (pa.NC >= 2 && waste(2 * tiles) + 0.03 < waste(tiles)) ? 2 : 1;
  if (pa.split == 2) {
    if (g != nullptr) CUDA_TRY(cudaMemset2DAsync(g, (size_t)g_ld * sizeof(float), 0, (size_t)P * sizeof(float), (size_t)p->D, st));
    if (e != nullptr) CUDA_TRY(cudaMemsetAsync(e, 0, (size_t)P * sizeof(float), st));
    if (e64 != nullptr) CUDA_TRY(cudaMemsetAsync(e64, 0, (size_t)P * sizeof(double), st));
  }
  const unsigned grid = (unsigned)(tiles * pa.split);
  if (e != nullptr || e64 != nullptr) {
    CUDA_TRY(cudaFuncSetAttribute(k_logistic_tcs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_logistic_tcs<true><<<grid, LTS_THREADS, sm, st>>>(theta, t_ld, P, g, g_ld, e, e64, pa);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(k_logistic_tcs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_logistic_tcs<false><<<grid, LTS_THREADS, sm, st>>>(theta, t_ld, P, g, g_ld, e, e64, pa);
  }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return EHMC_OK;
}

}  // namespace ehmc
