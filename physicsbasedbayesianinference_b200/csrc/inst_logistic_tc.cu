#include <algorithm>

#include "host_defs.h"
#include "k_logistic_tc.cuh"

namespace ehmc {

int logistic_grad_tc(ehmc_ctx* c, const ehmc_potential* p, const float* theta, long long t_ld, long long P, float* g,
                     long long g_ld, float* e, cudaStream_t st) {
  LogisticTcArgs pa;
  pa.chunks = static_cast<const unsigned char*>(p->d6);
  pa.NC = p->lt_nc;
  pa.DP = p->lt_dp;
  pa.D = p->D;
  pa.n_pad = p->lt_npad;
  pa.chunk_bytes = p->lt_chunk_bytes;
  pa.inv_s2 = (float)(1.0 / (p->scalars[0] * p->scalars[0]));
  const size_t fixed = 2 * LT_M * 4 + (2 * LT_MAX_STAGES + 5) * 8 + 16;
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
  }
  const unsigned grid = (unsigned)(tiles * pa.split);
  if (e != nullptr) {
    CUDA_TRY(cudaFuncSetAttribute(k_logistic_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_logistic_tc<true><<<grid, LT_THREADS, sm, st>>>(theta, t_ld, P, g, g_ld, e, pa);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(k_logistic_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_logistic_tc<false><<<grid, LT_THREADS, sm, st>>>(theta, t_ld, P, g, g_ld, e, pa);
  }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return EHMC_OK;
}

}  // namespace ehmc
