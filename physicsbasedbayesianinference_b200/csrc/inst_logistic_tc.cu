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
  const size_t sm = (size_t)pa.DP * LT_M * 2 + 2 * (size_t)pa.chunk_bytes + LT_NB * LT_M * 2 + 2 * LT_M * 4 + 12 * 8 + 16;
  if (sm > 227 * 1024) return fail(EHMC_ERR_UNSUPPORTED, "logistic tensor-core kernel needs %zu B shared memory", sm);
  const unsigned grid = (unsigned)((P + LT_M - 1) / LT_M);
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
