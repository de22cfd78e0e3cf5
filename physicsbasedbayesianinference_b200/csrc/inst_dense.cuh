// Instantiations of the dense-Gaussian fused kernel for one dtype.
#pragma once

#include "host_defs.h"
#include "k_dense.cuh"

namespace ehmc {

template <typename T, int TN, int MINB>
static int launch_dense_tn(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, int integ, bool hmc,
                           cudaStream_t st) {
  typedef DenseShape<T, TN> S;
  const size_t sm = S::smem_bytes(A.D);
  if (sm > 227 * 1024) return fail(EHMC_ERR_UNSUPPORTED, "dense kernel needs %zu B shared memory", sm);
  CUDA_TRY(cudaFuncSetAttribute(k_dense<T, TN, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  DenseArgs<T> pa;
  pa.Ls = static_cast<const T*>(p->d0);
  pa.mu = static_cast<const T*>(p->d1);
  const unsigned grid = (unsigned)((A.P + S::PT - 1) / S::PT);
  k_dense<T, TN, MINB><<<grid, K2_THREADS, sm, st>>>(A, pa, integ, hmc ? 1 : 0);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return EHMC_OK;
}

template <typename T>
int launch_dense(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, int integ, bool hmc, cudaStream_t st) {
  const bool two = sizeof(T) == 4 && c->dense_occupancy >= 2;
  switch (p->TN) {
    case 4: return launch_dense_tn<T, 4, 1>(c, p, A, integ, hmc, st);
    case 8: return launch_dense_tn<T, 8, 1>(c, p, A, integ, hmc, st);
    case 13:
      if constexpr (sizeof(T) == 4) {
        if (two) return launch_dense_tn<T, 13, 2>(c, p, A, integ, hmc, st);
      }
      return launch_dense_tn<T, 13, 1>(c, p, A, integ, hmc, st);
    case 16: return launch_dense_tn<T, 16, 1>(c, p, A, integ, hmc, st);
  }
  return fail(EHMC_ERR_INVALID, "dense potential not packed (TN = %d)", p->TN);
}

template <typename T>
int dense_particles_per_cta() {
  return 32 * DenseTile<T>::TM;
}

template <typename T>
int dense_tnp(int TN) {
  const int VW = 16 / (int)sizeof(T);
  return (TN + VW - 1) / VW * VW;
}

}  // namespace ehmc
