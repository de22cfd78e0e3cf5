#include <algorithm>

#include "host_defs.h"
#include "k_dense_tc.cuh"
#include "k_dense_tc2.cuh"

namespace ehmc {

template <int K8>
static int launch_tc_k8(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<float>& A, bool hmc, cudaStream_t st) {
  constexpr size_t sm = TcShape<K8>::smem_bytes();
  static_assert(sm <= 227 * 1024, "dense tensor-core kernel: shared memory");
  CUDA_TRY(cudaFuncSetAttribute(k_dense_tc<K8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  DenseTcArgs pa;
  pa.Bhi = static_cast<const float*>(p->d3);
  pa.Blo = static_cast<const float*>(p->d4);
  pa.mu = static_cast<const float*>(p->d5);
  pa.dbg = c->tc_debug;
  pa.prof = c->tc_prof ? static_cast<long long*>(c->tc_prof_buf.ptr) : nullptr;
  const unsigned grid = (unsigned)((A.P + TC_M - 1) / TC_M);
  k_dense_tc<K8><<<grid, TC_THREADS, sm, st>>>(A, pa, hmc ? 1 : 0);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return EHMC_OK;
}

template <int K8>
static int launch_tc2_k8(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<float>& A, bool hmc, cudaStream_t st) {
  constexpr size_t sm = Tc2Shape<K8>::smem_bytes();
  static_assert(sm <= 227 * 1024, "dense tensor-core kernel (2 tiles): shared memory");
  CUDA_TRY(cudaFuncSetAttribute(k_dense_tc2<K8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  DenseTcArgs pa;
  pa.Bhi = static_cast<const float*>(p->d3);
  pa.Blo = static_cast<const float*>(p->d4);
  pa.mu = static_cast<const float*>(p->d5);
  pa.dbg = c->tc_debug;
  pa.prof = c->tc_prof ? static_cast<long long*>(c->tc_prof_buf.ptr) : nullptr;
  const unsigned grid = (unsigned)((A.P + 2 * TC_M - 1) / (2 * TC_M));
  k_dense_tc2<K8><<<grid, TC2_THREADS, sm, st>>>(A, pa, hmc ? 1 : 0);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return EHMC_OK;
}

int launch_dense_tc(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<float>& A, int integ, bool hmc, cudaStream_t st) {
  if (c->dense_path == 0 || c->dense_path >= 4) return launch_dense_tc3(c, p, A, integ, hmc, st);
  if (c->dense_path != 2) {  // two-tile TS-mode kernel (default); dense_path = 2 selects the one-tile SS kernel
    switch (p->tc_kp / 8) {
      case 3: return launch_tc2_k8<3>(c, p, A, hmc, st);
      case 4: return launch_tc2_k8<4>(c, p, A, hmc, st);
      case 5: return launch_tc2_k8<5>(c, p, A, hmc, st);
      case 6: return launch_tc2_k8<6>(c, p, A, hmc, st);
      case 7: return launch_tc2_k8<7>(c, p, A, hmc, st);
      case 8: return launch_tc2_k8<8>(c, p, A, hmc, st);
      case 9: return launch_tc2_k8<9>(c, p, A, hmc, st);
      case 10: return launch_tc2_k8<10>(c, p, A, hmc, st);
      case 11: return launch_tc2_k8<11>(c, p, A, hmc, st);
      case 12: return launch_tc2_k8<12>(c, p, A, hmc, st);
      case 13: return launch_tc2_k8<13>(c, p, A, hmc, st);
    }
    return fail(EHMC_ERR_UNSUPPORTED, "dense tensor-core kernel: D = %d not packed", p->D);
  }
  switch (p->tc_kp / 8) {
    case 3: return launch_tc_k8<3>(c, p, A, hmc, st);
    case 4: return launch_tc_k8<4>(c, p, A, hmc, st);
    case 5: return launch_tc_k8<5>(c, p, A, hmc, st);
    case 6: return launch_tc_k8<6>(c, p, A, hmc, st);
    case 7: return launch_tc_k8<7>(c, p, A, hmc, st);
    case 8: return launch_tc_k8<8>(c, p, A, hmc, st);
    case 9: return launch_tc_k8<9>(c, p, A, hmc, st);
    case 10: return launch_tc_k8<10>(c, p, A, hmc, st);
    case 11: return launch_tc_k8<11>(c, p, A, hmc, st);
    case 12: return launch_tc_k8<12>(c, p, A, hmc, st);
    case 13: return launch_tc_k8<13>(c, p, A, hmc, st);
  }
  return fail(EHMC_ERR_UNSUPPORTED, "dense tensor-core kernel: D = %d not packed", p->D);
}

}  // namespace ehmc
