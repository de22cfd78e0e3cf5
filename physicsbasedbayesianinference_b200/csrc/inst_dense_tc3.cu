// 3xFP16 persistent tensor-core kernel of the dense Gaussian family (its own translation unit: the
// fully unrolled epilogues make it the slowest file of the build).
#include <algorithm>

#include "host_defs.h"
#include "k_dense_tc3.cuh"

namespace ehmc {

template <int C8>
static int launch_tc3_c8(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<float>& A, int integ, bool hmc,
                         cudaStream_t st) {
  constexpr size_t sm = Tc3Shape<C8>::smem_bytes();
  static_assert(sm <= 227 * 1024, "dense tensor-core kernel (fp16 split): shared memory");
  CUDA_TRY(cudaFuncSetAttribute(k_dense_tc3<C8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  DenseTc3Args pa;
  pa.Bhi = static_cast<const __half*>(p->d7);
  pa.Blo = static_cast<const __half*>(p->d8);
  pa.mu = static_cast<const float*>(p->d9);
  pa.inv_lscale = p->tc3_inv_lscale;
  pa.dbg = c->tc_debug;
  pa.prof = c->tc_prof ? static_cast<long long*>(c->tc_prof_buf.ptr) : nullptr;
  if (!hmc && c->overflow.ptr == nullptr) {
    TRY(c->overflow.ensure(sizeof(unsigned)));
    CUDA_TRY(cudaMemsetAsync(c->overflow.ptr, 0, sizeof(unsigned), st));
  }
  pa.overflow = static_cast<unsigned*>(c->overflow.ptr);
  // persistent: one CTA per SM (the CTA takes the whole TMEM), tiles dealt in contiguous ranges
  const long long ntiles = (A.P + TC_M - 1) / TC_M;
  const unsigned grid = (unsigned)std::min<long long>(c->prop.multiProcessorCount, ntiles);
  k_dense_tc3<C8><<<grid, TC3_THREADS, sm, st>>>(A, pa, hmc ? 1 : 0, integ);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return EHMC_OK;
}

int launch_dense_tc3(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<float>& A, int integ, bool hmc, cudaStream_t st) {
  switch (p->tc3_c8) {
    case 3: return launch_tc3_c8<3>(c, p, A, integ, hmc, st);
    case 4: return launch_tc3_c8<4>(c, p, A, integ, hmc, st);
    case 5: return launch_tc3_c8<5>(c, p, A, integ, hmc, st);
    case 6: return launch_tc3_c8<6>(c, p, A, integ, hmc, st);
    case 7: return launch_tc3_c8<7>(c, p, A, integ, hmc, st);
    case 8: return launch_tc3_c8<8>(c, p, A, integ, hmc, st);
    case 9: return launch_tc3_c8<9>(c, p, A, integ, hmc, st);
    case 10: return launch_tc3_c8<10>(c, p, A, integ, hmc, st);
    case 11: return launch_tc3_c8<11>(c, p, A, integ, hmc, st);
    case 12: return launch_tc3_c8<12>(c, p, A, integ, hmc, st);
    case 13: return launch_tc3_c8<13>(c, p, A, integ, hmc, st);
    case 14: return launch_tc3_c8<14>(c, p, A, integ, hmc, st);
    case 15: return launch_tc3_c8<15>(c, p, A, integ, hmc, st);
    case 16: return launch_tc3_c8<16>(c, p, A, integ, hmc, st);
  }
  return fail(EHMC_ERR_UNSUPPORTED, "dense tensor-core kernel (fp16 split): D = %d not packed", p->D);
}

}  // namespace ehmc
