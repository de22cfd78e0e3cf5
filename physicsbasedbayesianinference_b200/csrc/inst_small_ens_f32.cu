#include "inst_small_ens.cuh"
namespace ehmc {
template int run_small_ens<float>(ehmc_ctx*, const ehmc_potential*, const IterArgs<float>&, const EnsRunArgs<float>&, cudaStream_t);
}  // namespace ehmc
