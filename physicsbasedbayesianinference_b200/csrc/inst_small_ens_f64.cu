#include "inst_small_ens.cuh"
namespace ehmc {
template int run_small_ens<double>(ehmc_ctx*, const ehmc_potential*, const IterArgs<double>&, const EnsRunArgs<double>&, cudaStream_t);
}  // namespace ehmc
