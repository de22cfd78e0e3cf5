#include "inst_small.cuh"
namespace ehmc {
template int launch_small<double>(ehmc_ctx*, const ehmc_potential*, const IterArgs<double>&, int, bool, cudaStream_t);
template int eval_small<double>(ehmc_ctx*, const ehmc_potential*, const double*, long long, long long, double*,
                                double*, long long, cudaStream_t);
template int run_small<double>(ehmc_ctx*, const ehmc_potential*, const IterArgs<double>&, int, const RunArgs<double>&, cudaStream_t);
}  // namespace ehmc
