// Host-side definitions shared by the translation units of _ehmc.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/ehmc.h"
#include "common.cuh"

// ---- errors ------------------------------------------------------------------
int ehmc_fail(int code, const char* fmt, ...);
#define fail ehmc_fail

#define CUDA_TRY(expr)                                                                                         \
  do {                                                                                                         \
    cudaError_t e__ = (expr);                                                                                  \
    if (e__ != cudaSuccess)                                                                                    \
      return fail(EHMC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define TRY(expr)                     \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != EHMC_OK) return rc__; \
  } while (0)

// ---- context / potential objects ------------------------------------------------
struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes);
  void release();
};

constexpr int N_STAGE = 3;  // host path: chunks in flight

struct ehmc_ctx {
  int device = 0;
  cudaDeviceProp prop;
  uint64_t launches = 0;
  long long last_rows = 0;  // rows of statistics partials the last trajectory launch wrote
  DevBuf partials;        // block partial sums for the statistics
  DevBuf stage[N_STAGE];  // host path staging (one slab per in-flight chunk)
  DevBuf stage_stats;
  DevBuf pstats;          // per-particle statistics rows [P][3] (CTA-per-particle / unfused families)
  DevBuf uf[N_STAGE];     // scratch of the unfused path: w, v, g [D][P] + K0, U0, U1 [P]
  // endpoint cache of the unfused (logistic) family: grad U and U at the position the last ehmc_hmc_iter kept
  DevBuf ep_grad[2], ep_energy[2];
  int ep_cur = 0;                  // slot holding the valid cache
  bool ep_valid = false;
  bool ep_enabled = false;         // set by the device path of ehmc_hmc_iter only (the host path stages chunks)
  const void* ep_q = nullptr;      // identity of the cached call: q pointer, potential, P, dtype
  const ehmc_potential* ep_pot = nullptr;
  long long ep_P = 0;
  int ep_bits = 0;
  cudaStream_t streams[N_STAGE] = {nullptr, nullptr, nullptr};
  // tuning options (ehmc_ctx_set_option)
  int small_waves = 8;            // k_small grid = this many resident waves of CTAs (grid-stride over particles)
  int nbody_ti = 0;               // bodies per thread of k_nbody: 0 auto, 4 or 8
  int dense_occupancy = 2;        // CTAs/SM the float32 dense kernel is compiled for (1 or 2)
  int dense_path = 0;             // 0 auto (3xFP16 tensor cores when eligible), 1 CUDA cores (exact fp32 FMA),
                                  // 4 force 3xFP16 (even when the split of Lambda loses accuracy)
  long long host_chunk_bytes = 32LL << 20;
  int tc_prof = 0;                // 1: record a clock64 trace of CTA 0 into tc_prof_buf (64 x int64)
  DevBuf tc_prof_buf;
  int ens_groups = 0;             // fused ensemble run: groups of batches per iteration at most (one float64 row each); 0 = auto
  int ens_sshift = -1;            // fused ensemble run: log2(sub-batches per queue item), -1 = by shard size
  int ens_debug = 0;              // iterations of the fused ensemble run whose phase stamps are recorded
  DevBuf ens_dbg_buf;
  DevBuf ens_ctl;                 // control block of the fused ensemble run: step-size schedule, tickets, partial rows
  DevBuf overflow;                // unsigned: integrate() rows that saturated the fp16 operand range of k_dense_tc3
  int tc_debug = 0;               // profiling knobs of the tensor-core kernels (see DenseTcArgs::dbg)
};

struct ehmc_potential {
  ehmc_ctx* ctx = nullptr;
  int family = 0;
  int D = 0;
  int bits = 32;
  std::vector<double> hp0, hp1;  // host copies of the parameters (double)
  std::vector<double> scalars;
  void* d0 = nullptr;  // dense: packed Ls ; nbody: body masses ; logistic: X
  void* d1 = nullptr;  // dense: mu (padded) ; logistic: y
  void* d2 = nullptr;  // dense: plain Lambda row-major (eval kernel)
  int TN = 0;          // dense tile selection
  void* d6 = nullptr;  // logistic tensor-core path: packed bf16 X chunks + y
  int lt_nc = 0, lt_dp = 0, lt_npad = 0;
  unsigned lt_chunk_bytes = 0;
  void* d7 = nullptr;  // dense float32 3xFP16 tensor-core path: Lambda_hi fp16 [KP/8][NP][8]
  void* d8 = nullptr;  //                                       Lambda_lo
  void* d9 = nullptr;  //                                       mu [128]
  int tc3_c8 = 0;      // ceil(D / 8) (0 = not eligible)
  int tc3_ok = 0;      // 1: the (hi, lo) fp16 split represents Lambda to float32 accuracy (no entry flushes)
  float tc3_inv_lscale = 1.f;
  int use_tc = 0;      // logistic (scalars[1]): 0 = CUDA cores (exact), 1 = bf16 tensor-core gradient,
                       // 2 = float32-accurate tensor-core gradient (3-pass fp16 split)
  float lts_x_iscale = 1.f;  // 1 / (power-of-two scale of the packed fp16 X)
  int B = 0;           // nbody: bodies per particle
  int N = 0;           // logistic: data rows
};

// peer-memory mailboxes of the fused ensemble run (comm.cu)
struct ehmc_comm {
  ehmc_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  size_t bytes = 0;
  void* mailbox = nullptr;                       // [2][world][ENS_MB_STRIDE] doubles, local
  double* peers[EHMC_COMM_MAX_RANKS] = {};       // every rank's mailbox as mapped into this process
  bool opened[EHMC_COMM_MAX_RANKS] = {};
  double** peers_dev = nullptr;                  // device copy of peers[]
  bool connected = false;
  unsigned long long seq = 0;                    // iterations reduced so far (identical on all ranks)
};

// ---- launchers (explicitly instantiated for float / double in inst_*.cu) ----------
namespace ehmc {

constexpr int K1_THREADS_HOST = 128;
constexpr int K2_WARPS_HOST = 8;

// fused trajectory kernels, D <= 32 register-resident families
template <typename T>
int launch_small(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, int integ, bool hmc, cudaStream_t st);
// getSamples' whole loop in one launch (small-D families, Philox draws)
template <typename T>
int run_small(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, int integ, const RunArgs<T>& R, cudaStream_t st);
// the adaptive ensemble run as one persistent launch (k_small_ens.cuh)
template <typename T>
struct EnsRunArgs;
template <typename T>
int run_small_ens(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, const EnsRunArgs<T>& R, cudaStream_t st);
// dense Gaussian, 16 < D <= 128
template <typename T>
int launch_dense(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, int integ, bool hmc, cudaStream_t st);
// float32 3xFP16 persistent tensor-core variant (leapfrog and Stormer-Verlet; the default)
int launch_dense_tc3(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<float>& A, int integ, bool hmc, cudaStream_t st);
template <typename T>
int dense_particles_per_cta();
int dense_tn(int D);
template <typename T>
int dense_tnp(int TN);
// potential evaluation for the register-resident families
template <typename T>
int eval_small(ehmc_ctx* c, const ehmc_potential* p, const T* q, long long q_ld, long long P, T* e, T* g,
               long long g_ld, cudaStream_t st);

// pairwise gravitational family (one CTA per ensemble particle)
template <typename T>
int launch_nbody(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, int integ, bool hmc, cudaStream_t st);
template <typename T>
int eval_nbody(ehmc_ctx* c, const ehmc_potential* p, const T* q, long long q_ld, long long P, T* e, T* g,
               long long g_ld, cudaStream_t st);
// sum_i q[d,i], sum_i q[d,i]^2 -> out[3 + d], out[3 + D + d]
template <typename T>
int colstats(ehmc_ctx* c, const T* q, long long q_ld, long long P, int D, double* out, cudaStream_t st);
// logistic regression: unfused trajectory driver (gradient kernel + kick/drift kernels)
template <typename T>
int launch_logistic(ehmc_ctx* c, const ehmc_potential* p, const IterArgs<T>& A, int integ, bool hmc, cudaStream_t st,
                    int slot);
// bf16 tensor-core gradient (float32 state only)
int logistic_grad_tc(ehmc_ctx* c, const ehmc_potential* p, const float* theta, long long t_ld, long long P, float* g,
                     long long g_ld, float* e, double* e64, cudaStream_t st);
// float32-accurate tensor-core gradient: 3-pass fp16 split (float32 state only)
int logistic_grad_tcs(ehmc_ctx* c, const ehmc_potential* p, const float* theta, long long t_ld, long long P, float* g,
                      long long g_ld, float* e, double* e64, cudaStream_t st);
template <typename T>
int eval_logistic(ehmc_ctx* c, const ehmc_potential* p, const T* q, long long q_ld, long long P, T* e, T* g,
                  long long g_ld, cudaStream_t st);

}  // namespace ehmc
