// K6: the adaptive ensemble run as ONE persistent launch -- trajectory kernel, ensemble statistics, the all-reduce
// of those statistics across GPUs and the step-size update fused into a single cooperative kernel.
//
// What it replaces: the outer loop of HMC.getSamples (src/HMC.py:150-179) as HMC.run drives it for large ensembles
// (momentum refresh, trajectory, Metropolis per iteration; build-defined: per-iteration ensemble statistics
// {n_accept, sum acceptance probability, sum H, sum q_d, sum q_d^2} summed over ALL GPUs, Robbins-Monro step-size
// adaptation from them).  The iteration-by-iteration form costs two launches, an NCCL all-reduce of 2D+3 doubles, a
// pinned copy and the Python enqueue per iteration: 63 us per iteration at the 8-GPU shard of config 5 (2^19
// particles) against 33 us of kernel (profiles/r01_adapt_probe.txt).
//
// Structure (grid = one resident wave, cooperative launch so that every CTA is co-resident):
//   * COMPUTE CTAs run k_small_body for iteration `it` over their grid-stride share of the particles (the same
//     particles every iteration: q round-trips through HBM, nothing crosses CTAs), deposit their row of 2D+3 partial
//     sums and release a ticket (fire and forget).  They never wait for the statistics: iteration `it` only needs
//     the step size h[it], published two iterations earlier, and even that wait sits behind the momentum draw.
//   * 8 SERVICE CTAs = 32 independent reducer warps: warp w adds the rows of compute CTAs 32 w .. 32 w + 31 (fixed
//     order) when their tickets are in; the master warp adds the ~33 group rows, pushes the 2D+3 doubles into every
//     peer GPU's mailbox with plain stores over NVLink (peer memory mapped through CUDA IPC), waits for the peers'
//     pushes, adds the world's vectors in rank order (every rank computes the same bits), updates log h and
//     publishes h[it + 2]: the one-iteration-stale pipeline of HMC.run, so the reductions, the NVLink round trip
//     (~184 B out and in per peer, a few microseconds) and the update all hide behind iteration it + 1.
// Per-iteration cost on the compute CTAs' path: the block reduction of their own row.
#pragma once

#include "k_small.cuh"

namespace ehmc {

constexpr int ENS_PUB_COPIES = 64; // replicas of the "step sizes published" counter, one 128-byte line each
constexpr int ENS_SERVICE_CTAS = 8; // blocks of reducer warps (one of them also runs the all-reduce and the update)
constexpr int ENS_GROUP = 32;      // compute CTAs per first-level reduction group
constexpr int ENS_MB_STRIDE = 72;  // doubles per mailbox slot: up to 2 * 32 + 3 statistics, last one = sequence flag

template <typename T>
struct EnsRunArgs {
  int nIter;
  int adaptIters;      // Robbins-Monro updates during the first adaptIters iterations of this launch
  double target, maxMove, logLo, logHi;
  const double* gains;  // [adaptIters] Robbins-Monro gain of every update, gain0 / k^kappa (host-computed)
  double Ptot;         // particles of all ranks
  double* hsched;      // [nIter + 2] step size of every iteration
  long long* published;  // [ENS_PUB_COPIES][16] number of valid hsched entries (replicated, one line per copy)
  unsigned long long* arrived;  // [ENS_PUB_COPIES][16] compute CTAs that finished an iteration, ever (soft grid barrier)
  int lockstep;        // 1: iterations start in step (soft grid barrier)
  unsigned* ticket;    // [2][32] (entry 0 of each 128-byte line) groups whose rows of the iteration of this parity are reduced
  unsigned* gticket;   // [2][ngroups] compute CTAs of the group that finished the iteration
  double* rows;        // [2][ncompute][2D+3] per-CTA partial sums
  double* grows;       // [2][ngroups][2D+3] per-group sums
  double* state;       // [4] in/out: step size, log step size, updates k, iterations run
  double* history;     // [nIter][4] {accept rate, mean acceptance probability, mean H, step size used} or null
  double* moments;     // [2D] += sum q_d, sum q_d^2 of every iteration, or null
  T* trace;            // [D][ntrace][S] kept positions of the first ntrace local particles, or null
  long long ntrace, S, s0;
  int rank, world;
  double* const* peers;   // [world] every rank's mailbox [2][world][ENS_MB_STRIDE] (peers[rank] = the local one)
  unsigned long long seq0;  // iterations this communicator has reduced before this launch
  long long* dbg;      // optional %globaltimer stamps [dbg_iters][8] (ctx option "ens_debug"):
  int dbg_iters;       //   service: 0 tickets complete, 1 reduced, 2 pushed, 3 peers in, 4 published;
                       //   compute CTA 0: 5 waits for h, 6 starts, 7 done
};

__device__ __forceinline__ long long ld_acquire_gpu(const long long* p) {
  long long v;
  asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(long long* p, long long v) {
  asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Step size of iteration `it`, waited for INSIDE the trajectory body, behind the momentum draw of the CTA's first
// particles (~300 instructions of Philox work that do not depend on h): lane 0 of every warp polls its replica of the
// "published" counter, the warp shares the value.
template <typename T>
struct EnsStepHook {
  const long long* pub;
  const double* hsched;
  int it;
  long long* dbg;  // stamp slot of this iteration (thread 0 of CTA 0) or null
  __device__ __forceinline__ void operator()(T& h, T& h2) const {
    if ((threadIdx.x & 31) == 0) {
      unsigned long long t0 = 0, t1 = 0;
      if (dbg != nullptr && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      while (ld_acquire_gpu(pub) < it + 1) __nanosleep(250);
      if (dbg != nullptr && threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        dbg[6] = (long long)(t1 - t0);  // time spent waiting for the step size
      }
    }
    __syncwarp();
    h = (T)__ldcg(&hsched[it]);
    h2 = h * h;
  }
};

template <typename T, int DT, class Pot, int INTEG, bool EXACT>
__global__ void __launch_bounds__(K1_THREADS) k_small_ens(const IterArgs<T> Ain, const Pot pot, const EnsRunArgs<T> R) {
  extern __shared__ double k1_smem[];
  const int Dn = EXACT ? DT : Ain.D;
  const int NS = 2 * Dn + 3;
  const unsigned ncompute = gridDim.x - ENS_SERVICE_CTAS;
  const unsigned ngroups = (ncompute + ENS_GROUP - 1) / ENS_GROUP;
  const int tid = threadIdx.x;

  if (blockIdx.x < ncompute) {
    // ===== compute CTAs =====
    IterArgs<T> A = Ain;
    EnsStepHook<T> hook;
    // (64 replicas of the counter on separate 128-byte lines: a thousand waiting CTAs polling ONE line kept its L2
    // slice busy enough to slow the ticket atomics in its neighbourhood)
    hook.pub = R.published + (blockIdx.x % ENS_PUB_COPIES) * 16;
    hook.hsched = R.hsched;
    for (int it = 0; it < R.nIter; ++it) {
      auto stamp = [&](int k) {
        if (R.dbg != nullptr && blockIdx.x == 0 && tid == 0 && it < R.dbg_iters) {
          unsigned long long t;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
          R.dbg[it * 8 + k] = (long long)t;
        }
      };
      stamp(5);
      if (R.lockstep && it > 0) {
        // soft grid barrier: nobody starts iteration `it` before every CTA has finished iteration it - 1.  Without it
        // the warp schedulers' preference for the oldest warps lets the oldest CTAs of every SM run up to two
        // iterations ahead and then wait there, so that on average only about half of the resident warps are
        // runnable: measured 39.8 us per iteration at 2^19 particles against 33 us for iterations that start in step.
        if (tid == 0) {
          const unsigned long long want = (unsigned long long)ncompute * (unsigned long long)it;
          while (ld_acquire_gpu_u64(R.arrived + (blockIdx.x % ENS_PUB_COPIES) * 16) < want) __nanosleep(100);
        }
        __syncthreads();
      }
      hook.it = it;
      hook.dbg = (R.dbg != nullptr && blockIdx.x == 0 && it < R.dbg_iters) ? R.dbg + it * 8 : nullptr;
      A.iter = Ain.iter + (u64)it;
      A.partials = R.rows + ((size_t)(it & 1) * ncompute) * NS;
      k_small_body<T, DT, Pot, INTEG, true, EXACT>(A, pot, k1_smem, blockIdx.x, ncompute, hook);
      if (R.trace != nullptr) {
        // kept positions of the first ntrace local particles (each thread re-reads what it wrote itself)
        const long long stride = (long long)ncompute * K1_THREADS;
        for (long long i = (long long)blockIdx.x * K1_THREADS + tid; i < R.ntrace; i += stride)
          for (int d = 0; d < Dn; ++d) R.trace[((long long)d * R.ntrace + i) * R.S + R.s0 + it] = A.q[d * A.q_ld + i];
      }
      __syncthreads();  // the row of partial sums is written (and the reduction scratch is free again)
      // fire and forget: the row is released to the reducer warp of this CTA's group; nothing here waits
      if (tid == 0)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&R.gticket[(it & 1) * ngroups + blockIdx.x / ENS_GROUP])
                     : "memory");
      if (R.lockstep && tid < ENS_PUB_COPIES)  // arrival counters, replicated like `published` (all of them count every CTA)
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(R.arrived + tid * 16) : "memory");
      stamp(7);
    }
    return;
  }

  // ===== service CTAs (the last ENS_SERVICE_CTAS blocks): 32 independent warps, no block-wide barrier =====
  // Warp w reduces the rows of groups w, w + 32, ... (ENS_GROUP consecutive compute CTAs each; lane j adds column j over
  // the group's 32 rows: independent loads, one L2 round trip, fixed order) as soon as the group's tickets are in.
  // Warp 0 is also the MASTER: group rows -> this GPU's vector, all-reduce over NVLink, step-size update, publish.
  // Everything that polls or waits for L2 / NVLink latency lives here, off the compute CTAs' path; 190 KB of rows
  // through ONE CTA took longer than an iteration of the 8-GPU shard and set the pace of the whole run.
  const int lane = tid & 31;
  const unsigned sw = (blockIdx.x - ncompute) * (K1_THREADS / 32) + (tid >> 5);  // service warp index, 0 .. 31
  constexpr unsigned NSW = ENS_SERVICE_CTAS * (K1_THREADS / 32);
  double* loc = k1_smem;        // (master) [<= 67] this GPU's vector
  double* tot = k1_smem + 72;   // (master) [<= 67] the world's
  const bool master = sw == 0;
  double s_h = 0.0, s_logh = 0.0;
  unsigned long long s_k = 0ull;
  if (master) {
    if (lane == 0) {
      s_h = R.state[0];
      s_logh = R.state[1];
      s_k = (unsigned long long)R.state[2];
      R.hsched[0] = s_h;
      R.hsched[1] = s_h;
      __threadfence();
    }
    __syncwarp();
    for (int c = lane; c < ENS_PUB_COPIES; c += 32) st_release_gpu(R.published + c * 16, 2);
  }
  auto stamp = [&](int it, int k) {
    if (R.dbg != nullptr && master && lane == 0 && it < R.dbg_iters) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      R.dbg[it * 8 + k] = (long long)t;
    }
  };
  for (int it = 0; it < R.nIter; ++it) {
    const int par = it & 1;
    // ---- first level: my groups ----
    for (unsigned g = sw; g < ngroups; g += NSW) {
      const unsigned gsize = min((unsigned)ENS_GROUP, ncompute - g * ENS_GROUP);
      unsigned* gt = &R.gticket[par * ngroups + g];
      if (lane == 0) {
        while (ld_acquire_gpu_u32(gt) < gsize) __nanosleep(100);
        *gt = 0u;  // next used two iterations later, behind the published step size
      }
      __syncwarp();
      const double* rows = R.rows + ((size_t)par * ncompute + (size_t)g * ENS_GROUP) * NS;
      for (int j = lane; j < NS; j += 32) {
        double sum = 0.0;
#pragma unroll 8
        for (unsigned r = 0; r < gsize; ++r) sum += __ldcg(&rows[(size_t)r * NS + j]);
        R.grows[((size_t)par * ngroups + g) * NS + j] = sum;
      }
      __syncwarp();
      if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&R.ticket[32 * par]) : "memory");
    }
    if (!master) continue;
    // ---- master: second level, all-reduce, update, publish ----
    if (lane == 0) {
      while (ld_acquire_gpu_u32(&R.ticket[32 * par]) < ngroups) __nanosleep(100);
      R.ticket[32 * par] = 0u;  // nobody touches this parity again before h[it + 2] is published below
    }
    __syncwarp();
    stamp(it, 0);
    const double* grows = R.grows + ((size_t)par * ngroups) * NS;
    for (int j = lane; j < NS; j += 32) {
      double sum = 0.0;
#pragma unroll 8
      for (unsigned r = 0; r < ngroups; ++r) sum += __ldcg(&grows[(size_t)r * NS + j]);
      loc[j] = sum;
    }
    __syncwarp();
    stamp(it, 1);
    if (R.world > 1) {
      const unsigned long long seq = R.seq0 + (unsigned long long)it + 1ull;
      const size_t slot = ((size_t)par * R.world + R.rank) * ENS_MB_STRIDE;
      // push: payload to every peer, then the sequence flag (release at system scope, behind the payload)
      for (int r = 0; r < R.world; ++r) {
        if (r == R.rank) continue;
        double* dst = R.peers[r] + slot;
        for (int j = lane; j < NS; j += 32) dst[j] = loc[j];
      }
      __threadfence_system();
      __syncwarp();
      if (lane < R.world && lane != R.rank)
        st_release_sys(reinterpret_cast<unsigned long long*>(R.peers[lane] + slot + ENS_MB_STRIDE - 1), seq);
      stamp(it, 2);
      // wait for every peer's push of this iteration
      if (lane < R.world && lane != R.rank) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(
            R.peers[R.rank] + ((size_t)par * R.world + lane) * ENS_MB_STRIDE + ENS_MB_STRIDE - 1);
        while (ld_acquire_sys(f) < seq) {
        }
      }
      __syncwarp();
      stamp(it, 3);
      for (int j = lane; j < NS; j += 32) {
        double s = 0.0;
        for (int r = 0; r < R.world; ++r)  // rank order: the same bits on every rank
          s += r == R.rank ? loc[j]
                           : *reinterpret_cast<volatile const double*>(
                                 R.peers[R.rank] + ((size_t)par * R.world + r) * ENS_MB_STRIDE + j);
        tot[j] = s;
      }
    } else {
      for (int j = lane; j < NS; j += 32) tot[j] = loc[j];
    }
    __syncwarp();
    if (lane == 0) {
      // Robbins-Monro on log h (parallel.StepSizeAdapter / ehmc_adapt_step): the update computed from iteration
      // `it` is first used by iteration it + 2.  gains[i] = gain0 / (k0 + 1 + i)^kappa comes from the host.
      const double meanAcc = tot[1] / R.Ptot;
      if (it < R.adaptIters) {
        s_k += 1ull;
        const double acc = isfinite(meanAcc) ? meanAcc : 0.0;
        double move = R.gains[it] * (acc - R.target);
        move = fmin(fmax(move, -R.maxMove), R.maxMove);
        s_logh = fmin(fmax(s_logh + move, R.logLo), R.logHi);
        s_h = exp(s_logh);
      }
      R.hsched[it + 2] = s_h;
      __threadfence();
    }
    __syncwarp();
    for (int c = lane; c < ENS_PUB_COPIES; c += 32) st_release_gpu(R.published + c * 16, (long long)it + 3);
    stamp(it, 4);
    if (lane == 0 && R.history != nullptr) {
      R.history[4 * it + 0] = tot[0] / R.Ptot;
      R.history[4 * it + 1] = tot[1] / R.Ptot;
      R.history[4 * it + 2] = tot[2] / R.Ptot;
      R.history[4 * it + 3] = R.hsched[it];
    }
    if (R.moments != nullptr)
      for (int j = lane; j < 2 * Dn; j += 32) R.moments[j] += tot[3 + j];
    __syncwarp();
  }
  if (master && lane == 0) {
    // the step size the NEXT iteration would use (the host loop's stepSize after the run)
    R.state[0] = s_h;
    R.state[1] = s_logh;
    R.state[2] = (double)s_k;
    R.state[3] += (double)R.nIter;
  }
}

}  // namespace ehmc
