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
//   * compute CTAs 0 .. n-2 run k_small_body for iteration `it` over their grid-stride share of the particles (the
//     same particles every iteration: q round-trips through HBM, nothing crosses CTAs), deposit their row of 2D+3
//     partial sums and take a ticket.  They never wait for the statistics: iteration `it` only needs the step size
//     h[it], published two iterations earlier.
//   * the SERVICE CTA (last block) waits for all tickets of iteration `it`, adds the rows in a fixed order, pushes
//     the 2D+3 doubles into every peer GPU's mailbox with plain stores over NVLink (peer memory mapped through CUDA
//     IPC), waits for the peers' pushes, adds the world's vectors in rank order (every rank computes the same
//     bits), updates log h and publishes h[it + 2]: the one-iteration-stale pipeline of HMC.run, so the
//     reduction, the NVLink round trip (~184 B out and in per peer, a few microseconds) and the update all hide
//     behind iteration it + 1.
// Per-iteration cost on the critical path: nothing but the trajectory work itself.
#pragma once

#include "k_small.cuh"

namespace ehmc {

constexpr int ENS_MB_STRIDE = 72;  // doubles per mailbox slot: up to 2 * 32 + 3 statistics, last one = sequence flag

template <typename T>
struct EnsRunArgs {
  int nIter;
  int adaptIters;      // Robbins-Monro updates during the first adaptIters iterations of this launch
  double target, gain0, kappa, maxMove, logLo, logHi;
  double Ptot;         // particles of all ranks
  double* hsched;      // [nIter + 2] step size of every iteration
  long long* published;  // [1] number of valid hsched entries
  unsigned* ticket;    // [2] compute CTAs that finished the iteration of this parity
  double* rows;        // [2][ncompute][2D+3]
  double* state;       // [4] in/out: step size, log step size, updates k, iterations run
  double* history;     // [nIter][4] {accept rate, mean acceptance probability, mean H, step size used} or null
  double* moments;     // [2D] += sum q_d, sum q_d^2 of every iteration, or null
  T* trace;            // [D][ntrace][S] kept positions of the first ntrace local particles, or null
  long long ntrace, S, s0;
  int rank, world;
  double* const* peers;   // [world] every rank's mailbox [2][world][ENS_MB_STRIDE] (peers[rank] = the local one)
  unsigned long long seq0;  // iterations this communicator has reduced before this launch
};

__device__ __forceinline__ long long ld_acquire_gpu(const long long* p) {
  long long v;
  asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(long long* p, long long v) {
  asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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

template <typename T, int DT, class Pot, int INTEG, bool EXACT>
__global__ void __launch_bounds__(K1_THREADS) k_small_ens(const IterArgs<T> Ain, const Pot pot, const EnsRunArgs<T> R) {
  extern __shared__ double k1_smem[];
  const int Dn = EXACT ? DT : Ain.D;
  const int NS = 2 * Dn + 3;
  const unsigned ncompute = gridDim.x - 1;
  const int tid = threadIdx.x;

  if (blockIdx.x < ncompute) {
    // ===== compute CTAs =====
    IterArgs<T> A = Ain;
    for (int it = 0; it < R.nIter; ++it) {
      if (tid == 0) {
        while (ld_acquire_gpu(R.published) < it + 1) __nanosleep(200);
      }
      __syncthreads();
      const double h = __ldcg(&R.hsched[it]);
      A.h = (T)h;
      A.h2 = A.h * A.h;
      A.iter = Ain.iter + (u64)it;
      A.partials = R.rows + ((size_t)(it & 1) * ncompute) * NS;
      k_small_body<T, DT, Pot, INTEG, true, EXACT>(A, pot, k1_smem, blockIdx.x, ncompute);
      if (R.trace != nullptr) {
        // kept positions of the first ntrace local particles (each thread re-reads what it wrote itself)
        const long long stride = (long long)ncompute * K1_THREADS;
        for (long long i = (long long)blockIdx.x * K1_THREADS + tid; i < R.ntrace; i += stride)
          for (int d = 0; d < Dn; ++d) R.trace[((long long)d * R.ntrace + i) * R.S + R.s0 + it] = A.q[d * A.q_ld + i];
      }
      __syncthreads();  // the row of partial sums is written (and the reduction scratch is free again)
      if (tid == 0) {
        __threadfence();
        atomicAdd(&R.ticket[it & 1], 1u);
      }
    }
    return;
  }

  // ===== service CTA: reduce, all-reduce over NVLink, adapt, publish =====
  double* sm = k1_smem;              // [4][32] slice sums, then [NS] local vector, [NS] total
  __shared__ double s_h, s_logh;
  __shared__ unsigned long long s_k;
  if (tid == 0) {
    s_h = R.state[0];
    s_logh = R.state[1];
    s_k = (unsigned long long)R.state[2];
    R.hsched[0] = s_h;
    R.hsched[1] = s_h;
    __threadfence();
    st_release_gpu(R.published, 2);
  }
  __syncthreads();
  double* loc = sm + 4 * 32;   // [<= 67]
  double* tot = loc + 72;      // [<= 67]
  for (int it = 0; it < R.nIter; ++it) {
    const int par = it & 1;
    if (tid == 0) {
      while (ld_acquire_gpu_u32(&R.ticket[par]) < ncompute) __nanosleep(500);
      R.ticket[par] = 0u;  // nobody touches this parity again before h[it + 2] is published below
    }
    __syncthreads();
    // rows -> local vector: column j by thread (j % 32), four row slices, fixed order
    const double* rows = R.rows + ((size_t)par * ncompute) * NS;
    for (int j0 = 0; j0 < NS; j0 += 32) {
      const int j = j0 + (tid & 31), sl = tid >> 5;
      double s = 0.0;
      if (j < NS) {
        const unsigned r0 = ncompute * sl / 4, r1 = ncompute * (sl + 1) / 4;
        for (unsigned r = r0; r < r1; ++r) s += __ldcg(&rows[(size_t)r * NS + j]);
      }
      sm[sl * 32 + (tid & 31)] = s;
      __syncthreads();
      if (tid < 32 && j0 + tid < NS) loc[j0 + tid] = (sm[tid] + sm[32 + tid]) + (sm[64 + tid] + sm[96 + tid]);
      __syncthreads();
    }
    if (R.world > 1) {
      const unsigned long long seq = R.seq0 + (unsigned long long)it + 1ull;
      const size_t slot = ((size_t)par * R.world + R.rank) * ENS_MB_STRIDE;
      // push: payload to every peer, then the sequence flag (release at system scope orders it behind the payload)
      for (int x = tid; x < R.world * NS; x += K1_THREADS) {
        const int r = x / NS, j = x % NS;
        if (r != R.rank) R.peers[r][slot + j] = loc[j];
      }
      __threadfence_system();
      __syncthreads();
      if (tid < R.world && tid != R.rank)
        st_release_sys(reinterpret_cast<unsigned long long*>(R.peers[tid] + slot + ENS_MB_STRIDE - 1), seq);
      // wait for every peer's push of this iteration
      if (tid < R.world && tid != R.rank) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(
            R.peers[R.rank] + ((size_t)par * R.world + tid) * ENS_MB_STRIDE + ENS_MB_STRIDE - 1);
        while (ld_acquire_sys(f) < seq) __nanosleep(100);
      }
      __syncthreads();
      __threadfence_system();
      if (tid < NS) {
        double s = 0.0;
        for (int r = 0; r < R.world; ++r)  // rank order: the same bits on every rank
          s += r == R.rank ? loc[tid]
                           : *reinterpret_cast<volatile const double*>(
                                 R.peers[R.rank] + ((size_t)par * R.world + r) * ENS_MB_STRIDE + tid);
        tot[tid] = s;
      }
    } else if (tid < NS) {
      tot[tid] = loc[tid];
    }
    __syncthreads();
    if (R.moments != nullptr && tid < 2 * Dn) R.moments[tid] += tot[3 + tid];
    if (tid == 0) {
      const double meanAcc = tot[1] / R.Ptot;
      if (R.history != nullptr) {
        R.history[4 * it + 0] = tot[0] / R.Ptot;
        R.history[4 * it + 1] = meanAcc;
        R.history[4 * it + 2] = tot[2] / R.Ptot;
        R.history[4 * it + 3] = R.hsched[it];
      }
      // Robbins-Monro on log h (parallel.StepSizeAdapter / ehmc_adapt_step): the update computed from iteration
      // `it` is first used by iteration it + 2
      if (it < R.adaptIters) {
        s_k += 1ull;
        const double acc = isfinite(meanAcc) ? meanAcc : 0.0;
        double move = R.gain0 / pow((double)s_k, R.kappa) * (acc - R.target);
        move = fmin(fmax(move, -R.maxMove), R.maxMove);
        s_logh = fmin(fmax(s_logh + move, R.logLo), R.logHi);
        s_h = exp(s_logh);
      }
      R.hsched[it + 2] = s_h;
      __threadfence();
      st_release_gpu(R.published, (long long)it + 3);
    }
    __syncthreads();
  }
  if (tid == 0) {
    // the step size the NEXT iteration would use (the host loop's stepSize after the run)
    R.state[0] = s_h;
    R.state[1] = s_logh;
    R.state[2] = (double)s_k;
    R.state[3] += (double)R.nIter;
  }
}

}  // namespace ehmc
