// K4 (CUDA-core version): Bayesian logistic regression gradient and energy
//   U(theta)    = sum_n [softplus(x_n.theta) - y_n x_n.theta] + 0.5 |theta|^2 / s^2
//   grad U      = X^T (sigmoid(X theta) - y) + theta / s^2
// for a tile of particles, fused flash-attention style: the N x P logits are never
// materialised.  Per chunk of NC data rows:  S = Xc Theta  ->  r = sigmoid(S) - y (registers,
// energy accumulated on the fly)  ->  R to shared memory  ->  G += Xc^T R.
// This is the exact-fp32 / fp64 path (parity mode, any N, D <= 256); the tcgen05 tensor-core
// path for BASELINE config 3 replaces the two contractions with UMMA tiles.
//
// The logistic family is not fused over the L steps: the gradient is ~4 N D flop per
// particle-step (1e8 for config 3), so the O(D) state traffic of separate kick/drift
// kernels is < 0.1 % of the step time.  The generic unfused pieces are below.
#pragma once

#include "common.cuh"

namespace ehmc {

constexpr int LG_THREADS = 256;
constexpr int LG_NC = 32;    // data rows per chunk
constexpr int LG_DMAX = 256;

template <typename T>
struct LogiTile;
template <>
struct LogiTile<float> {
  static constexpr int PT = 64;
};
template <>
struct LogiTile<double> {
  static constexpr int PT = 32;
};

template <typename T>
struct LogisticArgs {
  const T* X;  // [N][D] row-major
  const T* y;  // [N]
  int N, D;
  T inv_s2;
};

template <typename T>
__device__ __forceinline__ T softplus_(T s) {  // log(1 + e^s), stable
  return (s > T(0) ? s : T(0)) + log1p(exp(-fabs(s)));
}
template <typename T>
__device__ __forceinline__ T sigmoid_(T s) {
  return T(1) / (T(1) + exp(-s));
}

// grad[D,P] and/or energy[P] at theta[D,P]
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) k_logistic_grad(const T* __restrict__ theta, long long t_ld, long long P,
                                                              T* __restrict__ grad, long long g_ld,
                                                              T* __restrict__ energy, double* __restrict__ energy64,
                                                              const LogisticArgs<T> pa) {
  constexpr int PT = LogiTile<T>::PT, NC = LG_NC;
  constexpr int RPT = NC * PT / LG_THREADS;        // S rows per thread (phase 1)
  constexpr int GP = 4;                            // particles per thread (phase 2)
  constexpr int NPG = PT / GP;                     // particle groups
  constexpr int NDG = LG_THREADS / NPG;            // dim groups
  constexpr int GD = LG_DMAX / NDG;                // dims per thread
  extern __shared__ __align__(16) unsigned char k4_smem_raw[];
  const int D = pa.D, N = pa.N;
  const int DS = (D + 3) & ~3;                     // padded row length of the X chunk
  T* Th = reinterpret_cast<T*>(k4_smem_raw);       // [D][PT]
  T* Xs = Th + (size_t)D * PT;                     // [NC][DS]
  T* Rs = Xs + (size_t)NC * DS;                    // [NC][PT]
  T* Es = Rs + (size_t)NC * PT;                    // [LG_THREADS / PT][PT] energy partials
  const int tid = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * PT;

  for (int i = tid; i < D * PT; i += LG_THREADS) {
    const int d = i / PT, pp = i % PT;
    Th[i] = (p0 + pp < P) ? theta[d * t_ld + p0 + pp] : T(0);
  }
  // phase-1 mapping: particle pp1, row group ng (rows ng*RPT ..)
  const int pp1 = tid % PT, ng = tid / PT;
  // phase-2 mapping
  const int pg = tid % NPG, dg = tid / NPG;
  T acc[GD][GP];
#pragma unroll
  for (int j = 0; j < GD; ++j)
#pragma unroll
    for (int t = 0; t < GP; ++t) acc[j][t] = T(0);
  T e_acc = T(0);
  const bool want_e = energy != nullptr || energy64 != nullptr;

  for (int n0 = 0; n0 < N; n0 += NC) {
    __syncthreads();  // previous chunk fully consumed (also orders the Th fill on the first pass)
    for (int i = tid; i < NC * DS; i += LG_THREADS) {
      const int r = i / DS, k = i % DS;
      Xs[i] = (n0 + r < N && k < D) ? pa.X[(size_t)(n0 + r) * D + k] : T(0);
    }
    __syncthreads();
    // ---- phase 1: s[r] = x_{n0 + ng*RPT + r} . theta_pp1
    T s[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) s[r] = T(0);
    for (int k = 0; k < D; ++k) {
      const T th = Th[k * PT + pp1];
#pragma unroll
      for (int r = 0; r < RPT; ++r) s[r] = fma(Xs[(ng * RPT + r) * DS + k], th, s[r]);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int n = n0 + ng * RPT + r;
      T res = T(0);
      if (n < N) {
        const T yn = pa.y[n];
        res = sigmoid_<T>(s[r]) - yn;
        if (want_e) e_acc += softplus_<T>(s[r]) - yn * s[r];
      }
      Rs[(ng * RPT + r) * PT + pp1] = res;
    }
    __syncthreads();
    // ---- phase 2: acc[j][t] += sum_r X[r][dg*GD + j] * R[r][pg*GP + t]
#pragma unroll 4
    for (int r = 0; r < NC; ++r) {
      T rv[GP];
#pragma unroll
      for (int t = 0; t < GP; ++t) rv[t] = Rs[r * PT + pg * GP + t];
#pragma unroll
      for (int j = 0; j < GD; ++j) {
        const int d = dg * GD + j;
        const T xv = d < DS ? Xs[r * DS + d] : T(0);
#pragma unroll
        for (int t = 0; t < GP; ++t) acc[j][t] = fma(xv, rv[t], acc[j][t]);
      }
    }
  }
  // ---- epilogue: prior term, stores
  if (grad) {
#pragma unroll
    for (int j = 0; j < GD; ++j) {
      const int d = dg * GD + j;
      if (d < D) {
#pragma unroll
        for (int t = 0; t < GP; ++t) {
          const long long pi = p0 + pg * GP + t;
          if (pi < P) grad[d * g_ld + pi] = acc[j][t] + Th[d * PT + pg * GP + t] * pa.inv_s2;
        }
      }
    }
  }
  if (want_e) {
    __syncthreads();
    Es[ng * PT + pp1] = e_acc;
    __syncthreads();
    if (tid < PT && p0 + tid < P) {
      T e = T(0);
      for (int g = 0; g < LG_THREADS / PT; ++g) e += Es[g * PT + tid];
      T t2 = T(0);
      for (int d = 0; d < D; ++d) t2 += Th[d * PT + tid] * Th[d * PT + tid];
      const T ev = e + T(0.5) * t2 * pa.inv_s2;
      if (energy) energy[p0 + tid] = ev;
      if (energy64) energy64[p0 + tid] = (double)ev;
    }
  }
}

// ---------------------------------------------------------------------------
// Generic unfused HMC pieces (one thread per particle, coalesced over particles).
// State: w = working position, v = velocity, g = gradient, all [D][P] compact (ld = P).
// ---------------------------------------------------------------------------
// init: w = q ; p = z * pstd (fed or Philox) or p = p_in ; v = p / m ; K0[i] = 0.5 |p|^2 / m
template <typename T>
__global__ void k_uf_init(const IterArgs<T> A, T* w, T* v, T* K0, int hmc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.P) return;
  const T m = A.mass[i], inv_m = T(1) / m;
  const T pstd = hmc ? momentum_std<T>(m, A.kB, A.temp, A.pscale) : T(0);
  T ks = T(0);
  constexpr int NB = NormalBlock<T>::N;
  const PhiloxKey K(A.seed, A.iter);
  for (int d0 = 0; d0 < A.D; d0 += NB) {
    T zz[NB];
    if (hmc && A.z == nullptr) NormalBlock<T>::draw(K, A.offset + (u64)i, (uint32_t)(d0 / NB), zz);
#pragma unroll
    for (int t = 0; t < NB; ++t) {
      const int d = d0 + t;
      if (d < A.D) {
        T p;
        if (!hmc)
          p = A.p[d * A.p_ld + i];
        else if (A.z != nullptr)
          p = A.z[d * A.z_ld + i] * pstd;
        else
          p = zz[t] * pstd;
        ks += p * p;
        v[d * A.P + i] = p * inv_m;
        w[d * A.P + i] = A.q[d * A.q_ld + i];
      }
    }
  }
  if (K0) K0[i] = T(0.5) * ks * inv_m;
}

// kick-drift: v -= ck * g / m ; w += cd * v   (ck, cd select half/full kicks and the drift)
template <typename T>
__global__ void k_uf_kick_drift(T* w, T* v, const T* g, const T* mass, long long P, int D, T ck, T cd) {
  // grid.y splits the dimensions: a thread per particle alone leaves too few loads in flight to reach HBM speed
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const T c = ck / mass[i];
  const int dper = (D + (int)gridDim.y - 1) / (int)gridDim.y;
  const int d0 = (int)blockIdx.y * dper, d1 = d0 + dper < D ? d0 + dper : D;
#pragma unroll 4
  for (int d = d0; d < d1; ++d) {
    const T vv = v[d * P + i] - c * g[d * P + i];
    v[d * P + i] = vv;
    if (cd != T(0)) w[d * P + i] += cd * vv;
  }
}

// Stormer-Verlet position update: wn = 2 w - wp + (-g/m) h2 (first: w + v h + 0.5 a h2); v holds qPast after
template <typename T>
__global__ void k_uf_sv_step(T* w, T* v, const T* g, const T* mass, long long P, int D, T h, T h2, int first) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const T inv_m = T(1) / mass[i];
  for (int d = 0; d < D; ++d) {
    const T x = w[d * P + i], a = -g[d * P + i] * inv_m, s = v[d * P + i];
    const T xn = first ? x + s * h + T(0.5) * a * h2 : T(2) * x - s + a * h2;
    v[d * P + i] = x;
    w[d * P + i] = xn;
  }
}

template <typename T>
__global__ void k_uf_sv_finish(const T* w, T* v, long long P, int D, T h) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  for (int d = 0; d < D; ++d) v[d * P + i] = (w[d * P + i] - v[d * P + i]) / h;
}

// final: p = v * m ; newH ; Metropolis ; write q / p_out / accept / per-particle stats partials
template <typename T>
__global__ void k_uf_final(const IterArgs<T> A, const T* w, const T* v, const T* K0, const double* U0,
                           const double* U1, int hmc, double* pstats /* [P][3] or null */, const T* g_start = nullptr,
                           const T* g_end = nullptr, T* g_keep = nullptr, double* u_keep = nullptr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.P) return;
  const T m = A.mass[i];
  if (!hmc) {
    for (int d = 0; d < A.D; ++d) {
      A.q[d * A.q_ld + i] = w[d * A.P + i];
      A.p[d * A.p_ld + i] = v[d * A.P + i] * m;
    }
    return;
  }
  T ks = T(0);
  for (int d = 0; d < A.D; ++d) {
    const T p = v[d * A.P + i] * m;
    ks += p * p;
  }
  // The potential of this family is a sum over N data rows (7e4 at config 3, where a float32 ulp is 0.008): the
  // energies are carried in double and only the DIFFERENCE oldH - newH is rounded to T.  For T = double these are
  // the reference's own expressions (src/HMC.py:108-115).
  const double oldH = (double)K0[i] + U0[i], newH = (double)(T(0.5) * ks / m) + U1[i];
  T u;
  if (A.u != nullptr)
    u = A.u[i];
  else
    u = NormalBlock<T>::uniform(PhiloxKey(A.seed, A.iter), A.offset + (u64)i, A.D);
  T accp;
  const bool rej = metropolis_reject<T>((T)(oldH - newH), T(0), u, A.flags, &accp);
  const T pstd = momentum_std<T>(m, A.kB, A.temp, A.pscale);
  constexpr int NB = NormalBlock<T>::N;
  for (int d = 0; d < A.D; ++d) {
    if (A.p != nullptr) {
      T pv = v[d * A.P + i] * m;
      if (rej) {
        if (A.flags & FLAG_BUGCOMPAT)
          pv = A.q[d * A.q_ld + i];  // HMC.py:176 (sic)
        else if (A.z != nullptr)
          pv = A.z[d * A.z_ld + i] * pstd;
        else {
          T zz[NB];
          NormalBlock<T>::draw(PhiloxKey(A.seed, A.iter), A.offset + (u64)i, (uint32_t)(d / NB), zz);
          T zsel = zz[0];
#pragma unroll
          for (int e = 1; e < NB; ++e)
            if ((d % NB) == e) zsel = zz[e];
          pv = zsel * pstd;
        }
      }
      A.p[d * A.p_ld + i] = pv;
    }
    if (!rej) A.q[d * A.q_ld + i] = w[d * A.P + i];
    // endpoint cache: gradient at the position this particle keeps
    if (g_keep != nullptr) g_keep[d * A.P + i] = rej ? g_start[d * A.P + i] : g_end[d * A.P + i];
  }
  if (u_keep != nullptr) u_keep[i] = rej ? U0[i] : U1[i];
  if (A.accept != nullptr) A.accept[i] = rej ? 0 : 1;
  if (pstats != nullptr) {
    pstats[i * 3 + 0] = rej ? 0.0 : 1.0;
    pstats[i * 3 + 1] = (double)accp;
    pstats[i * 3 + 2] = rej ? oldH : newH;
  }
}

}  // namespace ehmc
