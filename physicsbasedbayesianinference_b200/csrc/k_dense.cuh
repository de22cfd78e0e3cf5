// K2: fused whole-trajectory kernel for the dense-precision Gaussian family
// (U = 0.5 x^T Lambda x, x = q - mu), 16 < D <= 128 -- BASELINE config 2 (D = 100).
//
// The gradient Lambda x over a tile of particles is a small GEMM that is re-issued
// L+1 times on state that never leaves the SM:
//   * CTA = 8 warps; warp w owns output dimensions [w*TN, (w+1)*TN), lane l owns TM
//     consecutive particles -> thread tile TN x TM accumulators in registers.
//   * Lambda (transposed, padded: Ls[k][w][TNP]) and the particle tile x (Xs[k][PT])
//     live in shared memory; per k one 16-byte particle vector + TNP/VW broadcast
//     vectors of Lambda feed TN*TM FMAs.
//   * velocities stay in registers for the whole trajectory; x is updated in place in
//     shared memory between two __syncthreads per step.
// HBM traffic per particle-iteration = read q (+ mass), write q where accepted: the
// kernel is FP32-FMA bound by construction (2 D^2 flop per particle-step).
//
// Reference arithmetic replaced: src/integrator.py:105-120 / :142-163 with
// gradient = Lambda (q - mu), src/HMC.py:106-116,168-176, src/ensemble.py:88-91.
// The leapfrog is evaluated in the algebraically identical kick-drift-kick form
// (v_half = v + a h/2; q += v_half h; v = v_half + a' h/2), which differs from the
// reference's q += v h + a h^2/2 form only by rounding (covered by the 1e-5 / 1e-12
// tolerances); it halves the register state.
#pragma once

#include "common.cuh"

namespace ehmc {

constexpr int K2_WARPS = 8;
constexpr int K2_THREADS = 32 * K2_WARPS;

template <typename T>
struct DenseTile;
template <>
struct DenseTile<float> {
  static constexpr int TM = 4;
  typedef float4 Vec;
};
template <>
struct DenseTile<double> {
  static constexpr int TM = 2;
  typedef double2 Vec;
};

template <typename T, int TN>
struct DenseShape {
  static constexpr int TM = DenseTile<T>::TM;
  static constexpr int VW = 16 / (int)sizeof(T);                 // elements per 16-byte vector
  static constexpr int TNP = ((TN + VW - 1) / VW) * VW;          // padded dims per warp
  static constexpr int PT = 32 * TM;                             // particles per CTA
  static constexpr int DMAX = K2_WARPS * TN;
  static size_t smem_bytes(int D) {
    return sizeof(T) * ((size_t)D * K2_WARPS * TNP + (size_t)D * PT + (size_t)K2_WARPS * PT);
  }
};

template <typename T>
struct DenseArgs {
  const T* Ls;  // [D][8][TNP]: Ls[k][w][j] = Lambda[w*TN + j][k] (0 beyond D)
  const T* mu;  // [8*TN] zero padded
};

template <typename T>
__device__ __forceinline__ void vec_to_arr(const float4& v, T* a) {
  a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
}
template <typename T>
__device__ __forceinline__ void vec_to_arr(const double2& v, T* a) {
  a[0] = v.x; a[1] = v.y;
}
__device__ __forceinline__ float4 arr_to_vec(const float* a) { return make_float4(a[0], a[1], a[2], a[3]); }
__device__ __forceinline__ double2 arr_to_vec(const double* a) { return make_double2(a[0], a[1]); }

// one standard normal of the production stream (cold path: rejected particles' old momentum)
template <typename T>
__device__ __noinline__ T one_normal(u64 seed, u64 iter, u64 pid, int d) {
  constexpr int NB = NormalBlock<T>::N;
  T zz[NB];
  NormalBlock<T>::draw(PhiloxKey(seed, iter), pid, (uint32_t)(d / NB), zz);
  T zsel = zz[0];
#pragma unroll
  for (int e = 1; e < NB; ++e)
    if ((d % NB) == e) zsel = zz[e];
  return zsel;
}

template <typename T, int TN, int MINB>
__global__ void __launch_bounds__(K2_THREADS, MINB)
k_dense(const IterArgs<T> A, const DenseArgs<T> pa, const int integ, const int hmc) {
  typedef DenseShape<T, TN> S;
  typedef typename DenseTile<T>::Vec Vec;
  constexpr int TM = S::TM, VW = S::VW, TNP = S::TNP, PT = S::PT;

  extern __shared__ __align__(16) unsigned char k2_smem_raw[];
  const int D = A.D;
  T* Ls = reinterpret_cast<T*>(k2_smem_raw);      // [D][8][TNP]
  T* Xs = Ls + (size_t)D * K2_WARPS * TNP;         // [D][PT]
  T* red = Xs + (size_t)D * PT;                    // [8][PT]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int d0 = w * TN;
  const long long pbase = (long long)blockIdx.x * PT + lane * TM;

  // ---- stage Lambda ---------------------------------------------------------
  {
    const Vec* src = reinterpret_cast<const Vec*>(pa.Ls);
    Vec* dst = reinterpret_cast<Vec*>(Ls);
    const int nvec = D * K2_WARPS * TNP / VW;
    for (int i = tid; i < nvec; i += K2_THREADS) dst[i] = src[i];
  }

  // ---- per-particle scalars ---------------------------------------------------
  bool valid[TM];
  T m[TM], inv_m[TM], pstd[TM];
#pragma unroll
  for (int t = 0; t < TM; ++t) {
    valid[t] = pbase + t < A.P;
    m[t] = valid[t] ? A.mass[pbase + t] : T(1);
    inv_m[t] = T(1) / m[t];
    pstd[t] = hmc ? momentum_std<T>(m[t], A.kB, A.temp, A.pscale) : T(0);
  }
  const bool full = pbase + TM <= A.P;
  const bool qvec = full && (A.q_ld % TM == 0) && ((reinterpret_cast<uintptr_t>(A.q) & 15) == 0);

  // ---- load x = q - mu for the owned dims, stage into Xs ------------------------
  // st[j][t]: velocity (leapfrog) or past position (Stormer-Verlet)
  T st[TN][TM];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int d = d0 + j;
    T x[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t) x[t] = T(0);
    if (d < D) {
      const T mu = pa.mu[d];
      if (qvec) {
        const Vec qv = *reinterpret_cast<const Vec*>(A.q + d * A.q_ld + pbase);
        vec_to_arr<T>(qv, x);
      } else {
#pragma unroll
        for (int t = 0; t < TM; ++t)
          if (valid[t]) x[t] = A.q[d * A.q_ld + pbase + t];
      }
#pragma unroll
      for (int t = 0; t < TM; ++t) x[t] = valid[t] ? x[t] - mu : T(0);
      *reinterpret_cast<Vec*>(Xs + (size_t)d * PT + lane * TM) = arr_to_vec(x);
    }
  }

  // ---- momentum ------------------------------------------------------------------
  T Kpart[TM];  // sum over owned dims of p^2
#pragma unroll
  for (int t = 0; t < TM; ++t) Kpart[t] = T(0);
  if (hmc) {
    if (A.z != nullptr) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int d = d0 + j;
#pragma unroll
        for (int t = 0; t < TM; ++t)
          st[j][t] = (d < D && valid[t]) ? A.z[d * A.z_ld + pbase + t] * pstd[t] : T(0);
      }
    } else {
      const PhiloxKey K(A.seed, A.iter);
      constexpr int NB = NormalBlock<T>::N;
      constexpr int NBLK = (TN + NB - 2) / NB + 1;
      const int b0 = d0 / NB;
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int t = 0; t < TM; ++t) st[j][t] = T(0);
#pragma unroll 1
      for (int bi = 0; bi < NBLK; ++bi) {
        const int b = b0 + bi;
#pragma unroll
        for (int t = 0; t < TM; ++t) {
          T zz[NB];
          NormalBlock<T>::draw(K, A.offset + (u64)(pbase + t), (uint32_t)b, zz);
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            const int d = d0 + j;
            if (d / NB == b && d < D && valid[t]) {
              T zsel = zz[0];
#pragma unroll
              for (int e = 1; e < NB; ++e)
                if ((d % NB) == e) zsel = zz[e];
              st[j][t] = zsel * pstd[t];
            }
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int d = d0 + j;
#pragma unroll
      for (int t = 0; t < TM; ++t) st[j][t] = (d < D && valid[t]) ? A.p[d * A.p_ld + pbase + t] : T(0);
    }
  }
#pragma unroll
  for (int j = 0; j < TN; ++j)
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      Kpart[t] += st[j][t] * st[j][t];
      st[j][t] *= inv_m[t];  // v = p / m
    }

  __syncthreads();

  // ---- the gradient GEMM: acc[j][t] = sum_k Lambda[d0+j][k] * x[k][particle t] ---
  T acc[TN][TM];
  const T* xcol = Xs + lane * TM;
  const T* lrow = Ls + w * TNP;
  auto matvec = [&]() {
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int t = 0; t < TM; ++t) acc[j][t] = T(0);
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
      T x[TM], l[TNP];
      vec_to_arr<T>(*reinterpret_cast<const Vec*>(xcol + (size_t)k * PT), x);
#pragma unroll
      for (int c = 0; c < TNP / VW; ++c)
        vec_to_arr<T>(*reinterpret_cast<const Vec*>(lrow + (size_t)k * (K2_WARPS * TNP) + c * VW), l + c * VW);
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int t = 0; t < TM; ++t) acc[j][t] = fma(l[j], x[t], acc[j][t]);
    }
  };
  // sum over owned dims of x * g (for U = 0.5 x.g)
  auto xdotg = [&](T (&out)[TM]) {
#pragma unroll
    for (int t = 0; t < TM; ++t) out[t] = T(0);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int d = d0 + j;
      if (d < D) {
        T x[TM];
        vec_to_arr<T>(*reinterpret_cast<const Vec*>(Xs + (size_t)d * PT + lane * TM), x);
#pragma unroll
        for (int t = 0; t < TM; ++t) out[t] += x[t] * acc[j][t];
      }
    }
  };
  // H[t] = 0.5 * (sum_w Kpart) / m + 0.5 * (sum_w xg): deterministic cross-warp sum
  auto reduce_H = [&](const T (&kp)[TM], const T (&xg)[TM], T (&H)[TM]) {
    T a[TM], b[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t) a[t] = kp[t];
    *reinterpret_cast<Vec*>(red + w * PT + lane * TM) = arr_to_vec(a);
    __syncthreads();
#pragma unroll
    for (int t = 0; t < TM; ++t) a[t] = T(0);
#pragma unroll
    for (int ww = 0; ww < K2_WARPS; ++ww) {
      vec_to_arr<T>(*reinterpret_cast<const Vec*>(red + ww * PT + lane * TM), b);
#pragma unroll
      for (int t = 0; t < TM; ++t) a[t] += b[t];
    }
    __syncthreads();
    *reinterpret_cast<Vec*>(red + w * PT + lane * TM) = arr_to_vec(xg);
    __syncthreads();
    T c[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t) c[t] = T(0);
#pragma unroll
    for (int ww = 0; ww < K2_WARPS; ++ww) {
      vec_to_arr<T>(*reinterpret_cast<const Vec*>(red + ww * PT + lane * TM), b);
#pragma unroll
      for (int t = 0; t < TM; ++t) c[t] += b[t];
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < TM; ++t) H[t] = T(0.5) * a[t] * inv_m[t] + T(0.5) * c[t];
  };

  matvec();  // g0
  T oldH[TM], newH[TM];
  if (hmc) {
    T xg[TM];
    xdotg(xg);
    reduce_H(Kpart, xg, oldH);
  }

  const T h = A.h, h2 = A.h2;
  const int L = A.L;
  if (integ == INTEG_LEAPFROG) {
    // half kick
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int t = 0; t < TM; ++t) st[j][t] -= (T(0.5) * h * inv_m[t]) * acc[j][t];
    for (int s = 0; s < L; ++s) {
      __syncthreads();  // every warp finished reading Xs in the previous matvec
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int d = d0 + j;
        if (d < D) {
          T x[TM];
          Vec* px = reinterpret_cast<Vec*>(Xs + (size_t)d * PT + lane * TM);
          vec_to_arr<T>(*px, x);
#pragma unroll
          for (int t = 0; t < TM; ++t) x[t] = fma(h, st[j][t], x[t]);  // drift
          *px = arr_to_vec(x);
        }
      }
      __syncthreads();
      matvec();
      const T kf = (s == L - 1) ? T(0.5) * h : h;  // full kick, half at the end
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int t = 0; t < TM; ++t) st[j][t] -= (kf * inv_m[t]) * acc[j][t];
    }
    if (L == 0) {  // undo the half kick: numSteps == 0 leaves (q, p) unchanged
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int t = 0; t < TM; ++t) st[j][t] += (T(0.5) * h * inv_m[t]) * acc[j][t];
    }
  } else {
    // Stormer-Verlet, src/integrator.py:142-163.  st: v -> qPast
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int d = d0 + j;
      if (d < D) {
        T x[TM];
        Vec* px = reinterpret_cast<Vec*>(Xs + (size_t)d * PT + lane * TM);
        vec_to_arr<T>(*px, x);
#pragma unroll
        for (int t = 0; t < TM; ++t) {
          const T xn = x[t] + st[j][t] * h + T(0.5) * (-acc[j][t] * inv_m[t]) * h2;
          st[j][t] = x[t];
          x[t] = xn;
        }
        *px = arr_to_vec(x);
      }
    }
    for (int s = 0; s < L; ++s) {
      __syncthreads();
      matvec();
      __syncthreads();
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int d = d0 + j;
        if (d < D) {
          T x[TM];
          Vec* px = reinterpret_cast<Vec*>(Xs + (size_t)d * PT + lane * TM);
          vec_to_arr<T>(*px, x);
#pragma unroll
          for (int t = 0; t < TM; ++t) {
            const T xn = T(2) * x[t] - st[j][t] + (-acc[j][t] * inv_m[t]) * h2;
            st[j][t] = x[t];
            x[t] = xn;
          }
          *px = arr_to_vec(x);
        }
      }
    }
    __syncthreads();
    // v = (q - qPast) / h
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int d = d0 + j;
      T x[TM];
#pragma unroll
      for (int t = 0; t < TM; ++t) x[t] = T(0);
      if (d < D) vec_to_arr<T>(*reinterpret_cast<const Vec*>(Xs + (size_t)d * PT + lane * TM), x);
#pragma unroll
      for (int t = 0; t < TM; ++t) st[j][t] = d < D ? (x[t] - st[j][t]) / h : T(0);
    }
    if (hmc) matvec();  // gradient at the final position, for U(q_new) = 0.5 x.g
  }

  // ---- p = v * m, new Hamiltonian ----------------------------------------------------
#pragma unroll
  for (int t = 0; t < TM; ++t) Kpart[t] = T(0);
#pragma unroll
  for (int j = 0; j < TN; ++j)
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      st[j][t] *= m[t];
      Kpart[t] += st[j][t] * st[j][t];
    }

  bool rej[TM];
  T accp[TM];
#pragma unroll
  for (int t = 0; t < TM; ++t) {
    rej[t] = false;
    accp[t] = T(1);
  }
  if (hmc) {
    T xg[TM];
    xdotg(xg);
    reduce_H(Kpart, xg, newH);
    const PhiloxKey K(A.seed, A.iter);
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      T u = T(0);
      if (valid[t]) u = A.u != nullptr ? A.u[pbase + t] : NormalBlock<T>::uniform(K, A.offset + (u64)(pbase + t), A.D);
      rej[t] = metropolis_reject<T>(oldH[t], newH[t], u, A.flags, &accp[t]);
    }
  }

  // ---- write back ------------------------------------------------------------------------
  bool any_rej = false, all_rej = true;
#pragma unroll
  for (int t = 0; t < TM; ++t) {
    any_rej |= rej[t];
    all_rej &= rej[t];
  }
  const bool pvec = full && A.p != nullptr && (A.p_ld % TM == 0) && ((reinterpret_cast<uintptr_t>(A.p) & 15) == 0);
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int d = d0 + j;
    if (d >= D) continue;
    const T mu = pa.mu[d];
    T x[TM], qold[TM];
    vec_to_arr<T>(*reinterpret_cast<const Vec*>(Xs + (size_t)d * PT + lane * TM), x);
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      x[t] += mu;
      qold[t] = T(0);
    }
    const bool need_old = any_rej && (A.partials != nullptr || (A.p != nullptr && (A.flags & FLAG_BUGCOMPAT)));
    if (need_old) {
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (valid[t] && rej[t]) qold[t] = A.q[d * A.q_ld + pbase + t];
    }
    // positions: HMC.py:175 -- rejected particles keep the value already in HBM
    if (qvec && !any_rej) {
      *reinterpret_cast<Vec*>(A.q + d * A.q_ld + pbase) = arr_to_vec(x);
    } else if (!all_rej) {
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (valid[t] && !rej[t]) A.q[d * A.q_ld + pbase + t] = x[t];
    }
    // momenta
    if (A.p != nullptr) {
      T pv[TM];
#pragma unroll
      for (int t = 0; t < TM; ++t) pv[t] = st[j][t];
      if (hmc && any_rej) {
#pragma unroll
        for (int t = 0; t < TM; ++t)
          if (rej[t]) {
            if (A.flags & FLAG_BUGCOMPAT) {
              pv[t] = qold[t];  // HMC.py:176 (sic)
            } else if (A.z != nullptr) {
              pv[t] = valid[t] ? A.z[d * A.z_ld + pbase + t] * pstd[t] : T(0);
            } else {
              pv[t] = one_normal<T>(A.seed, A.iter, A.offset + (u64)(pbase + t), d) * pstd[t];
            }
          }
      }
      if (pvec) {
        *reinterpret_cast<Vec*>(A.p + d * A.p_ld + pbase) = arr_to_vec(pv);
      } else {
#pragma unroll
        for (int t = 0; t < TM; ++t)
          if (valid[t]) A.p[d * A.p_ld + pbase + t] = pv[t];
      }
    }
    // statistics over the kept state: sum q_d, sum q_d^2 (each warp owns its dims)
    if (A.partials != nullptr) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (valid[t]) {
          const double qk = (double)(rej[t] ? qold[t] : x[t]);
          s1 += qk;
          s2 += qk * qk;
        }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        double* out = A.partials + (size_t)blockIdx.x * (2 * D + 3);
        out[3 + d] = s1;
        out[3 + D + d] = s2;
      }
    }
  }
  if (hmc && w == 0) {
    if (A.accept != nullptr) {
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (valid[t]) A.accept[pbase + t] = rej[t] ? 0 : 1;
    }
    if (A.partials != nullptr) {
      double na = 0.0, sa = 0.0, sh = 0.0;
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (valid[t]) {
          na += rej[t] ? 0.0 : 1.0;
          sa += (double)accp[t];
          sh += (double)(rej[t] ? oldH[t] : newH[t]);
        }
      na = warp_sum(na);
      sa = warp_sum(sa);
      sh = warp_sum(sh);
      if (lane == 0) {
        double* out = A.partials + (size_t)blockIdx.x * (2 * D + 3);
        out[0] = na;
        out[1] = sa;
        out[2] = sh;
      }
    }
  }
}

}  // namespace ehmc
