// K1: fused whole-trajectory kernel for small D (one thread = one particle, the
// phase space q, v, a held in registers for all L steps).
//
// Replaces, per particle, the reference's
//   Ensemble.setMomentum           src/ensemble.py:88-91
//   Leapfrog.integrate loop body   src/integrator.py:105-120
//   StormerVerlet.integrate        src/integrator.py:142-163
//   HMC.getWeightsRatio            src/HMC.py:106-116
//   accept / restore               src/HMC.py:168-176
// HBM traffic per particle-iteration: read q (D) + mass, write q (D) where accepted;
// everything else stays on chip.  Loads/stores are particle-contiguous (SoA), so a
// warp touches 128 B per dimension.
#pragma once

#include "common.cuh"

namespace ehmc {

constexpr int K1_THREADS = 128;

// ---------------------------------------------------------------------------
// Potential functors.  grad(q, g, wantE) writes grad U into g and returns U when
// wantE (else an unspecified value).  Passed by value (constant bank).
// ---------------------------------------------------------------------------
template <typename T, int DT>
struct DiagPot {  // harmonicPotentialND, src/potential.py:27:  0.5 * dot(k, q**2)
  T k[DT];
  __device__ __forceinline__ T grad(const T (&q)[DT], T (&g)[DT], bool wantE) const {
    T s = T(0);
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      g[d] = Ar<T>::mul(k[d], q[d]);
      if (wantE) s = Ar<T>::add(s, Ar<T>::mul(k[d], Ar<T>::mul(q[d], q[d])));
    }
    return Ar<T>::mul(T(0.5), s);
  }
};

template <typename T, int DT>
struct DenseSmallPot {  // U = 0.5 x^T Lambda x, x = q - mu;  grad = Lambda x
  T lam[DT * DT];       // row-major, zero padded
  T mu[DT];
  __device__ __forceinline__ T grad(const T (&q)[DT], T (&g)[DT], bool wantE) const {
    T x[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) x[d] = q[d] - mu[d];
    T e = T(0);
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      T s = T(0);
#pragma unroll
      for (int j = 0; j < DT; ++j) s += lam[i * DT + j] * x[j];
      g[i] = s;
      e += x[i] * s;
    }
    return T(0.5) * e;
  }
};

template <typename T, int DT>
struct FunnelPot {  // Neal's funnel (see ehmc.h EHMC_FAMILY_FUNNEL)
  T inv_s2;         // 1 / sigma_v^2
  T half_dm1;       // 0.5 (D - 1) with the TRUE D
  __device__ __forceinline__ T grad(const T (&q)[DT], T (&g)[DT], bool wantE) const {
    const T v = q[0];
    const T ev = Ar<T>::exp_(-v);
    T s2 = T(0);
#pragma unroll
    for (int d = 1; d < DT; ++d) {
      s2 += q[d] * q[d];
      g[d] = ev * q[d];
    }
    const T hs = T(0.5) * ev * s2;
    g[0] = v * inv_s2 - hs + half_dm1;
    return T(0.5) * v * v * inv_s2 + hs + half_dm1 * v;
  }
};

// ---------------------------------------------------------------------------
// momentum draw: p = z * pstd, z fed or from Philox
// ---------------------------------------------------------------------------
template <typename T, int DT>
__device__ __forceinline__ void draw_momentum(const IterArgs<T>& A, long long i, T pstd, T (&p)[DT]) {
  if (A.z != nullptr) {
#pragma unroll
    for (int d = 0; d < DT; ++d) p[d] = d < A.D ? Ar<T>::mul(A.z[d * A.z_ld + i], pstd) : T(0);
  } else {
    const PhiloxKey K(A.seed, A.iter);
    constexpr int NB = NormalBlock<T>::N;
#pragma unroll
    for (int b = 0; b < (DT + NB - 1) / NB; ++b) {
      T zz[NB];
      NormalBlock<T>::draw(K, A.offset + (u64)i, (uint32_t)b, zz);
#pragma unroll
      for (int t = 0; t < NB; ++t) {
        const int d = b * NB + t;
        if (d < DT) p[d] = d < A.D ? Ar<T>::mul(zz[t], pstd) : T(0);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// The integrators on register state.  On entry p holds the momentum; on exit q, p
// hold the integrated state.  Returns U(q_final) when wantE.  U0 (energy at the
// start) is returned through *U0 when wantE.
// ---------------------------------------------------------------------------
template <typename T, int DT, class Pot, int INTEG>
__device__ __forceinline__ T integrate_regs(const Pot& pot, T (&q)[DT], T (&p)[DT], T m, T h, T h2, int L,
                                            bool wantE, T* U0) {
  typedef Ar<T> R;
  const T inv_m = T(1) / m;
  T a[DT], g[DT];
  T Uend;
  // v = p / m                                   integrator.py:106 / :143
#pragma unroll
  for (int d = 0; d < DT; ++d) p[d] = R::divm(p[d], m, inv_m);
  T (&v)[DT] = p;
  *U0 = pot.grad(q, g, wantE);
  Uend = *U0;
  // a = -gradient(q) / m                        integrator.py:73
#pragma unroll
  for (int d = 0; d < DT; ++d) a[d] = R::divm(-g[d], m, inv_m);

  if (INTEG == INTEG_LEAPFROG) {
    for (int j = 0; j < L; ++j) {
      // q += v*h + 0.5*a*h**2                   integrator.py:112-115
#pragma unroll
      for (int d = 0; d < DT; ++d)
        q[d] = R::add(q[d], R::add(R::mul(v[d], h), R::mul(R::mul(T(0.5), a[d]), h2)));
      Uend = pot.grad(q, g, wantE && (j == L - 1));
      // v += 0.5*(a + a')*h ; a = a'            integrator.py:116-118
#pragma unroll
      for (int d = 0; d < DT; ++d) {
        const T a2 = R::divm(-g[d], m, inv_m);
        v[d] = R::add(v[d], R::mul(R::mul(T(0.5), R::add(a[d], a2)), h));
        a[d] = a2;
      }
    }
  } else {
    T qp[DT];
    // qPast = q ; q = q + v*h + 0.5*a*h**2      integrator.py:145-150
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      qp[d] = q[d];
      q[d] = R::add(R::add(q[d], R::mul(v[d], h)), R::mul(R::mul(T(0.5), a[d]), h2));
    }
    for (int j = 0; j < L; ++j) {
      pot.grad(q, g, false);
      // q = 2*q - qPast + a*h**2                integrator.py:155-160
#pragma unroll
      for (int d = 0; d < DT; ++d) {
        const T t = q[d];
        q[d] = R::add(R::sub(R::mul(T(2), q[d]), qp[d]), R::mul(R::divm(-g[d], m, inv_m), h2));
        qp[d] = t;
      }
    }
    // v = (q - qPast) / h                       integrator.py:162
#pragma unroll
    for (int d = 0; d < DT; ++d) v[d] = R::sub(q[d], qp[d]) / h;
    if (wantE) Uend = pot.grad(q, g, true);
  }
  // p = v * m                                   integrator.py:120 / :163
#pragma unroll
  for (int d = 0; d < DT; ++d) p[d] = R::mul(v[d], m);
  return Uend;
}

template <typename T, int DT>
__device__ __forceinline__ T kinetic(const T (&p)[DT], T m, T inv_m) {
  // 0.5 * dot(p, p) / m                         HMC.py:109
  T s = T(0);
#pragma unroll
  for (int d = 0; d < DT; ++d) s = Ar<T>::add(s, Ar<T>::mul(p[d], p[d]));
  return Ar<T>::divm(Ar<T>::mul(T(0.5), s), m, inv_m);
}

// Block-level partial sums of NS values -> partials[blockIdx.x * NS + j].
// Each thread contributes vals via the callback `get(j)`.
template <int NTHREADS, class F>
__device__ __forceinline__ void block_partials(double* out, int NS, double* smem /*[NTHREADS/32][NS]*/, F get) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int j = 0; j < NS; ++j) {
    const double s = warp_sum(get(j));
    if (lane == 0) smem[w * NS + j] = s;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < NS; j += NTHREADS) {
    double s = 0.0;
#pragma unroll
    for (int ww = 0; ww < NTHREADS / 32; ++ww) s += smem[ww * NS + j];
    out[(size_t)blockIdx.x * NS + j] = s;
  }
}

// ---------------------------------------------------------------------------
// Kernel.  HMC = false: Leapfrog/StormerVerlet.integrate on (q, p) in place.
//          HMC = true : one full iteration of HMC.getSamples' loop body.
// ---------------------------------------------------------------------------
template <typename T, int DT, class Pot, int INTEG, bool HMC>
__global__ void __launch_bounds__(K1_THREADS) k_small(const IterArgs<T> A, const Pot pot) {
  extern __shared__ double k1_smem[];
  const long long i = (long long)blockIdx.x * K1_THREADS + threadIdx.x;
  const bool active = i < A.P;
  const long long ic = active ? i : 0;  // inactive threads shadow particle 0, never store

  T q[DT], p[DT];
  const T m = A.mass[ic];
#pragma unroll
  for (int d = 0; d < DT; ++d) q[d] = d < A.D ? A.q[d * A.q_ld + ic] : T(0);

  T pstd = T(0);
  if (HMC) {
    pstd = momentum_std<T>(m, A.kB, A.temp, A.pscale);
    draw_momentum<T, DT>(A, ic, pstd, p);
  } else {
#pragma unroll
    for (int d = 0; d < DT; ++d) p[d] = d < A.D ? A.p[d * A.p_ld + ic] : T(0);
  }

  T K0 = T(0);
  const T inv_m = T(1) / m;
  if (HMC) K0 = kinetic<T, DT>(p, m, inv_m);
  T U0;
  const T U1 = integrate_regs<T, DT, Pot, INTEG>(pot, q, p, m, A.h, A.h2, A.L, HMC, &U0);

  if (!HMC) {
    if (active) {
#pragma unroll
      for (int d = 0; d < DT; ++d)
        if (d < A.D) {
          A.q[d * A.q_ld + i] = q[d];
          A.p[d * A.p_ld + i] = p[d];
        }
    }
    return;
  }

  const T oldH = Ar<T>::add(K0, U0);
  const T newH = Ar<T>::add(kinetic<T, DT>(p, m, inv_m), U1);  // dot(-p,-p) == dot(p,p), HMC.py:164
  T u;
  if (A.u != nullptr)
    u = A.u[ic];
  else
    u = NormalBlock<T>::uniform(PhiloxKey(A.seed, A.iter), A.offset + (u64)ic);
  T accp;
  const bool rej = metropolis_reject<T>(oldH, newH, u, A.flags, &accp);

  if (active) {
    if (!rej) {
#pragma unroll
      for (int d = 0; d < DT; ++d)
        if (d < A.D) A.q[d * A.q_ld + i] = q[d];  // HMC.py:175 (rejected: q in HBM is still oldQ)
    }
    if (A.p != nullptr) {
      if (rej) {
        if (A.flags & FLAG_BUGCOMPAT) {
#pragma unroll
          for (int d = 0; d < DT; ++d) p[d] = d < A.D ? A.q[d * A.q_ld + i] : T(0);  // HMC.py:176 (sic)
        } else {
          draw_momentum<T, DT>(A, i, pstd, p);  // oldP
        }
      }
#pragma unroll
      for (int d = 0; d < DT; ++d)
        if (d < A.D) A.p[d * A.p_ld + i] = p[d];  // un-flipped, HMC.py:164,179
    }
    if (A.accept != nullptr) A.accept[i] = rej ? 0 : 1;
  }

  if (A.partials != nullptr) {
    // kept state for the statistics
    if (rej) {
#pragma unroll
      for (int d = 0; d < DT; ++d) q[d] = d < A.D ? A.q[d * A.q_ld + ic] : T(0);
    }
    const double w = active ? 1.0 : 0.0;
    const double hk = (double)(rej ? oldH : newH);
    const int D = A.D;
    block_partials<K1_THREADS>(A.partials, 2 * D + 3, k1_smem, [&](int j) -> double {
      if (j == 0) return w * (rej ? 0.0 : 1.0);
      if (j == 1) return w * (double)accp;
      if (j == 2) return w * hk;
      const int d = (j - 3) % D;
      double qd = 0.0;
#pragma unroll
      for (int dd = 0; dd < DT; ++dd)
        if (dd == d) qd = (double)q[dd];
      return w * ((j - 3) < D ? qd : qd * qd);
    });
  }
}

}  // namespace ehmc
