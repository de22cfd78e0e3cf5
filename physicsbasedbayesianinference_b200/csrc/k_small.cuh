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

#include <type_traits>

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
  static constexpr bool kPacked = true;
  __device__ __forceinline__ float grad2(const f32x2 (&Q)[(DT + 1) / 2], f32x2 (&G)[(DT + 1) / 2], bool wantE) const {
    f32x2 S = 0ull;
#pragma unroll
    for (int i = 0; i < (DT + 1) / 2; ++i) {
      G[i] = mul2(pk2((float)k[2 * i], 2 * i + 1 < DT ? (float)k[2 * i + 1] : 0.f), Q[i]);
      if (wantE) S = fma2(G[i], Q[i], S);
    }
    return 0.5f * (pk_lo(S) + pk_hi(S));
  }
};

// Packed float32 form of a potential functor (two dimensions per FFMA2 / FMUL2 issue slot), for the
// functors that define grad2; the others run the scalar form.
template <class Pot, typename = void>
struct HasPacked : std::false_type {};
template <class Pot>
struct HasPacked<Pot, std::void_t<decltype(Pot::kPacked)>> : std::true_type {};
template <class Pot, typename = void>
struct HasFusedKick : std::false_type {};
template <class Pot>
struct HasFusedKick<Pot, std::void_t<decltype(Pot::kFusedKick)>> : std::true_type {};

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
struct FunnelPot {  // Neal's funnel (see ehmc.h EHMC_FAMILY_FUNNEL), optionally in rescaled coordinates:
                    // v = a q[0], x_k = c q[k]:  U = inv_s2 q0^2 / 2 + 0.5 e^{lnb - a q0} sum q_k^2 + half_dm1 q0
  T inv_s2;         // a^2 / sigma_v^2
  T half_dm1;       // 0.5 a (D - 1) with the TRUE D
  T a;              // scale of v (1 for the plain funnel)
  T lnb;            // ln c^2   (0 for the plain funnel; a = 1, lnb = 0 reproduce the plain funnel bit for bit:
                    //           0 - 1 * v and 1 * hs are exact)
  __device__ __forceinline__ T grad(const T (&q)[DT], T (&g)[DT], bool wantE) const {
    const T v = q[0];
    const T ev = Ar<T>::exp_(Ar<T>::sub(lnb, Ar<T>::mul(a, v)));
    T s2 = T(0);
#pragma unroll
    for (int d = 1; d < DT; ++d) {
      s2 += q[d] * q[d];
      g[d] = ev * q[d];
    }
    const T hs = T(0.5) * ev * s2;
    const T ahs = a * hs;
    g[0] = v * inv_s2 - ahs + half_dm1;
    return T(0.5) * v * v * inv_s2 + hs + half_dm1 * v;
  }
  static constexpr bool kPacked = true;
  // float32 packed forms: e^{lnb - a v} is ONE FFMA + MUFU.EX2 (ex2.approx.ftz, relative error 2^-22, on the argument
  // scaled by log2 e) instead of expf's 9 instructions (range reduction by FFMA.SAT / FFMA.RM, two-term log2 e, scale
  // by shift): the trajectory loop is issue / FMA-pipe bound and this was a quarter of a leapfrog step.
  __device__ __forceinline__ float expo(float v) const {
    return ex2_ftz(fmaf(-(float)a * 1.4426950408889634f, v, (float)lnb * 1.4426950408889634f));
  }
  // pair 0 = (v, q_1); padded dims hold q = 0 and contribute nothing
  __device__ __forceinline__ float grad2(const f32x2 (&Q)[(DT + 1) / 2], f32x2 (&G)[(DT + 1) / 2], bool) const {
    const float v = pk_lo(Q[0]), q1 = pk_hi(Q[0]);
    const float ev = expo(v);
    const f32x2 ev2 = pk2(ev, ev);
    f32x2 S = 0ull;
#pragma unroll
    for (int i = 1; i < (DT + 1) / 2; ++i) {
      S = fma2(Q[i], Q[i], S);
      G[i] = mul2(Q[i], ev2);
    }
    const float s2 = fmaf(q1, q1, pk_lo(S) + pk_hi(S));
    const float hs = 0.5f * ev * s2;
    const float ahs = (float)a * hs;
    G[0] = pk2(v * (float)inv_s2 - ahs + (float)half_dm1, ev * q1);
    return 0.5f * v * v * (float)inv_s2 + hs + (float)half_dm1 * v;
  }
  // Gradient and kick in one: V += c * grad U(Q).  For the pairs above the first, grad = e^{-v} q, so the kick is
  // ONE FFMA2 with the scalar c e^{-v} instead of FMUL2 (gradient) + FFMA2 (kick); the first pair's
  // c (v / s^2 - a e^{-v} s2 / 2 + (D-1) a / 2) and c e^{-v} q_1 are formed with c folded into the coefficients
  // (c / s^2, c (D-1) a / 2: loop invariants), 6 scalar instructions.  Returns U(Q) when wantE.
  static constexpr bool kFusedKick = true;
  __device__ __forceinline__ float kick2(const f32x2 (&Q)[(DT + 1) / 2], f32x2 (&V)[(DT + 1) / 2], float c,
                                         bool wantE) const {
    const float v = pk_lo(Q[0]), q1 = pk_hi(Q[0]);
    const float ev = expo(v);
    const float cev = c * ev;
    const f32x2 cev2 = pk2(cev, cev);
    f32x2 S = 0ull;
#pragma unroll
    for (int i = 1; i < (DT + 1) / 2; ++i) {
      S = fma2(Q[i], Q[i], S);
      V[i] = fma2(Q[i], cev2, V[i]);
    }
    const float s2 = fmaf(q1, q1, pk_lo(S) + pk_hi(S));
    const float w = (0.5f * (float)a) * cev;                             // c a e^{-v} / 2
    const float t = fmaf(c * (float)inv_s2, v, c * (float)half_dm1);     // c (v / s^2 + (D-1) a / 2)
    V[0] = pk2(pk_lo(V[0]) + fmaf(-w, s2, t), fmaf(cev, q1, pk_hi(V[0])));
    return wantE ? fmaf(0.5f * ev, s2, fmaf(0.5f * v * (float)inv_s2, v, (float)half_dm1 * v)) : 0.f;
  }
};

template <typename T, int DT>
struct CoinPot {  // EHMC_FAMILY_COIN_TOSS: U = -sum k ln q + (n - k) ln(1 - q)
  T k[DT];        // successes (0 for padded dims)
  T nk[DT];       // failures n - k
  __device__ __forceinline__ T grad(const T (&q)[DT], T (&g)[DT], bool wantE) const {
    T e = T(0);
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      const T a = q[d], b = T(1) - q[d];
      // padded dims (k = nk = 0, q = 0) must stay exactly 0, not 0/0
      g[d] = (k[d] != T(0) ? -k[d] / a : T(0)) + (nk[d] != T(0) ? nk[d] / b : T(0));
      if (wantE) e -= (k[d] != T(0) ? k[d] * log(a) : T(0)) + (nk[d] != T(0) ? nk[d] * log(b) : T(0));
    }
    return e;
  }
};

// ---------------------------------------------------------------------------
// momentum draw: p = z * pstd, z fed or from Philox
// ---------------------------------------------------------------------------
// uword (optional, float32 Philox draws only): receives word z of normal block Dn / 4 -- the Metropolis uniform's
// word when NormalBlock<float>::uniform_in_normal_block(Dn) (stream version 2, common.cuh)
template <typename T, int DT>
__device__ __forceinline__ void draw_momentum(const IterArgs<T>& A, int Dn, long long i, T pstd, T (&p)[DT],
                                              uint32_t* uword = nullptr) {
  if (A.z != nullptr) {
#pragma unroll
    for (int d = 0; d < DT; ++d) p[d] = d < Dn ? Ar<T>::mul(A.z[d * A.z_ld + i], pstd) : T(0);
  } else {
    const PhiloxKey K(A.seed, A.iter);
    constexpr int NB = NormalBlock<T>::N;
#pragma unroll
    for (int b = 0; b < (DT + NB - 1) / NB; ++b) {
      T zz[NB];
      if constexpr (sizeof(T) == 4) {
        const uint4 r = K.block(A.offset + (u64)i, (uint32_t)b);
        NormalBlock<float>::transform(r, zz);
        if (uword != nullptr && b == (Dn >> 2)) *uword = r.z;
      } else {
        NormalBlock<T>::draw(K, A.offset + (u64)i, (uint32_t)b, zz);
      }
#pragma unroll
      for (int t = 0; t < NB; ++t) {
        const int d = b * NB + t;
        if (d < DT) p[d] = d < Dn ? Ar<T>::mul(zz[t], pstd) : T(0);
      }
    }
  }
}

// Metropolis uniform of particle i: fed, or the word kept from the momentum draw, or its own Philox block
template <typename T>
__device__ __forceinline__ T metropolis_uniform(const IterArgs<T>& A, int Dn, long long i, uint32_t uword) {
  if (A.u != nullptr) return A.u[i];
  if (A.z == nullptr && NormalBlock<T>::uniform_in_normal_block(Dn)) return NormalBlock<T>::uniform_from_word(uword);
  return NormalBlock<T>::uniform(PhiloxKey(A.seed, A.iter), A.offset + (u64)i, Dn);
}

// ---------------------------------------------------------------------------
// The integrators on register state.  On entry p holds the momentum; on exit q, p
// hold the integrated state.  Returns U(q_final) when wantE.  U0 (energy at the
// start) is returned through *U0 when wantE.
// ---------------------------------------------------------------------------
// float32 leapfrog on PACKED state (pairs of dimensions: FFMA2 / FMUL2 do two lanes of work per issue slot and this
// loop is issue bound): kick-drift-kick, Q positions, V velocities, hm = h / m.  The last step is peeled so that the
// loop body carries no select of the kick coefficient and no energy.
template <int DT, class Pot>
__device__ __forceinline__ float integrate_packed(const Pot& pot, f32x2 (&Q)[(DT + 1) / 2], f32x2 (&V)[(DT + 1) / 2],
                                                  float hm, float h, int L, bool wantE, float* U0) {
  constexpr int NP2 = (DT + 1) / 2;
  const float hmh = 0.5f * hm;
  const f32x2 h2p = pk2(h, h);
  f32x2 G[NP2];
  if (L <= 0) {  // L = 0: nothing moves
    *U0 = pot.grad2(Q, G, wantE);
    return *U0;
  }
  if constexpr (HasFusedKick<Pot>::value) {
    *U0 = pot.kick2(Q, V, -hmh, wantE);
    for (int j = 1; j < L; ++j) {
#pragma unroll
      for (int i = 0; i < NP2; ++i) Q[i] = fma2(V[i], h2p, Q[i]);
      pot.kick2(Q, V, -hm, false);
    }
#pragma unroll
    for (int i = 0; i < NP2; ++i) Q[i] = fma2(V[i], h2p, Q[i]);
    return pot.kick2(Q, V, -hmh, wantE);
  } else {
    const f32x2 nhm = pk2(-hm, -hm), nhmh = pk2(-hmh, -hmh);
    *U0 = pot.grad2(Q, G, wantE);
#pragma unroll
    for (int i = 0; i < NP2; ++i) V[i] = fma2(G[i], nhmh, V[i]);
    for (int j = 1; j < L; ++j) {
#pragma unroll
      for (int i = 0; i < NP2; ++i) Q[i] = fma2(V[i], h2p, Q[i]);
      pot.grad2(Q, G, false);
#pragma unroll
      for (int i = 0; i < NP2; ++i) V[i] = fma2(G[i], nhm, V[i]);
    }
#pragma unroll
    for (int i = 0; i < NP2; ++i) Q[i] = fma2(V[i], h2p, Q[i]);
    const float Uend = pot.grad2(Q, G, wantE);
#pragma unroll
    for (int i = 0; i < NP2; ++i) V[i] = fma2(G[i], nhmh, V[i]);
    return Uend;
  }
}

// One HMC trajectory of the float32 packed path from UNIT normals z (fed or drawn): velocities V = z pstd / m directly
// (no momentum in between), kinetic energies from sum z^2 and sum V^2 with packed FFMA2.  On exit q holds the end
// position, K0 / K1 / U0 / U1 the energies; the momentum p = V m is formed by the caller only when it is stored.
// Shared by k_small (all its grids) and k_small_run, whose results must agree bit for bit.
template <int DT, class Pot>
struct PackedHmc {
  static constexpr int NP2 = (DT + 1) / 2;
  f32x2 Q[NP2], V[NP2];
  float K0, K1, U0, U1;
  __device__ __forceinline__ void start(const float (&q)[DT], const float (&z)[DT], float m, float inv_m, float pstd) {
    const float sv = pstd * inv_m;
    const f32x2 sv2 = pk2(sv, sv);
    f32x2 S = 0ull;
#pragma unroll
    for (int i = 0; i < NP2; ++i) {
      Q[i] = pk2(q[2 * i], 2 * i + 1 < DT ? q[2 * i + 1] : 0.f);
      const f32x2 Z = pk2(z[2 * i], 2 * i + 1 < DT ? z[2 * i + 1] : 0.f);
      S = fma2(Z, Z, S);
      V[i] = mul2(Z, sv2);
    }
    // 0.5 dot(p, p) / m with p = z pstd                       HMC.py:109
    K0 = (0.5f * pstd) * sv * (pk_lo(S) + pk_hi(S));
  }
  __device__ __forceinline__ void run(const Pot& pot, float m, float inv_m, float h, int L) {
    U1 = integrate_packed<DT, Pot>(pot, Q, V, h * inv_m, h, L, true, &U0);
    f32x2 S = 0ull;
#pragma unroll
    for (int i = 0; i < NP2; ++i) S = fma2(V[i], V[i], S);
    K1 = (0.5f * m) * (pk_lo(S) + pk_hi(S));  // dot(-p,-p) == dot(p,p), HMC.py:164
  }
  __device__ __forceinline__ void position(float (&q)[DT]) const {
#pragma unroll
    for (int i = 0; i < NP2; ++i) {
      q[2 * i] = pk_lo(Q[i]);
      if (2 * i + 1 < DT) q[2 * i + 1] = pk_hi(Q[i]);
    }
  }
  __device__ __forceinline__ void momentum(float (&p)[DT], float m) const {
#pragma unroll
    for (int i = 0; i < NP2; ++i) {
      p[2 * i] = pk_lo(V[i]) * m;
      if (2 * i + 1 < DT) p[2 * i + 1] = pk_hi(V[i]) * m;
    }
  }
};

template <typename T, int DT, class Pot, int INTEG>
__device__ __forceinline__ T integrate_regs(const Pot& pot, T (&q)[DT], T (&p)[DT], T m, T h, T h2, int L,
                                            bool wantE, T* U0) {
  typedef Ar<T> R;
  const T inv_m = R::rcp_(m);
  if constexpr (sizeof(T) == 4 && INTEG == INTEG_LEAPFROG && HasPacked<Pot>::value) {
    constexpr int NP2 = (DT + 1) / 2;
    f32x2 Q[NP2], V[NP2];
#pragma unroll
    for (int i = 0; i < NP2; ++i) {
      Q[i] = pk2(q[2 * i], 2 * i + 1 < DT ? q[2 * i + 1] : 0.f);
      V[i] = pk2(p[2 * i] * inv_m, 2 * i + 1 < DT ? p[2 * i + 1] * inv_m : 0.f);
    }
    const T Uend = integrate_packed<DT, Pot>(pot, Q, V, h * inv_m, h, L, wantE, U0);
#pragma unroll
    for (int i = 0; i < NP2; ++i) {
      q[2 * i] = pk_lo(Q[i]);
      p[2 * i] = pk_lo(V[i]) * m;
      if (2 * i + 1 < DT) {
        q[2 * i + 1] = pk_hi(Q[i]);
        p[2 * i + 1] = pk_hi(V[i]) * m;
      }
    }
    return Uend;
  } else if constexpr (sizeof(T) == 4 && INTEG == INTEG_LEAPFROG) {
    // float32: the same recurrence in kick-drift-kick form (v_half = v + a h / 2 ; q += h v_half ;
    // v = v_half + a' h / 2, consecutive half kicks merged): 2 FFMA per dimension and step instead of
    // the 8 operations of the reference's expression order, which only the float64 mode reproduces
    // bit for bit.  Differs from it by rounding only (tolerance 1e-5, north_star).
    T g[DT];
    T (&v)[DT] = p;
#pragma unroll
    for (int d = 0; d < DT; ++d) v[d] *= inv_m;
    *U0 = pot.grad(q, g, wantE);
    T Uend = *U0;
    const T hm = h * inv_m, hmh = T(0.5) * hm;
    if (L > 0) {
#pragma unroll
      for (int d = 0; d < DT; ++d) v[d] = fmaf(-hmh, g[d], v[d]);
    }
    for (int j = 0; j < L; ++j) {
#pragma unroll
      for (int d = 0; d < DT; ++d) q[d] = fmaf(h, v[d], q[d]);
      const bool last = j == L - 1;
      Uend = pot.grad(q, g, wantE && last);
      const T ck = last ? hmh : hm;
#pragma unroll
      for (int d = 0; d < DT; ++d) v[d] = fmaf(-ck, g[d], v[d]);
    }
#pragma unroll
    for (int d = 0; d < DT; ++d) p[d] = v[d] * m;
    return Uend;
  }
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

// Ensemble statistics {n_accept, sum acceptance probability, sum H, sum q_d, sum q_d^2}: one row of 2D+3 float64
// partial sums per CTA (or per slice of the fused ensemble run).
//
// Every warp hands the DT+3 values of its 32 particles {accepted, acceptance probability, H, q_0 .. q_DT-1} through a
// warp-private shared-memory tile in the state's own precision; lane j then owns value j: it reads the 32 entries of
// its tile row with 16-byte loads, adds them and their squares in the state's precision (fixed order, four partial
// sums), and adds the two results to its float64 REGISTER accumulators.  So a statistic is a float64 sum of
// per-warp sums over 32 consecutive particles; in float32 mode those inner sums are float32 (relative error of a
// statistic over 2^22 particles ~1e-9, far below its Monte-Carlo error, and independent of timing and -- for shards
// that are multiples of 32 particles -- of the number of GPUs); in float64 mode everything is float64.
// Measured on B200 (profiles/microbench/op_rates_r02.txt): the scheme this replaced -- 2DT+3 float64 accumulators
// per THREAD in shared memory -- moved 16 B through shared memory per particle and statistic, 92 cycles of the SM's
// 128 B/clk per warp of particles: 40 us per iteration of 2^22 particles whatever the trajectory length, a fifth of
// config 5.  Summing exactly in float64 from the tile was no better (F2F.F64.F32 issues at a quarter rate: 32
// conversions per lane and warp of particles).  This form moves 4 B in and 4 B out per value (29 cycles) and converts
// twice per lane.
template <typename T, int DT>
struct WarpStats {
  static constexpr int NV = DT + 3;                    // values per particle
  static constexpr int NA = 2 * DT + 3;                // statistics columns
  static constexpr int LD = 32 + 16 / (int)sizeof(T);  // tile row: 32 particles + 16 bytes (conflict-free 16-byte reads)
  static constexpr int NSLOT = (NV + 31) / 32;         // values per lane
  static constexpr int PER16 = 16 / (int)sizeof(T);
  static constexpr size_t kTileBytes = sizeof(T) * NV * LD;                             // per warp
  static constexpr size_t kSmemBytes = (K1_THREADS / 32) * (kTileBytes + sizeof(double) * NA);  // tiles | warp rows
  double sum[NSLOT], sq[NSLOT];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int n = 0; n < NSLOT; ++n) sum[n] = sq[n] = 0.0;
  }
  // all 32 lanes call this together; lanes without a particle (on = false) contribute zeros.  Afterwards lane j holds
  // in s[n], r[n] the sum and the sum of squares of value v = j + 32 n over the warp's 32 particles.
  static __device__ __forceinline__ void reduce32(unsigned char* smem, bool on, T accepted, T accp, T H, const T (&q)[DT],
                                                  T (&s)[NSLOT], T (&r)[NSLOT]) {
    const int lane = threadIdx.x & 31;
    T* tile = reinterpret_cast<T*>(smem + (threadIdx.x >> 5) * kTileBytes);
    __syncwarp();  // the reads of the previous batch are done
    tile[0 * LD + lane] = on ? accepted : T(0);
    tile[1 * LD + lane] = on ? accp : T(0);
    tile[2 * LD + lane] = on ? H : T(0);
#pragma unroll
    for (int d = 0; d < DT; ++d) tile[(3 + d) * LD + lane] = on ? q[d] : T(0);
    __syncwarp();
#pragma unroll
    for (int n = 0; n < NSLOT; ++n) {
      const int v = lane + 32 * n;
      s[n] = r[n] = T(0);
      if (v < NV) {
        const T* row = tile + v * LD;
        T ps[PER16], pr[PER16];
#pragma unroll
        for (int t = 0; t < PER16; ++t) ps[t] = pr[t] = T(0);
        // (not unrolled further: with all 32 values of the row in registers at once the trajectory kernel loses
        // resident CTAs)
        if constexpr (PER16 == 4) {
          // float32: the four partial sums as two packed pairs (FADD2 / FFMA2: the same sums for half the issue slots)
          f32x2 s01 = 0ull, s23 = 0ull, r01 = 0ull, r23 = 0ull;
#pragma unroll 2
          for (int k = 0; k < 32; k += 4) {
            const uint4 x = *reinterpret_cast<const uint4*>(row + k);
            const f32x2 x01 = ((f32x2)x.y << 32) | x.x, x23 = ((f32x2)x.w << 32) | x.z;
            s01 = add2(s01, x01);
            s23 = add2(s23, x23);
            r01 = fma2(x01, x01, r01);
            r23 = fma2(x23, x23, r23);
          }
          ps[0] = pk_lo(s01), ps[1] = pk_hi(s01), ps[2] = pk_lo(s23), ps[3] = pk_hi(s23);
          pr[0] = pk_lo(r01), pr[1] = pk_hi(r01), pr[2] = pk_lo(r23), pr[3] = pk_hi(r23);
        } else {
#pragma unroll 2
          for (int k = 0; k < 32; k += PER16) {
            alignas(16) T x[PER16];
            *reinterpret_cast<uint4*>(x) = *reinterpret_cast<const uint4*>(row + k);
#pragma unroll
            for (int t = 0; t < PER16; ++t) {
              ps[t] += x[t];
              pr[t] = x[t] * x[t] + pr[t];
            }
          }
        }
        if constexpr (PER16 == 4) {
          s[n] = (ps[0] + ps[1]) + (ps[2] + ps[3]);
          r[n] = (pr[0] + pr[1]) + (pr[2] + pr[3]);
        } else {
          s[n] = ps[0] + ps[1];
          r[n] = pr[0] + pr[1];
        }
      }
    }
  }
  __device__ __forceinline__ void add(unsigned char* smem, bool on, T accepted, T accp, T H, const T (&q)[DT]) {
    T s[NSLOT], r[NSLOT];
    reduce32(smem, on, accepted, accp, H, q, s, r);
#pragma unroll
    for (int n = 0; n < NSLOT; ++n) {
      sum[n] += (double)s[n];
      sq[n] += (double)r[n];
    }
  }
  // The warp's 32 particles as ONE row of NA values in the state's precision, padded layout ([0..2], [3 .. 3+DT) sums,
  // [3+DT .. 3+2DT) squares): the fused ensemble run's per-batch rows.
  static __device__ __forceinline__ void batch_row(unsigned char* smem, T* out, bool on, T accepted, T accp, T H,
                                                   const T (&q)[DT]) {
    T s[NSLOT], r[NSLOT];
    reduce32(smem, on, accepted, accp, H, q, s, r);
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int n = 0; n < NSLOT; ++n) {
      const int v = lane + 32 * n;
      if (v < NV) {
        out[v] = s[n];
        if (v >= 3) out[DT + v] = r[n];
      }
    }
  }
  // the CTA's row: warp sums through shared memory, added in warp order; padded dimensions are dropped
  // (row layout with the true D: [0..2], [3 .. 3+D) sum q_d, [3+D .. 3+2D) sum q_d^2).  Contains block-wide barriers.
  __device__ __forceinline__ void store_row(unsigned char* smem, double* out_row, int D) const {
    constexpr int NWARP = K1_THREADS / 32;
    double* wrows = reinterpret_cast<double*>(smem + NWARP * kTileBytes);  // [NWARP][NA]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < NSLOT; ++n) {
      const int v = lane + 32 * n;
      if (v < NV) {
        wrows[w * NA + v] = sum[n];
        if (v >= 3) wrows[w * NA + DT + v] = sq[n];
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < NA; j += K1_THREADS) {
      double t = wrows[j];
#pragma unroll
      for (int k = 1; k < NWARP; ++k) t += wrows[k * NA + j];
      int o = j;
      if (j >= 3 + DT) {
        const int d = j - 3 - DT;
        o = d < D ? 3 + D + d : -1;
      } else if (j >= 3) {
        o = (j - 3) < D ? j : -1;
      }
      if (o >= 0) out_row[o] = t;
    }
  }
};

// ---------------------------------------------------------------------------
// Kernel.  HMC = false: Leapfrog/StormerVerlet.integrate on (q, p) in place.
//          HMC = true : one full iteration of HMC.getSamples' loop body.
// Persistent: the grid is one resident wave of CTAs and every thread walks particles
// i, i + gridDim * 128, ...  so that the ensemble statistics (2D+3 sums) are accumulated per
// thread (in shared memory) and reduced across the block ONCE per launch (a per-particle block reduction of 23
// doubles cost 3x the trajectory itself at config 5).
// ---------------------------------------------------------------------------
// EXACT: A.D == DT (no padded dimensions): the per-dimension bounds tests fold away.  With them the ten loads of
// config 5 cost 140 instructions (predicated address arithmetic, selects against zero) and -- worse -- the selects
// consumed the loaded values at once, so every warp waited for HBM BEFORE its ~300 instructions of Philox work
// instead of behind them (profiles/r01_k1_dbg_probe.txt, r01_ncu_full_k1l4_before.txt).
struct NoStepHook {};  // k_small_body: the step size is A.h

// WARP: the calling WARP runs particles 0 .. A.P - 1 (one batch of the fused ensemble run, 32 at a time) on its own --
// no block-wide barrier anywhere, A.partials holds one row of statistics per 32 particles in the state's precision
// (WarpStats::batch_row); otherwise the
// CTA is number blk of nblk CTAs walking the particles, 128 at a time, and A.partials holds one float64 row per CTA.
template <typename T, int DT, class Pot, int INTEG, bool HMC, bool EXACT, class StepHook = NoStepHook, bool WARP = false>
__device__ __forceinline__ void k_small_body(const IterArgs<T>& A, const Pot& pot, double* k1_smem, unsigned blk,
                                             unsigned nblk, const StepHook& hook = StepHook()) {
  // the fused ensemble run (k_small_ens) learns the step size of the iteration from `hook`, called once, behind the
  // first momentum draw (work that does not depend on h)
  T h = A.h, h2 = A.h2;
  bool h_known = std::is_same<StepHook, NoStepHook>::value;
  const int Dn = EXACT ? DT : A.D;
  const bool want_stats = HMC && A.partials != nullptr;
  unsigned char* stats_smem = reinterpret_cast<unsigned char*>(k1_smem);
  WarpStats<T, DT> ws;
  ws.clear();
  const long long stride = WARP ? 32 : (long long)nblk * K1_THREADS;
  const int t_in = WARP ? (threadIdx.x & 31) : threadIdx.x;

  for (long long base = WARP ? 0 : (long long)blk * K1_THREADS; base < A.P; base += stride) {
    const long long i = base + t_in;
    const bool active = i < A.P;
    const long long ic = active ? i : 0;  // inactive threads shadow particle 0, never store

    T q[DT], p[DT];
    const T m = A.mass[ic];
#pragma unroll
    for (int d = 0; d < DT; ++d) q[d] = EXACT ? A.q[d * A.q_ld + ic] : ld_if(A.q + (d * A.q_ld + ic), d < Dn);

    // float32 leapfrog with a packed potential: the trajectory runs on pairs of dimensions straight from the unit
    // normals (PackedHmc); every other combination goes through integrate_regs on (q, p)
    constexpr bool FAST = HMC && sizeof(T) == 4 && INTEG == INTEG_LEAPFROG && HasPacked<Pot>::value;
    T pstd = T(0);
    uint32_t uword = 0u;
    if (HMC) {
      // the unit normals first: ~300 instructions that depend on nothing in flight; the mass is first touched
      // after them (z * 1 is exact, so p = z * pstd below is the product the one-step form gave)
      draw_momentum<T, DT>(A, Dn, ic, T(1), p, &uword);
      pstd = momentum_std<T>(m, A.kB, A.temp, A.pscale);
      if constexpr (!FAST) {
#pragma unroll
        for (int d = 0; d < DT; ++d) p[d] = Ar<T>::mul(p[d], pstd);
      }
    } else {
#pragma unroll
      for (int d = 0; d < DT; ++d) p[d] = EXACT ? A.p[d * A.p_ld + ic] : ld_if(A.p + (d * A.p_ld + ic), d < Dn);
    }

    const T inv_m = Ar<T>::rcp_(m);
    if constexpr (!std::is_same<StepHook, NoStepHook>::value) {
      if (!h_known) {
        hook(h, h2);
        h_known = true;
      }
    }
    T oldH, newH;
    [[maybe_unused]] PackedHmc<DT, Pot> traj;
    if constexpr (FAST) {
      traj.start(q, p, m, inv_m, pstd);
      traj.run(pot, m, inv_m, h, A.L);
      traj.position(q);
      oldH = traj.K0 + traj.U0;
      newH = traj.K1 + traj.U1;
    } else {
      T K0 = T(0);
      if (HMC) K0 = kinetic<T, DT>(p, m, inv_m);
      T U0;
      const T U1 = integrate_regs<T, DT, Pot, INTEG>(pot, q, p, m, h, h2, A.L, HMC, &U0);

      if (!HMC) {
        if (active) {
#pragma unroll
          for (int d = 0; d < DT; ++d)
            if (d < Dn) {
              A.q[d * A.q_ld + i] = q[d];
              A.p[d * A.p_ld + i] = p[d];
            }
        }
        continue;
      }
      oldH = Ar<T>::add(K0, U0);
      newH = Ar<T>::add(kinetic<T, DT>(p, m, inv_m), U1);  // dot(-p,-p) == dot(p,p), HMC.py:164
    }
    const T u = metropolis_uniform<T>(A, Dn, ic, uword);
    T accp;
    const bool rej = metropolis_reject<T>(oldH, newH, u, A.flags, &accp);

    if (active) {
      if (!rej) {
#pragma unroll
        for (int d = 0; d < DT; ++d)
          if (d < Dn) A.q[d * A.q_ld + i] = q[d];  // HMC.py:175 (rejected: q in HBM is still oldQ)
      }
      if (A.p != nullptr) {
        if constexpr (FAST) traj.momentum(p, m);
        if (rej) {
          if (A.flags & FLAG_BUGCOMPAT) {
#pragma unroll
            for (int d = 0; d < DT; ++d) p[d] = d < Dn ? A.q[d * A.q_ld + i] : T(0);  // HMC.py:176 (sic)
          } else {
            draw_momentum<T, DT>(A, Dn, i, pstd, p);  // oldP
          }
        }
#pragma unroll
        for (int d = 0; d < DT; ++d)
          if (d < Dn) A.p[d * A.p_ld + i] = p[d];  // un-flipped, HMC.py:164,179
      }
      if (A.accept != nullptr) A.accept[i] = rej ? 0 : 1;
    }

    if (want_stats) {
      // kept state for the statistics (lanes without a particle add zeros)
      if (rej && active) {
#pragma unroll
        for (int d = 0; d < DT; ++d) q[d] = d < Dn ? A.q[d * A.q_ld + i] : T(0);
      }
      if constexpr (WARP)  // one row per 32 particles (sub-batch), consecutive
        WarpStats<T, DT>::batch_row(stats_smem, reinterpret_cast<T*>(A.partials) + (base >> 5) * (2 * DT + 3), active,
                                    rej ? T(0) : T(1), accp, rej ? oldH : newH, q);
      else
        ws.add(stats_smem, active, rej ? T(0) : T(1), accp, rej ? oldH : newH, q);
    }
  }
  if constexpr (!WARP) {
    if (want_stats) ws.store_row(stats_smem, A.partials + (size_t)blk * (2 * Dn + 3), Dn);
  }
}

// Two kernels (picked on the host by D == DT) rather than one with both bodies: a kernel gets the register
// count of its hungrier body.  (Capping the float32 HMC kernel at 64 registers for 8 CTAs per SM instead of 7
// measured 7 % slower at config 5: the Philox rounds of the three blocks no longer interleave.)
template <typename T, int DT, class Pot, int INTEG, bool HMC, bool EXACT>
__global__ void __launch_bounds__(K1_THREADS) k_small(const IterArgs<T> Ain, const Pot pot) {
  extern __shared__ __align__(16) double k1_smem[];
  const IterArgs<T> A = resolve_dynamic(Ain);
  k_small_body<T, DT, Pot, INTEG, HMC, EXACT>(A, pot, k1_smem, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------------------
// HMC.getSamples' WHOLE loop (src/HMC.py:150-179) as one launch: every thread keeps its particle in
// registers across `nIter` iterations (momentum refresh from Philox iteration A.iter + it, trajectory,
// Metropolis, restore) and stores each iteration's position / momentum into the reference's own
// (D, P, S) sample arrays.  For the launch-bound regime of the reference's runnable configurations
// (config 1: P = 1024, 1000 iterations: one launch instead of 1000 launches + 2000 copies).
// ---------------------------------------------------------------------------
template <typename T, int DT, class Pot, int INTEG>
__global__ void __launch_bounds__(K1_THREADS) k_small_run(const IterArgs<T> Ain, const Pot pot, const RunArgs<T> R) {
  const long long stride = (long long)gridDim.x * K1_THREADS;
  for (long long i = (long long)blockIdx.x * K1_THREADS + threadIdx.x; i < Ain.P; i += stride) {
    IterArgs<T> A = Ain;
    T q[DT], qold[DT], p[DT], p0[DT];
    const T m = A.mass[i], inv_m = Ar<T>::rcp_(m);
    const T pstd = momentum_std<T>(m, A.kB, A.temp, A.pscale);
#pragma unroll
    for (int d = 0; d < DT; ++d) q[d] = d < A.D ? A.q[d * A.q_ld + i] : T(0);
    int nacc = 0;
    constexpr bool FAST = sizeof(T) == 4 && INTEG == INTEG_LEAPFROG && HasPacked<Pot>::value;  // as in k_small_body
    for (int it = 0; it < R.nIter; ++it) {
      A.iter = Ain.iter + (u64)it;
      uint32_t uword = 0u;
      draw_momentum<T, DT>(A, A.D, i, FAST ? T(1) : pstd, p, &uword);
#pragma unroll
      for (int d = 0; d < DT; ++d) {
        qold[d] = q[d];
        p0[d] = FAST ? p[d] * pstd : p[d];
      }
      T oldH, newH;
      if constexpr (FAST) {
        PackedHmc<DT, Pot> traj;
        traj.start(q, p, m, inv_m, pstd);
        traj.run(pot, m, inv_m, A.h, A.L);
        traj.position(q);
        traj.momentum(p, m);
        oldH = traj.K0 + traj.U0;
        newH = traj.K1 + traj.U1;
      } else {
        const T K0 = kinetic<T, DT>(p, m, inv_m);
        T U0;
        const T U1 = integrate_regs<T, DT, Pot, INTEG>(pot, q, p, m, A.h, A.h2, A.L, true, &U0);
        oldH = Ar<T>::add(K0, U0);
        newH = Ar<T>::add(kinetic<T, DT>(p, m, inv_m), U1);
      }
      const T u = metropolis_uniform<T>(A, A.D, i, uword);
      T accp;
      const bool rej = metropolis_reject<T>(oldH, newH, u, A.flags, &accp);
      nacc += rej ? 0 : 1;
#pragma unroll
      for (int d = 0; d < DT; ++d) {
        if (rej) {
          q[d] = qold[d];                                         // HMC.py:175
          p[d] = (A.flags & FLAG_BUGCOMPAT) ? qold[d] : p0[d];    // HMC.py:176 (sic) / the old momentum
        }
        if (d < A.D) {
          const long long o = ((long long)d * A.P + i) * R.S + R.s0 + it;
          if (R.samples != nullptr) R.samples[o] = q[d];          // HMC.py:178
          if (R.momenta != nullptr) R.momenta[o] = p[d];          // HMC.py:179 (un-flipped, :164)
        }
      }
    }
#pragma unroll
    for (int d = 0; d < DT; ++d)
      if (d < A.D) {
        A.q[d * A.q_ld + i] = q[d];
        if (A.p != nullptr) A.p[d * A.p_ld + i] = p[d];
      }
    if (R.accepted != nullptr) R.accepted[i] = nacc;
  }
}

}  // namespace ehmc
