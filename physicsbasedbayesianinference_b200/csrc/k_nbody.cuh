// K3: fused whole-trajectory kernel for the pairwise gravitational family
// (BASELINE config 4: every ensemble particle is a B-body system, D = 3 B).
//
//   U(q) = -G sum_{i<j} m_i m_j / sqrt(|r_i - r_j|^2 + eps^2)
//   dU/dr_i = G m_i sum_{j != i} m_j (r_i - r_j) / (|r_i - r_j|^2 + eps^2)^{3/2}
//   (for eps = 0:  -dU/dr_i / m_i = getAccelNBody(q, m, i), src/potential.py:30-53;
//    potential sign as samples/NBody/MiscFunctions.py:163-169)
//
// One CTA per ensemble particle.  The B bodies' positions (x, y, z, m) live in shared
// memory as 16/32-byte vectors; every thread owns 4 or 8 bodies (positions, velocities
// and force accumulators in registers) and sweeps all j through shared-memory
// broadcasts: one LDS.128 feeds 8 pair interactions (~13 issue slots each, one MUFU.RSQ).
// Per trajectory the all-pairs sweep runs L+1 times without touching HBM; HBM traffic is
// the 3B coordinates in and out once per HMC iteration.
// Coordinates are flattened component-major, d = c*B + b (src/potential.py:83-84).
#pragma once

#include <type_traits>

#include "common.cuh"
#include "k_dense.cuh"  // one_normal

namespace ehmc {

constexpr int NB_TI_MAX = 8;  // bodies per thread: 8, or 4 when that still covers B with <= 1024 threads

template <typename T>
struct NBodyArgs {
  const T* bmass;  // [B]
  int B;
  T G;
  T eps2;
  // endpoint cache (EHMC_FLAG_REUSE_ENDPOINT): f = sum_j m_j (r_j - r_i) / r^3 per coordinate, [3B][P] like q,
  // and U per particle, at the position the previous iteration kept
  T* fcache;
  T* ucache;
  int cache_read;  // 1: start from the cache instead of the first all-pairs sweep
};

template <typename T>
struct V4;
template <>
struct V4<float> {
  typedef float4 type;
  static __device__ __forceinline__ float4 make(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
};
template <>
struct V4<double> {
  typedef double4 type;
  static __device__ __forceinline__ double4 make(double x, double y, double z, double w) {
    return make_double4(x, y, z, w);
  }
};

// bare MUFU.RSQ: rsqrtf() wraps it in a denormal-input rescue (FSETP + two predicated FMULs per call),
// which squared distances never need
__device__ __forceinline__ float rsqrt_ftz(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// shared-memory record of one body: double (x, y, z, m); float (x, x, y, y)(z, z, m, m) -- every
// component duplicated so that one LDS.128 delivers two ready-made packed operands
template <typename T>
struct BodyRec;
template <>
struct BodyRec<double> {
  static constexpr int VECS = 1;
  static __device__ __forceinline__ void store(double4* pos, int b, double x, double y, double z, double m) {
    pos[b] = make_double4(x, y, z, m);
  }
};
template <>
struct BodyRec<float> {
  static constexpr int VECS = 2;
  static __device__ __forceinline__ void store(float4* pos, int b, float x, float y, float z, float m) {
    pos[2 * b] = make_float4(x, x, y, y);
    pos[2 * b + 1] = make_float4(z, z, m, m);
  }
};

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red, int nwarps) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T s = T(0);
  for (int w = 0; w < nwarps; ++w) s += red[w];
  return s;
}

template <typename T, bool EPS0, int NTMAX, int NB_TI>
__global__ void __launch_bounds__(NTMAX) k_nbody(const IterArgs<T> A, const NBodyArgs<T> pa, const int integ,
                                                const int hmc) {
  typedef typename V4<T>::type Vec;
  extern __shared__ __align__(32) unsigned char k3_smem_raw[];
  const int B = pa.B, NT = blockDim.x, tid = threadIdx.x, nwarps = NT >> 5;
  Vec* pos = reinterpret_cast<Vec*>(k3_smem_raw);                 // [B] body records (BodyRec)
  T* red = reinterpret_cast<T*>(pos + BodyRec<T>::VECS * B);      // [32] + 2 broadcast slots
  const long long part = blockIdx.x;
  const T M = A.mass[part], inv_M = T(1) / M;
  const T G = pa.G, eps2 = pa.eps2;

  // body of slot s: groups of 4 consecutive bodies per thread (= one Philox block per component)
  int body[NB_TI];
  bool own[NB_TI];
#pragma unroll
  for (int s = 0; s < NB_TI; ++s) {
    body[s] = 4 * tid + (s & 3) + 4 * NT * (s >> 2);
    own[s] = body[s] < B;
  }
  T x[NB_TI][3], st[NB_TI][3], bm[NB_TI];
#pragma unroll
  for (int s = 0; s < NB_TI; ++s) {
    bm[s] = own[s] ? pa.bmass[body[s]] : T(0);
#pragma unroll
    for (int c = 0; c < 3; ++c) x[s][c] = own[s] ? A.q[((long long)c * B + body[s]) * A.q_ld + part] : T(0);
    if (own[s]) BodyRec<T>::store(pos, body[s], x[s][0], x[s][1], x[s][2], bm[s]);
  }

  // ---- momentum ---------------------------------------------------------------------
  const T pstd = hmc ? momentum_std<T>(M, A.kB, A.temp, A.pscale) : T(0);
  auto draw = [&](T (&p)[NB_TI][3]) {
    if (A.z != nullptr) {
#pragma unroll
      for (int s = 0; s < NB_TI; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          p[s][c] = own[s] ? A.z[((long long)c * B + body[s]) * A.z_ld + part] * pstd : T(0);
    } else {
      constexpr int NBLK = NormalBlock<T>::N;
      const PhiloxKey K(A.seed, A.iter);
      if ((B % 4) == 0 && NBLK == 4) {
#pragma unroll
        for (int g = 0; g < NB_TI / 4; ++g)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            T zz[4] = {T(0), T(0), T(0), T(0)};
            if (own[4 * g]) NormalBlock<T>::draw(K, A.offset + (u64)part, (uint32_t)((c * B + body[4 * g]) / 4), zz);
#pragma unroll
            for (int e = 0; e < 4; ++e) p[4 * g + e][c] = own[4 * g + e] ? zz[e % NBLK] * pstd : T(0);
          }
      } else {
#pragma unroll 1
        for (int s = 0; s < NB_TI; ++s)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            T zv = T(0);
            if (own[s]) zv = one_normal<T>(A.seed, A.iter, A.offset + (u64)part, c * B + body[s]);
#pragma unroll
            for (int s2 = 0; s2 < NB_TI; ++s2)
              if (s2 == s) p[s2][c] = zv * pstd;
          }
      }
    }
  };
  if (hmc) {
    draw(st);
  } else {
#pragma unroll
    for (int s = 0; s < NB_TI; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c) st[s][c] = own[s] ? A.p[((long long)c * B + body[s]) * A.p_ld + part] : T(0);
  }
  T ksum = T(0);
#pragma unroll
  for (int s = 0; s < NB_TI; ++s)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ksum += st[s][c] * st[s][c];
      st[s][c] *= inv_M;  // v = p / M
    }

  // ---- all-pairs sweep: f[s] = sum_j m_j (r_j - r_i) / r^3 ; pot[s] = sum_j m_j / r -------
  T f[NB_TI][3], pot[NB_TI];
  auto sweep_scalar = [&](bool wantE) {
#pragma unroll
    for (int s = 0; s < NB_TI; ++s) {
      f[s][0] = f[s][1] = f[s][2] = T(0);
      pot[s] = T(0);
    }
    if constexpr (sizeof(T) == 8) {
#pragma unroll 2
      for (int j = 0; j < B; ++j) {
        const Vec pj = pos[j];
#pragma unroll
        for (int s = 0; s < NB_TI; ++s) {
          const T dx = pj.x - x[s][0], dy = pj.y - x[s][1], dz = pj.z - x[s][2];
          const T r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
          T inv = Ar<T>::rsqrt_(r2);
          if (EPS0) inv = r2 > T(0) ? inv : T(0);  // self pair / coincident bodies contribute nothing
          const T mi = pj.w * inv;
          const T w3 = mi * inv * inv;
          f[s][0] = fma(dx, w3, f[s][0]);
          f[s][1] = fma(dy, w3, f[s][1]);
          f[s][2] = fma(dz, w3, f[s][2]);
          if (wantE) pot[s] += mi;
        }
      }
    }
  };
  // float32: slots (2k, 2k+1) of a thread ride in one packed register pair
  auto sweep_packed = [&](auto wantE_tag) {
    constexpr bool wantE = decltype(wantE_tag)::value;
    if constexpr (sizeof(T) == 4) {
      constexpr int NP2 = NB_TI / 2;
      f32x2 X[NP2][3], F[NP2][3], PT[NP2];
      const f32x2 m1 = pk2(-1.f, -1.f), e2 = pk2(eps2, eps2);
#pragma unroll
      for (int k = 0; k < NP2; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          X[k][c] = pk2(x[2 * k][c], x[2 * k + 1][c]);
          F[k][c] = 0ull;
        }
        PT[k] = 0ull;
      }
      const ulonglong2* rec = reinterpret_cast<const ulonglong2*>(pos);
#pragma unroll 2
      for (int j = 0; j < B; ++j) {
        const ulonglong2 a = rec[2 * j], b = rec[2 * j + 1];  // (xx, yy), (zz, mm)
#pragma unroll
        for (int k = 0; k < NP2; ++k) {
          const f32x2 dx = fma2(X[k][0], m1, a.x), dy = fma2(X[k][1], m1, a.y), dz = fma2(X[k][2], m1, b.x);
          const f32x2 r2 = fma2(dx, dx, fma2(dy, dy, fma2(dz, dz, e2)));
          float i0 = rsqrt_ftz(pk_lo(r2)), i1 = rsqrt_ftz(pk_hi(r2));
          if (EPS0) {  // self pair / coincident bodies contribute nothing
            i0 = pk_lo(r2) > 0.f ? i0 : 0.f;
            i1 = pk_hi(r2) > 0.f ? i1 : 0.f;
          }
          const f32x2 inv = pk2(i0, i1);
          const f32x2 mi = mul2(b.y, inv);
          const f32x2 w3 = mul2(mul2(mi, inv), inv);
          F[k][0] = fma2(dx, w3, F[k][0]);
          F[k][1] = fma2(dy, w3, F[k][1]);
          F[k][2] = fma2(dz, w3, F[k][2]);
          if (wantE) PT[k] = add2(PT[k], mi);
        }
      }
#pragma unroll
      for (int k = 0; k < NP2; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          f[2 * k][c] = pk_lo(F[k][c]);
          f[2 * k + 1][c] = pk_hi(F[k][c]);
        }
        pot[2 * k] = pk_lo(PT[k]);
        pot[2 * k + 1] = pk_hi(PT[k]);
      }
    }
  };
  auto sweep = [&](bool wantE) {
    if constexpr (sizeof(T) == 4) {
      if (wantE)
        sweep_packed(std::true_type{});
      else
        sweep_packed(std::false_type{});
    } else {
      sweep_scalar(wantE);
    }
  };
  // U = -0.5 G sum_i m_i (pot_i - self term)
  auto energy = [&]() -> T {
    T e = T(0);
    const T self = EPS0 ? T(0) : Ar<T>::rsqrt_(eps2);
#pragma unroll
    for (int s = 0; s < NB_TI; ++s)
      if (own[s]) e += bm[s] * (pot[s] - bm[s] * self);
    return T(-0.5) * G * block_sum<T>(e, red, nwarps);
  };
  auto publish = [&]() {
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NB_TI; ++s)
      if (own[s]) BodyRec<T>::store(pos, body[s], x[s][0], x[s][1], x[s][2], bm[s]);
    __syncthreads();
  };

  __syncthreads();
  T oldH = T(0), newH = T(0);
  const bool caching = hmc && pa.fcache != nullptr;
  if (caching && pa.cache_read) {
    // the previous iteration left f and U of this very position behind: no sweep
#pragma unroll
    for (int s = 0; s < NB_TI; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c) f[s][c] = own[s] ? pa.fcache[((long long)c * B + body[s]) * A.P + part] : T(0);
    oldH = T(0.5) * block_sum<T>(ksum, red, nwarps) * inv_M + pa.ucache[part];
  } else {
    sweep(hmc);
    if (hmc) {
      const T U0 = energy();
      oldH = T(0.5) * block_sum<T>(ksum, red, nwarps) * inv_M + U0;
      if (caching) {  // seed the cache with the start state (kept if the proposal is rejected)
#pragma unroll
        for (int s = 0; s < NB_TI; ++s)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (own[s]) pa.fcache[((long long)c * B + body[s]) * A.P + part] = f[s][c];
        if (tid == 0) pa.ucache[part] = U0;
      }
    }
  }
  // acceleration of coordinate (s, c) = -dU/dq / M = G m_i f / M
  const T h = A.h, h2 = A.h2;
  const int L = A.L;
  if (integ == INTEG_LEAPFROG) {
    // kick-drift-kick form of src/integrator.py:112-118 (identical up to rounding; no stored a)
    if (L > 0) {
#pragma unroll
      for (int s = 0; s < NB_TI; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) st[s][c] += (T(0.5) * h * G * inv_M) * bm[s] * f[s][c];
    }
    for (int step = 0; step < L; ++step) {
#pragma unroll
      for (int s = 0; s < NB_TI; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) x[s][c] = fma(h, st[s][c], x[s][c]);
      publish();
      sweep(hmc && step == L - 1);
      const T kf = (step == L - 1 ? T(0.5) * h : h) * G * inv_M;
#pragma unroll
      for (int s = 0; s < NB_TI; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) st[s][c] += kf * bm[s] * f[s][c];
    }
  } else {
    // Stormer-Verlet, src/integrator.py:142-163; st: v -> qPast
#pragma unroll
    for (int s = 0; s < NB_TI; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const T xn = x[s][c] + st[s][c] * h + T(0.5) * (G * bm[s] * f[s][c] * inv_M) * h2;
        st[s][c] = x[s][c];
        x[s][c] = xn;
      }
    publish();
    for (int step = 0; step < L; ++step) {
      sweep(false);
#pragma unroll
      for (int s = 0; s < NB_TI; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const T xn = T(2) * x[s][c] - st[s][c] + (G * bm[s] * f[s][c] * inv_M) * h2;
          st[s][c] = x[s][c];
          x[s][c] = xn;
        }
      publish();
    }
#pragma unroll
    for (int s = 0; s < NB_TI; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c) st[s][c] = (x[s][c] - st[s][c]) / h;
    if (hmc) sweep(true);
  }
  // p = v * M
  ksum = T(0);
#pragma unroll
  for (int s = 0; s < NB_TI; ++s)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      st[s][c] *= M;
      ksum += st[s][c] * st[s][c];
    }

  bool rej = false;
  T accp = T(1);
  T U1 = T(0);
  if (hmc) {
    U1 = energy();
    newH = T(0.5) * block_sum<T>(ksum, red, nwarps) * inv_M + U1;
    T u;
    if (A.u != nullptr)
      u = A.u[part];
    else
      u = NormalBlock<T>::uniform(PhiloxKey(A.seed, A.iter), A.offset + (u64)part, A.D);
    rej = metropolis_reject<T>(oldH, newH, u, A.flags, &accp);  // identical in every thread
  }

  // ---- write back ------------------------------------------------------------------------
  if (!rej) {
#pragma unroll
    for (int s = 0; s < NB_TI; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (own[s]) A.q[((long long)c * B + body[s]) * A.q_ld + part] = x[s][c];
    if (caching && integ == INTEG_LEAPFROG && L > 0) {  // f, U of the accepted end state (the last sweep's)
#pragma unroll
      for (int s = 0; s < NB_TI; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (own[s]) pa.fcache[((long long)c * B + body[s]) * A.P + part] = f[s][c];
      if (tid == 0) pa.ucache[part] = U1;
    }
  }
  if (A.p != nullptr) {
    if (rej) {
      if (A.flags & FLAG_BUGCOMPAT) {
#pragma unroll
        for (int s = 0; s < NB_TI; ++s)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            st[s][c] = own[s] ? A.q[((long long)c * B + body[s]) * A.q_ld + part] : T(0);  // HMC.py:176 (sic)
      } else {
        draw(st);
      }
    }
#pragma unroll
    for (int s = 0; s < NB_TI; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (own[s]) A.p[((long long)c * B + body[s]) * A.p_ld + part] = st[s][c];
  }
  if (hmc && tid == 0) {
    if (A.accept != nullptr) A.accept[part] = rej ? 0 : 1;
    if (A.partials != nullptr) {  // [P][3]: per-particle scalars; coordinate sums come from k_colstats
      A.partials[part * 3 + 0] = rej ? 0.0 : 1.0;
      A.partials[part * 3 + 1] = (double)accp;
      A.partials[part * 3 + 2] = (double)(rej ? oldH : newH);
    }
  }
}

// U and grad U for the N-body family (one CTA per particle, same sweep).
template <typename T, bool EPS0>
__global__ void __launch_bounds__(1024) k_nbody_eval(const T* q, long long q_ld, T* energy, T* grad, long long g_ld,
                                                     const NBodyArgs<T> pa) {
  typedef typename V4<T>::type Vec;
  extern __shared__ __align__(32) unsigned char k3_smem_raw[];
  const int B = pa.B, NT = blockDim.x, tid = threadIdx.x, nwarps = NT >> 5;
  Vec* pos = reinterpret_cast<Vec*>(k3_smem_raw);
  T* red = reinterpret_cast<T*>(pos + B);
  const long long part = blockIdx.x;
  for (int b = tid; b < B; b += NT)
    pos[b] = V4<T>::make(q[(0LL * B + b) * q_ld + part], q[(1LL * B + b) * q_ld + part], q[(2LL * B + b) * q_ld + part],
                         pa.bmass[b]);
  __syncthreads();
  T e = T(0);
  const T self = EPS0 ? T(0) : Ar<T>::rsqrt_(pa.eps2);
  for (int b = tid; b < B; b += NT) {
    const Vec pi = pos[b];
    T fx = T(0), fy = T(0), fz = T(0), pt = T(0);
    for (int j = 0; j < B; ++j) {
      const Vec pj = pos[j];
      const T dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
      const T r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, pa.eps2)));
      T inv = Ar<T>::rsqrt_(r2);
      if (EPS0) inv = r2 > T(0) ? inv : T(0);
      const T mi = pj.w * inv, w3 = mi * inv * inv;
      fx = fma(dx, w3, fx);
      fy = fma(dy, w3, fy);
      fz = fma(dz, w3, fz);
      pt += mi;
    }
    if (grad) {  // dU/dr_i = -G m_i f
      grad[(0LL * B + b) * g_ld + part] = -pa.G * pi.w * fx;
      grad[(1LL * B + b) * g_ld + part] = -pa.G * pi.w * fy;
      grad[(2LL * B + b) * g_ld + part] = -pa.G * pi.w * fz;
    }
    e += pi.w * (pt - pi.w * self);
  }
  const T tot = block_sum<T>(e, red, nwarps);
  if (energy && tid == 0) energy[part] = T(-0.5) * pa.G * tot;
}

// sum_i q[d, i] and sum_i q[d, i]^2 per coordinate d (one CTA per d): the statistics vector for
// families whose trajectory kernel is not particle-per-thread.
template <typename T>
__global__ void __launch_bounds__(256) k_colstats(const T* __restrict__ q, long long q_ld, long long P, int D,
                                                  double* __restrict__ out /* [3 + 2D] */) {
  const int d = blockIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (long long i = threadIdx.x; i < P; i += blockDim.x) {
    const double v = (double)q[d * q_ld + i];
    s1 += v;
    s2 += v * v;
  }
  __shared__ double sm[2][8];
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    sm[0][threadIdx.x >> 5] = s1;
    sm[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += sm[0][w];
      b += sm[1][w];
    }
    out[3 + d] = a;
    out[3 + D + d] = b;
  }
}

}  // namespace ehmc
