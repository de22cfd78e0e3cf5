// K2-TC: dense-precision Gaussian trajectory kernel on the 5th-generation tensor cores.
//
// The gradient of a tile of 128 particles, G[128 x D] = X[128 x D] Lambda^T, is a GEMM that
// is re-issued L+1 times on on-chip state.  Here it runs as tcgen05.mma (kind::tf32,
// M = 128, N = 16*NCH, K = 8 per instruction) with the accumulator in TMEM; float32
// accuracy is recovered with the 3xTF32 split
//       x = x_hi + x_lo,  Lambda = L_hi + L_lo   (hi = the 19 bits the MMA reads, lo = exact remainder)
//       G ~= x_hi L_hi + x_lo L_hi + x_hi L_lo   (the dropped x_lo L_lo term is 2^-22 relative)
// which tracks the float64 oracle to ~1e-6 over a 50-step trajectory (float32 FMA: 4e-7).
//
// CTA = 4 epilogue warps (thread t <-> particle row t <-> TMEM lane t) + 1 MMA warp.
//   shared memory : A_hi, A_lo [KP/4][128][4]  (the particle tile, canonical K-major no-swizzle
//                   UMMA layout = one conflict-free 16-byte store per thread and 4 dims),
//                   B_hi, B_lo [KP/4][NP][4]   (Lambda, same layout)
//   tensor memory : D[128 lanes][NP columns] fp32 accumulator
//   registers     : the thread's D velocities for the whole trajectory
// Per evaluation: MMA warp issues 3*KP/8 tcgen05.mma + tcgen05.commit -> mbarrier; epilogue
// threads tcgen05.ld their row of G, kick the velocities, drift x = hi + lo, re-split and
// store the operand for the next evaluation.  HBM traffic: q in, q out, once per iteration.
//
// Reference arithmetic replaced: src/integrator.py:105-120 with gradient = Lambda (q - mu),
// src/HMC.py:106-116,168-176, src/ensemble.py:88-91 (kick-drift-kick form, see k_dense.cuh).
#pragma once

#include "common.cuh"
#include "k_dense.cuh"  // one_normal

namespace ehmc {

constexpr int TC_M = 128;            // particles per CTA
constexpr int TC_EPI_WARPS = 8;      // epilogue warps: warp w -> TMEM lane quarter w % 4, column half w / 4
constexpr int TC_THREADS = 32 * (TC_EPI_WARPS + 1);  // + 1 MMA warp

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// one elected lane of a converged warp (lets the compiler keep tcgen05 operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with A in tensor memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  uint32_t u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

// issue only (no wait): 8 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&u)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// round-to-nearest (ties away) to the 10-bit TF32 mantissa; two integer ops (cvt.rna.tf32.f32 expands
// to four because it also special-cases inf/nan, which a diverged trajectory does not need here)
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// what the tf32 MMA actually multiplies when handed a float32 bit pattern: the low 13 mantissa bits
// are ignored.  The 3xTF32 split used here is x_hi = trunc(x) (implicit: x itself is the operand),
// x_lo = x - trunc(x) (exact, stored explicitly).
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// K-major, no-swizzle shared-memory matrix descriptor of one K = 8 (two 16-byte chunks) slice of an
// operand stored as [K/4][R][4] floats: core matrices (8 rows x 16 B) contiguous along the rows
// (SBO = 128 B), the next K chunk R*16 B further (LBO).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t rows) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((rows * 16u >> 4) & 0x3FFF) << 16;  // leading-dimension byte offset (K direction)
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;        // stride-dimension byte offset (M/N direction)
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  return d;                                           // layout_type = 0 (no swizzle), base_offset = 0
}

__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4)                      // D format F32
         | (2u << 7) | (2u << 10)      // A, B format TF32
         | ((uint32_t)(N >> 3) << 17)  // N / 8
         | ((uint32_t)(M >> 4) << 24); // M / 16   (A, B K-major: bits 15, 16 = 0)
}

// K8 = KP / 8 MMA K-steps (KP = D rounded up to 8); MMA N = NP = KP rounded up to 16.
template <int K8>
struct TcShape {
  static constexpr int KP = 8 * K8;
  static constexpr int K4 = 2 * K8;                    // 16-byte chunks along K
  static constexpr int NP = (KP + 15) / 16 * 16;       // MMA N (multiple of 16 for M = 128)
  static constexpr int C0 = (K8 + 1) / 2;              // 8-column chunks owned by column half 0
  static constexpr int TMEM_COLS = NP <= 32 ? 32 : NP <= 64 ? 64 : 128;
  static constexpr size_t smem_bytes() {
    return (size_t)K4 * (TC_M + NP) * 16 * 2 + 8 * TC_M * sizeof(float) + 64;
  }
};

struct DenseTcArgs {
  const float* Bhi;  // [K4][NP][4]
  const float* Blo;
  const float* mu;   // [KP] zero padded
  int dbg;           // profiling knobs: 1 = issue no MMAs (commit only), 2 = skip the epilogue arithmetic
  long long* prof;   // optional: clock64() trace of CTA 0 / tile 0 / thread 0 (see ehmc_ctx_set_option "tc_prof")
};

// One evaluation's epilogue for this thread's 8*NC columns starting at 8-column chunk cb:
// read G from TMEM, (energy), kick, drift, re-split, store the next operand.
template <int K8, bool FIRST, bool LAST, bool KICK>
__device__ __forceinline__ void tc_epilogue(float (&v)[8 * TcShape<K8>::C0], float4* Ahi, float4* Alo, int row,
                                            uint32_t trow, int cb, int nc, float ck, float h, bool wantE,
                                            float* Uout) {
  constexpr int C0 = TcShape<K8>::C0;
  uint32_t g[C0][8];
#pragma unroll
  for (int c = 0; c < C0; ++c)
    if (c < nc) tmem_ld8_issue(trow + (uint32_t)(8 * (cb + c)), g[c]);
  tmem_ld_wait();
  float Uacc = 0.f;
#pragma unroll
  for (int c = 0; c < C0; ++c)
    if (c < nc) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float4 hi = Ahi[(2 * (cb + c) + t) * TC_M + row];  // x itself (see tf32_trunc)
        float x[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float gd = __uint_as_float(g[c][4 * t + e]);
          if (FIRST || LAST) Uacc = fmaf(x[e], gd, Uacc);
          if (KICK) v[8 * c + 4 * t + e] = fmaf(-ck, gd, v[8 * c + 4 * t + e]);
          if (!LAST) x[e] = fmaf(h, v[8 * c + 4 * t + e], x[e]);
        }
        if (!LAST) {
          float4 nl;
          nl.x = x[0] - tf32_trunc(x[0]); nl.y = x[1] - tf32_trunc(x[1]);
          nl.z = x[2] - tf32_trunc(x[2]); nl.w = x[3] - tf32_trunc(x[3]);
          Ahi[(2 * (cb + c) + t) * TC_M + row] = make_float4(x[0], x[1], x[2], x[3]);
          Alo[(2 * (cb + c) + t) * TC_M + row] = nl;
        }
      }
    }
  if ((FIRST || LAST) && wantE) *Uout = 0.5f * Uacc;
}

template <int K8>
__global__ void __launch_bounds__(TC_THREADS, 1) k_dense_tc(const IterArgs<float> A, const DenseTcArgs pa,
                                                            const int hmc) {
  typedef TcShape<K8> S;
  constexpr int NP = S::NP, K4 = S::K4, C0 = S::C0, VN = 8 * C0;
  extern __shared__ __align__(128) unsigned char tc_smem_raw[];
  const int D = A.D;
  float4* Ahi = reinterpret_cast<float4*>(tc_smem_raw);          // [K4][128]
  float4* Alo = Ahi + (size_t)K4 * TC_M;
  float4* Bhi = Alo + (size_t)K4 * TC_M;                         // [K4][NP]
  float4* Blo = Bhi + (size_t)K4 * NP;
  float* xch = reinterpret_cast<float*>(Blo + (size_t)K4 * NP);  // [2 halves][4][128] energy exchange
  uint64_t* mbar = reinterpret_cast<uint64_t*>(xch + 8 * TC_M);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_epi = warp < TC_EPI_WARPS;
  const int row = (warp & 3) * 32 + lane;        // particle row in the tile = TMEM lane
  const int half = (warp >> 2) & 1;              // column half
  const int cb = half ? C0 : 0;                  // first 8-column chunk of this thread
  const int nc = half ? K8 - C0 : C0;            // number of chunks
  const long long prow = (long long)blockIdx.x * TC_M + row;
  const bool valid = is_epi && prow < A.P;
  const long long pc = valid ? prow : 0;

  // ---- setup ---------------------------------------------------------------------------
  if (warp == TC_EPI_WARPS) {
    tmem_alloc(tmem_slot, S::TMEM_COLS);
    if (lane == 0) {
      mbar_init(mbar, 1);
      fence_barrier_init();
    }
  }
  {  // Lambda (host-packed hi / lo in the canonical layout): straight vector copy
    const float4* s0 = reinterpret_cast<const float4*>(pa.Bhi);
    const float4* s1 = reinterpret_cast<const float4*>(pa.Blo);
    for (int i = tid; i < K4 * NP; i += TC_THREADS) {
      Bhi[i] = s0[i];
      Blo[i] = s1[i];
    }
  }

  float v[VN];  // velocities of this thread's columns (dims >= D stay 0)
  float m = 1.f, inv_m = 1.f, pstd = 0.f, Kpart = 0.f;
  if (is_epi) {
    m = valid ? A.mass[pc] : 1.f;
    inv_m = 1.f / m;
    pstd = hmc ? momentum_std<float>(m, A.kB, A.temp, A.pscale) : 0.f;
    const PhiloxKey K(A.seed, A.iter);
#pragma unroll
    for (int c = 0; c < C0; ++c)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int k4 = 2 * (cb + c) + t;
        float x[4] = {0.f, 0.f, 0.f, 0.f}, p[4] = {0.f, 0.f, 0.f, 0.f}, zz[4];
        if (c < nc) {
          if (hmc && A.z == nullptr && k4 * 4 < D) NormalBlock<float>::draw(K, A.offset + (u64)pc, (uint32_t)k4, zz);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int d = 4 * k4 + e;
            if (d < D && valid) {
              x[e] = A.q[d * A.q_ld + pc] - pa.mu[d];
              if (!hmc)
                p[e] = A.p[d * A.p_ld + pc];
              else if (A.z != nullptr)
                p[e] = A.z[d * A.z_ld + pc] * pstd;
              else
                p[e] = zz[e] * pstd;
            }
            Kpart = fmaf(p[e], p[e], Kpart);
          }
          float4 lo;
          lo.x = x[0] - tf32_trunc(x[0]); lo.y = x[1] - tf32_trunc(x[1]);
          lo.z = x[2] - tf32_trunc(x[2]); lo.w = x[3] - tf32_trunc(x[3]);
          Ahi[k4 * TC_M + row] = make_float4(x[0], x[1], x[2], x[3]);
          Alo[k4 * TC_M + row] = lo;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) v[8 * c + 4 * t + e] = p[e] * inv_m;
      }
  }
  fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int L = A.L;
  const float h = A.h;
  float U0 = 0.f, U1 = 0.f;

  if (warp == TC_EPI_WARPS) {
    // ===== MMA issuer warp =====
    const uint32_t idesc = umma_idesc_tf32(TC_M, NP);
    const uint64_t da_hi = umma_desc(smem_u32(Ahi), TC_M), da_lo = umma_desc(smem_u32(Alo), TC_M);
    const uint64_t db_hi = umma_desc(smem_u32(Bhi), NP), db_lo = umma_desc(smem_u32(Blo), NP);
    constexpr uint64_t a_step = (2u * TC_M * 16u) >> 4, b_step = (2u * NP * 16u) >> 4;  // start-address field units
    for (int ev = 0; ev <= L; ++ev) {
      if (ev > 0) {
        asm volatile("bar.sync 1, %0;" ::"r"(TC_THREADS) : "memory");  // operands of this evaluation are in smem
        tc_fence_after();
      }
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < K8; ++j) umma_tf32_ss(tmem_d, da_hi + j * a_step, db_hi + j * b_step, idesc, j > 0);
#pragma unroll
        for (int j = 0; j < K8; ++j) umma_tf32_ss(tmem_d, da_lo + j * a_step, db_hi + j * b_step, idesc, 1);
#pragma unroll
        for (int j = 0; j < K8; ++j) umma_tf32_ss(tmem_d, da_hi + j * a_step, db_lo + j * b_step, idesc, 1);
        umma_commit(mbar);  // implies tcgen05.fence::before_thread_sync
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue warps =====
    const uint32_t trow = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    const float ckh = 0.5f * h * inv_m, ckf = h * inv_m;
    const bool wantE = hmc != 0;
    for (int ev = 0; ev <= L; ++ev) {
      mbar_wait(mbar, (uint32_t)(ev & 1));
      tc_fence_after();
      if (L == 0)
        tc_epilogue<K8, true, true, false>(v, Ahi, Alo, row, trow, cb, nc, 0.f, h, wantE, &U0);
      else if (ev == 0)
        tc_epilogue<K8, true, false, true>(v, Ahi, Alo, row, trow, cb, nc, ckh, h, wantE, &U0);
      else if (ev == L)
        tc_epilogue<K8, false, true, true>(v, Ahi, Alo, row, trow, cb, nc, ckh, h, wantE, &U1);
      else
        tc_epilogue<K8, false, false, true>(v, Ahi, Alo, row, trow, cb, nc, ckf, h, false, &U1);
      if (ev < L) {
        fence_proxy_async();
        tc_fence_before();  // our tcgen05.ld of D are complete (wait::ld) before the next MMA overwrites D
        asm volatile("bar.arrive 1, %0;" ::"r"(TC_THREADS) : "memory");
      }
    }
    if (L == 0) U1 = U0;
  }

  // ---- Hamiltonians: combine the two column halves through shared memory -----------------------
  float K1part = 0.f;
  if (is_epi) {
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      v[i] *= m;  // p = v * m
      K1part = fmaf(v[i], v[i], K1part);
    }
    if (hmc) {
      xch[(half * 4 + 0) * TC_M + row] = Kpart;
      xch[(half * 4 + 1) * TC_M + row] = U0;
      xch[(half * 4 + 2) * TC_M + row] = K1part;
      xch[(half * 4 + 3) * TC_M + row] = U1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS) tmem_dealloc(tmem_d, S::TMEM_COLS);
  if (!is_epi) return;

  bool rej = false;
  float accp = 1.f, oldH = 0.f, newH = 0.f;
  if (hmc) {
    const float K0 = xch[0 * TC_M + row] + xch[4 * TC_M + row], Ua = xch[1 * TC_M + row] + xch[5 * TC_M + row];
    const float K1 = xch[2 * TC_M + row] + xch[6 * TC_M + row], Ub = xch[3 * TC_M + row] + xch[7 * TC_M + row];
    oldH = 0.5f * K0 * inv_m + Ua;
    newH = 0.5f * K1 * inv_m + Ub;
    float u = 0.f;
    if (valid) u = A.u != nullptr ? A.u[pc] : NormalBlock<float>::uniform(PhiloxKey(A.seed, A.iter), A.offset + (u64)pc);
    rej = metropolis_reject<float>(oldH, newH, u, A.flags, &accp);  // identical in both halves
  }
  const bool need_old = rej && (A.partials != nullptr || (A.p != nullptr && (A.flags & FLAG_BUGCOMPAT)));
  double* prow_out = A.partials ? A.partials + ((size_t)blockIdx.x * TC_EPI_WARPS + warp) * (2 * D + 3) : nullptr;
#pragma unroll
  for (int c = 0; c < C0; ++c) {
    if (c >= nc) break;  // warp-uniform
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int k4 = 2 * (cb + c) + t;
      const float4 hi = Ahi[k4 * TC_M + row];
      const float x[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 4 * k4 + e;
        if (d >= D) continue;  // warp-uniform
        const float qn = x[e] + pa.mu[d];
        float qold = 0.f;
        if (valid && need_old) qold = A.q[d * A.q_ld + pc];
        if (valid && !rej) A.q[d * A.q_ld + pc] = qn;  // HMC.py:175: rejected rows keep the value in HBM
        if (A.p != nullptr && valid) {
          float pv = v[8 * c + 4 * t + e];
          if (rej) {
            if (A.flags & FLAG_BUGCOMPAT)
              pv = qold;  // HMC.py:176 (sic)
            else if (A.z != nullptr)
              pv = A.z[d * A.z_ld + pc] * pstd;
            else
              pv = one_normal<float>(A.seed, A.iter, A.offset + (u64)pc, d) * pstd;
          }
          A.p[d * A.p_ld + pc] = pv;
        }
        if (prow_out != nullptr) {
          const double qk = valid ? (double)(rej ? qold : qn) : 0.0;
          const double s1 = warp_sum(qk), s2 = warp_sum(qk * qk);
          if (lane == 0) {
            prow_out[3 + d] = s1;
            prow_out[3 + D + d] = s2;
          }
        }
      }
    }
  }
  if (half == 0) {
    if (hmc && valid && A.accept != nullptr) A.accept[pc] = rej ? 0 : 1;
  }
  if (prow_out != nullptr) {
    // rows of different warps cover different (quarter, half) pairs; per-particle scalars only from half 0
    double s_acc = 0.0, s_accp = 0.0, s_h = 0.0;
    if (valid && half == 0) {
      s_acc = rej ? 0.0 : 1.0;
      s_accp = (double)accp;
      s_h = (double)(rej ? oldH : newH);
    }
    s_acc = warp_sum(s_acc);
    s_accp = warp_sum(s_accp);
    s_h = warp_sum(s_h);
    if (lane == 0) {
      prow_out[0] = s_acc;
      prow_out[1] = s_accp;
      prow_out[2] = s_h;
      // dims not owned by this warp's column half contribute nothing: zero them
      for (int d = 0; d < D; ++d) {
        const int ch = d / 8;
        const bool mine = ch >= cb && ch < cb + nc;
        if (!mine) {
          prow_out[3 + d] = 0.0;
          prow_out[3 + D + d] = 0.0;
        }
      }
    }
  }
}

}  // namespace ehmc
