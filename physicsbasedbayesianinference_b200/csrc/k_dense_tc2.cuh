// K2-TC2: dense-precision Gaussian trajectory kernel, tensor cores, two particle tiles per CTA.
//
// Same arithmetic as k_dense_tc.cuh (3xTF32 split GEMM, kick-drift-kick), restructured so the
// tensor pipe and the CUDA-core epilogue overlap:
//   * a CTA owns TWO tiles of 128 particles; each tile is served by its own group of 4 warps
//     (thread <-> particle row <-> TMEM lane, all D columns of the row in one thread);
//   * the tile's x operand (full fp32; the tf32 MMA uses its top 19 bits = x_hi) lives in TENSOR MEMORY (tcgen05.mma with A from TMEM, written by
//     the row's thread with tcgen05.st), x_lo in shared memory, Lambda_hi / Lambda_lo in shared
//     memory shared by both tiles, the accumulator D in TMEM:
//         TMEM columns  [0,128) D0 | [128,256) A_hi0 | [256,384) D1 | [384,512) A_hi1
//   * the two groups run the same loop independently
//         bar.sync(group) -> lane 0 issues 3*K/8 MMAs + commit -> mbarrier wait -> epilogue
//     so while one tile's MMAs execute, the other tile's threads kick/drift/re-split.
#pragma once

#include "k_dense_tc.cuh"

namespace ehmc {

constexpr int TC2_THREADS = 256;  // 2 groups x 4 warps

template <int N>
struct TmemRegs {
  uint32_t r[N];
};

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&u)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_issue2(uint32_t taddr, uint32_t (&u)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
}
// tcgen05.wait::ld with the loaded registers as read-write operands: every later use of them
// depends on this statement, so the compiler cannot hoist arithmetic above the wait.
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&u)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
                 "+r"(u[8]), "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld8(uint32_t (&u)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&u)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]),
      "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&u)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(u[0]),
               "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int K8>
struct Tc2Shape {
  static constexpr int KP = 8 * K8, K4 = 2 * K8;
  static constexpr int NP = (KP + 15) / 16 * 16;
  static constexpr size_t smem_bytes() { return (size_t)K4 * (2 * TC_M + 2 * NP) * 16 + (size_t)KP * 4 + 64; }
};

// ---- software-pipelined epilogue ---------------------------------------------------------
// Batches of 16 columns (a trailing batch of 8 when K8 is odd).  tcgen05.wait::ld waits for ALL
// outstanding TMEM loads, so the pipeline is: wait(batch b) -> issue loads(batch b+1) ->
// compute(batch b) -> store(batch b): the TMEM / shared-memory loads of the next batch are in
// flight while the current batch is kicked, drifted and re-split.
template <int NCOL>
__device__ __forceinline__ void tc2_issue(uint32_t (&g)[16], uint32_t (&hh)[16], uint32_t t_d, uint32_t t_a,
                                          int col0) {
  if constexpr (NCOL == 16) {
    tmem_ld16_issue(t_d + (uint32_t)col0, g);
    tmem_ld16_issue(t_a + (uint32_t)col0, hh);
  } else {
    uint32_t(&g8)[8] = reinterpret_cast<uint32_t(&)[8]>(g);
    uint32_t(&h8)[8] = reinterpret_cast<uint32_t(&)[8]>(hh);
    tmem_ld8_issue2(t_d + (uint32_t)col0, g8);
    tmem_ld8_issue2(t_a + (uint32_t)col0, h8);
  }
}

template <int NCOL>
__device__ __forceinline__ void tc2_wait(uint32_t (&g)[16], uint32_t (&hh)[16]) {
  if constexpr (NCOL == 16) {
    tmem_wait_ld16(g);
    tmem_wait_ld16(hh);
  } else {
    tmem_wait_ld8(reinterpret_cast<uint32_t(&)[8]>(g));
    tmem_wait_ld8(reinterpret_cast<uint32_t(&)[8]>(hh));
  }
}

// hh holds the row's x (full fp32) as stored in TMEM: the tf32 MMA ignores the 13 low mantissa
// bits, so x itself is the "hi" operand and only lo = x - trunc_tf32(x) needs explicit storage.
// One code path for every evaluation (first / middle / last differ only in ck, h and `store`):
// keeps the unrolled epilogue at one copy, which matters for the instruction cache.
template <int NCOL>
__device__ __forceinline__ void tc2_compute(float* v, uint32_t (&g)[16], uint32_t (&hh)[16], float4* Alo, int row,
                                            uint32_t t_a, int col0, float ck, float h, bool store, float& Uacc) {
#pragma unroll
  for (int q = 0; q < NCOL / 4; ++q) {
    float nl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = 4 * q + e;
      float x = __uint_as_float(hh[i]);
      const float gd = __uint_as_float(g[i]);
      Uacc = fmaf(x, gd, Uacc);
      v[i] = fmaf(-ck, gd, v[i]);
      x = fmaf(h, v[i], x);
      hh[i] = __float_as_uint(x);
      nl[e] = x - tf32_trunc(x);
    }
    if (store) Alo[(col0 / 4 + q) * TC_M + row] = make_float4(nl[0], nl[1], nl[2], nl[3]);
  }
  if (store) {
    if constexpr (NCOL == 16)
      tmem_st16(t_a + (uint32_t)col0, hh);
    else
      tmem_st8(t_a + (uint32_t)col0, reinterpret_cast<uint32_t(&)[8]>(hh));
  }
}

template <int K8, int B>
__device__ __forceinline__ void tc2_pipe(float (&v)[8 * K8], uint32_t (&g)[2][16], uint32_t (&hh)[2][16],
                                         float4* Alo, int row, uint32_t t_d, uint32_t t_a, float ck, float h,
                                         bool store, float& Uacc) {
  constexpr int NB = (K8 + 1) / 2;
  constexpr int NCOL = (B == NB - 1 && (K8 % 2) == 1) ? 8 : 16;
  constexpr int cur = B & 1;
  tc2_wait<NCOL>(g[cur], hh[cur]);
  if constexpr (B + 1 < NB) {
    constexpr int NCOL_N = (B + 1 == NB - 1 && (K8 % 2) == 1) ? 8 : 16;
    tc2_issue<NCOL_N>(g[cur ^ 1], hh[cur ^ 1], t_d, t_a, 16 * (B + 1));
  }
  tc2_compute<NCOL>(&v[16 * B], g[cur], hh[cur], Alo, row, t_a, 16 * B, ck, h, store, Uacc);
  if constexpr (B + 1 < NB) tc2_pipe<K8, B + 1>(v, g, hh, Alo, row, t_d, t_a, ck, h, store, Uacc);
}

template <int K8>
__device__ __forceinline__ float tc2_epilogue(float (&v)[8 * K8], float4* Alo, int row, uint32_t t_d, uint32_t t_a,
                                              float ck, float h, bool store) {
  float Uacc = 0.f;
  uint32_t g[2][16], hh[2][16];
  constexpr int NCOL0 = (K8 == 1) ? 8 : 16;
  tc2_issue<NCOL0>(g[0], hh[0], t_d, t_a, 0);
  tc2_pipe<K8, 0>(v, g, hh, Alo, row, t_d, t_a, ck, h, store, Uacc);
  if (store) tmem_wait_st();
  return 0.5f * Uacc;
}

template <int K8>
__global__ void __launch_bounds__(TC2_THREADS, 1) k_dense_tc2(const IterArgs<float> A, const DenseTcArgs pa,
                                                              const int hmc) {
  typedef Tc2Shape<K8> S;
  constexpr int NP = S::NP, K4 = S::K4, KP = S::KP;
  extern __shared__ __align__(128) unsigned char tc2_smem_raw[];
  const int D = A.D;
  float4* Alo_all = reinterpret_cast<float4*>(tc2_smem_raw);      // [2][K4][128]
  float4* Bhi = Alo_all + (size_t)2 * K4 * TC_M;                  // [K4][NP]
  float4* Blo = Bhi + (size_t)K4 * NP;
  float* mus = reinterpret_cast<float*>(Blo + (size_t)K4 * NP);           // [KP] mean
  uint64_t* mbar = reinterpret_cast<uint64_t*>(mus + KP);                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
  const int tile = warp >> 2, quarter = warp & 3;
  const int row = quarter * 32 + lane;
  float4* Alo = Alo_all + (size_t)tile * K4 * TC_M;
  const long long prow = (long long)blockIdx.x * (2 * TC_M) + tile * TC_M + row;
  const bool valid = prow < A.P;
  const long long pc = valid ? prow : 0;

  if (pa.prof != nullptr && blockIdx.x == 0 && tid == 0) pa.prof[63] = clock64();
  // ---- setup ---------------------------------------------------------------------------
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    if (lane == 0) {
      mbar_init(&mbar[0], 1);
      mbar_init(&mbar[1], 1);
      fence_barrier_init();
    }
  }
  {
    const float4* s0 = reinterpret_cast<const float4*>(pa.Bhi);
    const float4* s1 = reinterpret_cast<const float4*>(pa.Blo);
    for (int i = tid; i < K4 * NP; i += TC2_THREADS) {
      Bhi[i] = s0[i];
      Blo[i] = s1[i];
    }
    if (tid < KP) mus[tid] = pa.mu[tid];
  }
  const bool prof = pa.prof != nullptr && blockIdx.x == 0 && tid == 0;
  int pi = 0;
  auto stamp = [&]() {
    if (prof && pi < 62) pa.prof[pi++] = clock64();
  };
  stamp();  // after TMEM alloc issue + Lambda copy issue
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  stamp();  // after the first CTA barrier
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
  const uint32_t t_d = tmem_base + (uint32_t)(tile * 256) + lane_off;        // this row's accumulator
  const uint32_t t_a = tmem_base + (uint32_t)(tile * 256 + 128) + lane_off;  // this row's x_hi operand

  float v[KP];
  const float m = valid ? A.mass[pc] : 1.f;
  const float inv_m = 1.f / m;
  const float pstd = hmc ? momentum_std<float>(m, A.kB, A.temp, A.pscale) : 0.f;
  float K0 = 0.f;
  {
    // positions: all global loads in flight before the first use (v doubles as the staging array)
#pragma unroll
    for (int d = 0; d < KP; ++d) v[d] = (d < D && valid) ? A.q[d * A.q_ld + pc] : 0.f;
#pragma unroll
    for (int c = 0; c < K8; ++c) {
      uint32_t hh[8];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float lo4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int d = 8 * c + 4 * t + e;
          const float x = (d < D && valid) ? v[d] - mus[d] : 0.f;
          hh[4 * t + e] = __float_as_uint(x);
          lo4[e] = x - tf32_trunc(x);
        }
        Alo[(2 * c + t) * TC_M + row] = make_float4(lo4[0], lo4[1], lo4[2], lo4[3]);
      }
      tmem_st8(t_a + (uint32_t)(8 * c), hh);
    }
    stamp();  // positions loaded, x operands stored
    // momenta
    if (!hmc) {
#pragma unroll
      for (int d = 0; d < KP; ++d) v[d] = (d < D && valid) ? A.p[d * A.p_ld + pc] : 0.f;
    } else if (A.z != nullptr) {
#pragma unroll
      for (int d = 0; d < KP; ++d) v[d] = (d < D && valid) ? A.z[d * A.z_ld + pc] * pstd : 0.f;
    } else {
      // 4 Philox blocks (16 normals) at a time so their serial rounds interleave; K4 = 2*K8 blocks
      const PhiloxKey K(A.seed, A.iter);
      constexpr int NW = 4;
#pragma unroll
      for (int k4 = 0; k4 < K4; k4 += NW) {
        float zz[4 * NW];
        if (k4 + NW <= K4) {
          NormalBlock<float>::draw_multi<NW>(K, A.offset + (u64)pc, (uint32_t)k4, zz);
        } else {
          NormalBlock<float>::draw_multi<2>(K, A.offset + (u64)pc, (uint32_t)k4, zz);  // K4 is even
        }
#pragma unroll
        for (int e = 0; e < 4 * NW; ++e)
          if (4 * k4 + e < KP) v[4 * k4 + e] = (4 * k4 + e < D && valid) ? zz[e] * pstd : 0.f;
      }
    }
#pragma unroll
    for (int d = 0; d < KP; ++d) {
      K0 = fmaf(v[d], v[d], K0);
      v[d] *= inv_m;
    }
    tmem_wait_st();
    K0 *= 0.5f * inv_m;
    stamp();  // momenta drawn
  }

  const int L = A.L;
  const float h = A.h;
  const uint32_t idesc = umma_idesc_tf32(TC_M, NP);
  const uint64_t da_lo = umma_desc(smem_u32(Alo), TC_M);
  const uint64_t db_hi = umma_desc(smem_u32(Bhi), NP), db_lo = umma_desc(smem_u32(Blo), NP);
  constexpr uint64_t a_step = (2u * TC_M * 16u) >> 4, b_step = (2u * NP * 16u) >> 4;
  const uint32_t mma_d = tmem_base + (uint32_t)(tile * 256), mma_a = mma_d + 128u;
  const float ckh = 0.5f * h * inv_m, ckf = h * inv_m;
  float U0 = 0.f, U1 = 0.f;

  stamp();
  for (int ev = 0; ev <= L; ++ev) {
    // the tile's operands (x_hi in TMEM, x_lo in smem) are complete once all 128 rows arrive
    fence_proxy_async();
    if (ev < 6) stamp();
    tc_fence_before();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + tile) : "memory");
    if (ev < 6) stamp();
    if (quarter == 0 && elect_one()) {
      tc_fence_after();
      if (!(pa.dbg & 1) && !((pa.dbg & 4) && tile == 0)) {
#pragma unroll
        for (int j = 0; j < K8; ++j) umma_tf32_ts(mma_d, mma_a + 8u * j, db_hi + j * b_step, idesc, j > 0);
#pragma unroll
        for (int j = 0; j < K8; ++j) umma_tf32_ss(mma_d, da_lo + j * a_step, db_hi + j * b_step, idesc, 1);
#pragma unroll
        for (int j = 0; j < K8; ++j) umma_tf32_ts(mma_d, mma_a + 8u * j, db_lo + j * b_step, idesc, 1);
      }
      umma_commit(&mbar[tile]);
    }
    if (ev < 6) stamp();
    mbar_wait(&mbar[tile], (uint32_t)(ev & 1));
    tc_fence_after();
    if (ev < 6) stamp();
    const bool first = ev == 0, last = ev == L;
    const float ck = L == 0 ? 0.f : ((first || last) ? ckh : ckf);
    const float Uev = ((pa.dbg & 2) || ((pa.dbg & 4) && tile == 1) || ((pa.dbg & 8) && tile == 0)) ? 0.f : tc2_epilogue<K8>(v, Alo, row, t_d, t_a, ck, last ? 0.f : h, !last);
    if (first) U0 = Uev;
    if (last) U1 = Uev;
    if (ev < 6) stamp();
  }
  if (L == 0) U1 = U0;
  stamp();

  // ---- Metropolis + write back ---------------------------------------------------------------
  float K1 = 0.f;
#pragma unroll
  for (int d = 0; d < KP; ++d) {
    v[d] *= m;  // p = v * m
    K1 = fmaf(v[d], v[d], K1);
  }
  bool rej = false;
  float accp = 1.f, oldH = 0.f, newH = 0.f;
  if (hmc) {
    oldH = K0 + U0;
    newH = 0.5f * K1 * inv_m + U1;
    float u = 0.f;
    if (valid) u = A.u != nullptr ? A.u[pc] : NormalBlock<float>::uniform(PhiloxKey(A.seed, A.iter), A.offset + (u64)pc);
    rej = metropolis_reject<float>(oldH, newH, u, A.flags, &accp);
  }
  stamp();
  if (A.p == nullptr && A.partials == nullptr) {
    // production fast path: only the accepted positions go back to HBM (HMC.py:175)
    // tcgen05.ld is warp-collective (.sync.aligned): every lane loads, only accepted rows store
    const bool wr = valid && !rej;
#pragma unroll
    for (int b = 0; b < K8 / 2; ++b) {
      uint32_t hh[16];
      tmem_ld16_issue(t_a + (uint32_t)(16 * b), hh);
      tmem_wait_ld16(hh);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int d = 16 * b + e;
        if (d < D && wr) A.q[d * A.q_ld + pc] = __uint_as_float(hh[e]) + mus[d];
      }
    }
    if constexpr (K8 % 2 == 1) {
      uint32_t hh[8];
      tmem_ld8_issue2(t_a + (uint32_t)(8 * (K8 - 1)), hh);
      tmem_wait_ld8(hh);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int d = 8 * (K8 - 1) + e;
        if (d < D && wr) A.q[d * A.q_ld + pc] = __uint_as_float(hh[e]) + mus[d];
      }
    }
    if (hmc && valid && A.accept != nullptr) A.accept[pc] = rej ? 0 : 1;
    stamp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    return;
  }
  const bool need_old = rej && (A.partials != nullptr || (A.p != nullptr && (A.flags & FLAG_BUGCOMPAT)));
  double* prow_out = A.partials ? A.partials + ((size_t)blockIdx.x * 8 + warp) * (2 * D + 3) : nullptr;
#pragma unroll
  for (int c = 0; c < K8; ++c) {
    if (8 * c >= D) break;  // uniform
    uint32_t hh[8];
    tmem_ld8_issue2(t_a + (uint32_t)(8 * c), hh);
    tmem_wait_ld8(hh);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int d = 8 * c + e;
      if (d >= D) continue;  // uniform
      const float qn = __uint_as_float(hh[e]) + pa.mu[d];
      float qold = 0.f;
      if (valid && need_old) qold = A.q[d * A.q_ld + pc];
      if (valid && !rej) A.q[d * A.q_ld + pc] = qn;  // HMC.py:175: rejected rows keep the value in HBM
      if (A.p != nullptr && valid) {
        float pv = v[d];
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
  if (hmc && valid && A.accept != nullptr) A.accept[pc] = rej ? 0 : 1;
  if (prow_out != nullptr) {
    double s_acc = valid ? (rej ? 0.0 : 1.0) : 0.0, s_accp = valid ? (double)accp : 0.0;
    double s_h = valid ? (double)(rej ? oldH : newH) : 0.0;
    s_acc = warp_sum(s_acc);
    s_accp = warp_sum(s_accp);
    s_h = warp_sum(s_h);
    if (lane == 0) {
      prow_out[0] = s_acc;
      prow_out[1] = s_accp;
      prow_out[2] = s_h;
    }
  }
  stamp();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace ehmc
