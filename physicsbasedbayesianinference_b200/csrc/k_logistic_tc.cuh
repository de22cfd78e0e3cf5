// K4-TC: Bayesian logistic-regression gradient (and energy) on the tensor cores -- BASELINE
// config 3 (X 100k x 256, 65 536 particles).  Flash-attention-shaped GEMM chain; the N x P
// logits never leave the SM:
//
//   per CTA: 128 particles (Theta tile, bf16, shared memory, resident)
//   per chunk of 128 data rows (X chunk bf16 + y, ONE TMA bulk copy into a 2-stage ring):
//     GEMM1  S[128 x 128]  = Theta_tile . X_chunk^T        tcgen05.mma kind::f16, K = D, D in TMEM
//     epilogue             r = sigmoid(S) - y  (tanh.approx: one MUFU), energy terms, bf16 -> smem
//     GEMM2  G[128 x D]   += R . X_chunk                   A = R (smem), B = the SAME smem chunk
//                                                          read MN-major, accumulator in TMEM
//   TMEM columns: S0 [0,128) | S1 [128,256) | G [256, 256+D)     (S double-buffered)
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue
// (thread <-> particle row and one half of the 128 logit columns).
// Shared-memory operand layout everywhere: canonical K-major no-swizzle UMMA layout
// [K/8][128 rows][8 bf16]; an X chunk stored that way is simultaneously the K-major B operand
// of GEMM1 (rows = data rows, K = d) and the MN-major B operand of GEMM2 (N = d, K = data rows).
//
// Inputs are rounded to bf16 (fp32 accumulation): this is the throughput path; the exact
// fp32 / fp64 path is k_logistic.cuh.  A trajectory driven by this gradient is still reversible
// and volume preserving (the gradient is a deterministic function of q).
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"
#include "k_dense_tc.cuh"
#include "k_dense_tc2.cuh"

namespace ehmc {

constexpr int LT_M = 128;    // particles per CTA
constexpr int LT_NB = 128;   // data rows per chunk
constexpr int LT_EPI_WARPS = 8;
constexpr int LT_THREADS = 32 * (2 + LT_EPI_WARPS);

struct LogisticTcArgs {
  const unsigned char* chunks;  // [NC] blocks of chunk_bytes: X part [DP/8][128][8] bf16, then y[128] float
  int NC;                       // number of chunks
  int DP;                       // D rounded up to 16
  int D;
  int n_pad;                    // zero rows appended to the last chunk (y = 0.5 there)
  unsigned chunk_bytes;
  float inv_s2;
};

__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int b_mn_major) {
  return (1u << 4)                           // D format F32
         | (1u << 7) | (1u << 10)           // A, B format BF16
         | ((uint32_t)b_mn_major << 16)     // B major: 0 = K, 1 = MN
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// generic no-swizzle descriptor: lbo / sbo in bytes
__device__ __forceinline__ uint64_t umma_desc2(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// grad[D,P] (and energy[P] when WITH_E) at theta[D,P]
template <bool WITH_E>
__global__ void __launch_bounds__(LT_THREADS, 1) k_logistic_tc(const float* __restrict__ theta, long long t_ld,
                                                               long long P, float* __restrict__ grad, long long g_ld,
                                                               float* __restrict__ energy, const LogisticTcArgs pa) {
  extern __shared__ __align__(128) unsigned char lt_smem[];
  const int DP = pa.DP, D = pa.D, NC = pa.NC;
  const uint32_t a_bytes = (uint32_t)DP * LT_M * 2;             // Theta tile, [DP/8][128][16 B]
  unsigned char* As = lt_smem;
  unsigned char* Xs0 = As + a_bytes;                            // 2 stages of chunk_bytes
  unsigned char* Rs = Xs0 + 2 * (size_t)pa.chunk_bytes;         // [NB/8][128][16 B]
  float* xch = reinterpret_cast<float*>(Rs + LT_NB * LT_M * 2);  // [2][128] energy exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * LT_M);
  uint64_t* x_full = bars;        // [2]
  uint64_t* x_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;    // [2]
  uint64_t* s_empty = bars + 6;   // [2]
  uint64_t* r_full = bars + 8;
  uint64_t* r_empty = bars + 9;
  uint64_t* g_done = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const long long p0 = (long long)blockIdx.x * LT_M;

  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&x_full[i], 1);
        mbar_init(&x_empty[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&s_empty[i], LT_EPI_WARPS);
      }
      mbar_init(r_full, LT_EPI_WARPS);
      mbar_init(r_empty, 1);
      mbar_init(g_done, 1);
      fence_barrier_init();
    }
  }
  // Theta tile -> bf16 canonical layout: thread (row, dk) builds 16-byte units
  for (int i = tid; i < (DP / 8) * LT_M; i += LT_THREADS) {
    const int dk = i / LT_M, row = i % LT_M;
    const long long pi = p0 + row;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int d = 8 * dk + e;
      v[e] = (d < D && pi < P) ? theta[d * t_ld + pi] : 0.f;
    }
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]);
    u.w = pack_bf16x2(v[6], v[7]);
    reinterpret_cast<uint4*>(As)[i] = u;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t t_g = tmem_base + 256u;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int c = 0; c < NC; ++c) {
        const int s = c & 1;
        mbar_wait(&x_empty[s], (uint32_t)(((c >> 1) & 1) ^ 1));
        mbar_expect_tx(&x_full[s], pa.chunk_bytes);
        tma_bulk_g2s(Xs0 + (size_t)s * pa.chunk_bytes, pa.chunks + (size_t)c * pa.chunk_bytes, pa.chunk_bytes,
                     &x_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc1 = umma_idesc_bf16(LT_M, LT_NB, 0);
      const uint32_t idesc2 = umma_idesc_bf16(LT_M, DP, 1);
      const uint64_t dA = umma_desc2(smem_u32(As), LT_M * 16, 128);       // K-major: LBO = next K chunk, SBO = 8 rows
      const uint64_t dR = umma_desc2(smem_u32(Rs), LT_M * 16, 128);
      const uint64_t k_step = (2u * LT_M * 16u) >> 4;                     // two 16-byte K chunks per MMA (K = 16)
      auto gemm2 = [&](int cc) {
        const int s = cc & 1;
        mbar_wait(r_full, (uint32_t)(cc & 1));
        tc_fence_after();
        // B = X chunk read MN-major: N = d (unit stride LT_NB*16 B between 8-d groups = SBO),
        // K = data rows (16 B apart, groups of 8 rows 128 B apart = LBO)
        const uint64_t dB2 = umma_desc2(smem_u32(Xs0 + (size_t)s * pa.chunk_bytes), 128, LT_NB * 16);
        for (int j = 0; j < LT_NB / 16; ++j)
          umma_bf16_ss(t_g, dR + j * k_step, dB2 + j * ((16u * 16u) >> 4), idesc2, (cc > 0 || j > 0) ? 1u : 0u);
        umma_commit(&x_empty[s]);  // the X stage and R are free once these MMAs have executed
        umma_commit(r_empty);
      };
      for (int c = 0; c < NC; ++c) {
        const int s = c & 1;
        mbar_wait(&x_full[s], (uint32_t)((c >> 1) & 1));
        mbar_wait(&s_empty[s], (uint32_t)(((c >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint64_t dB1 = umma_desc2(smem_u32(Xs0 + (size_t)s * pa.chunk_bytes), LT_NB * 16, 128);
        for (int j = 0; j < DP / 16; ++j)
          umma_bf16_ss(tmem_base + (uint32_t)(s * 128), dA + j * k_step, dB1 + j * k_step, idesc1, j > 0 ? 1u : 0u);
        umma_commit(&s_full[s]);
        if (c >= 1) gemm2(c - 1);
      }
      gemm2(NC - 1);
      umma_commit(g_done);
    }
  } else {
    // ===== epilogue warps: sigmoid-residual between the two GEMMs =====
    // a warp can only touch the TMEM lanes of ITS hardware quarter (warp index % 4)
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float Uacc = 0.f;
    for (int c = 0; c < NC; ++c) {
      const int s = c & 1;
      mbar_wait(&s_full[s], (uint32_t)((c >> 1) & 1));
      tc_fence_after();
      uint32_t sv[4][16];
#pragma unroll
      for (int b = 0; b < 4; ++b) tmem_ld16_issue(tmem_base + lane_off + (uint32_t)(s * 128 + half * 64 + 16 * b), sv[b]);
#pragma unroll
      for (int b = 0; b < 4; ++b) tmem_wait_ld16(sv[b]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);  // S buffer may be overwritten by GEMM1 of chunk c + 2
      const float* yv = reinterpret_cast<const float*>(Xs0 + (size_t)s * pa.chunk_bytes + (size_t)DP * LT_NB * 2);
      uint32_t rp[32];
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float r2[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const float sgm = __uint_as_float(sv[b][i + t]);
            const float yn = yv[half * 64 + 16 * b + i + t];
            const float sig = fmaf(0.5f, tanh_approx(0.5f * sgm), 0.5f);
            r2[t] = sig - yn;
            if (WITH_E) Uacc += fmaxf(sgm, 0.f) + __logf(1.f + __expf(-fabsf(sgm))) - yn * sgm;
          }
          rp[(16 * b + i) / 2] = pack_bf16x2(r2[0], r2[1]);
        }
      mbar_wait(r_empty, (uint32_t)((c & 1) ^ 1));  // GEMM2 of the previous chunk has consumed R
#pragma unroll
      for (int u8 = 0; u8 < 8; ++u8)
        reinterpret_cast<uint4*>(Rs)[(half * 8 + u8) * LT_M + row] =
            make_uint4(rp[4 * u8], rp[4 * u8 + 1], rp[4 * u8 + 2], rp[4 * u8 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(r_full);
    }
    // ---- final: G from TMEM, prior term, stores --------------------------------------------------
    mbar_wait(g_done, 0u);
    tc_fence_after();
    const long long pi = p0 + row;
    const int dh = DP / 2;  // columns of this half
    float t2 = 0.f;
    for (int b = 0; b < dh / 16; ++b) {
      uint32_t gv[16];
      tmem_ld16_issue(t_g + lane_off + (uint32_t)(half * dh + 16 * b), gv);
      tmem_wait_ld16(gv);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int d = half * dh + 16 * b + i;
        if (d < D && pi < P) {
          const float th = theta[d * t_ld + pi];
          t2 = fmaf(th, th, t2);
          if (grad) grad[d * g_ld + pi] = __uint_as_float(gv[i]) + th * pa.inv_s2;
        }
      }
    }
    if (dh % 16) {  // DP = 16 * odd: one 8-column tail per half
      uint32_t gv[8];
      tmem_ld8_issue2(t_g + lane_off + (uint32_t)(half * dh + (dh / 16) * 16), gv);
      tmem_wait_ld8(gv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int d = half * dh + (dh / 16) * 16 + i;
        if (d < D && pi < P) {
          const float th = theta[d * t_ld + pi];
          t2 = fmaf(th, th, t2);
          if (grad) grad[d * g_ld + pi] = __uint_as_float(gv[i]) + th * pa.inv_s2;
        }
      }
    }
    if (WITH_E) xch[half * LT_M + row] = Uacc + 0.5f * t2 * pa.inv_s2;
  }
  tc_fence_before();
  __syncthreads();
  if (WITH_E && tid < LT_M && p0 + tid < P)
    energy[p0 + tid] = xch[tid] + xch[LT_M + tid] - (float)pa.n_pad * 0.6931471805599453f;
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace ehmc
