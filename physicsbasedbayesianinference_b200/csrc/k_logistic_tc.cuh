// K4-TC: Bayesian logistic-regression gradient (and energy) on the tensor cores -- BASELINE
// config 3 (X 100k x 256, 65 536 particles).  Flash-attention-shaped GEMM chain; the N x P
// logits never leave the SM and BOTH GEMMs take their A operand from tensor memory, so shared
// memory carries nothing but a deep ring of X chunks:
//
//   per CTA: 128 particles.  Theta tile -> bf16 pairs in TMEM (A of GEMM1, resident).
//   per chunk of 64 data rows (X chunk bf16 + y, ONE TMA bulk copy into an NS-stage ring):
//     GEMM1  S[128 x 64]   = Theta_tile . X_chunk^T       tcgen05.mma kind::f16, A = Theta (TMEM)
//     epilogue             r = sigmoid(S) - y  (tanh.approx: one MUFU), energy terms; r (bf16
//                          pairs) is written back INTO the S columns it came from (tcgen05.st)
//     GEMM2  G[128 x D]   += R . X_chunk                   A = R (TMEM, aliasing S), B = the SAME
//                                                          smem chunk read MN-major
//   TMEM columns: Theta [0, DP/2) | S0 [128,192) | S1 [192,256) | G [256, 256+DP)
//
// Why this shape (measured on the first version, which kept Theta and R in shared memory with a
// 2-stage ring of 128-row chunks: 43 % tensor-pipe active): (1) with A and B both in shared memory
// GEMM1 alone needs 128 B/clk, the whole shared-memory bandwidth, before TMA writes and the R
// round trip; (2) two stages leave no prefetch distance -- a stage is refilled only when GEMM2 of
// its chunk retires and is needed again by the very next GEMM1, so every chunk exposed the TMA
// latency.  Here shared-memory traffic is ~1/2 of its bandwidth and 6 stages are in flight.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue
// (thread <-> particle row and one half of the 64 logit columns).
// Shared-memory operand layout: canonical K-major no-swizzle UMMA layout [K/8][64 rows][8 bf16];
// an X chunk stored that way is simultaneously the K-major B operand of GEMM1 (rows = data rows,
// K = d) and the MN-major B operand of GEMM2 (N = d, K = data rows).
//
// Inputs are rounded to bf16 (fp32 accumulation): this is the throughput path; the exact
// fp32 / fp64 path is k_logistic.cuh.  A trajectory driven by this gradient is still reversible
// and volume preserving (the gradient is a deterministic function of q).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace ehmc {

constexpr int LT_M = 128;    // particles per CTA
constexpr int LT_NB = 64;    // data rows per chunk
constexpr int LT_MAX_STAGES = 8;
constexpr int LT_EPI_WARPS = 8;
constexpr int LT_THREADS = 32 * (2 + LT_EPI_WARPS);

struct LogisticTcArgs {
  const unsigned char* chunks;  // [NC] blocks of chunk_bytes: X part [DP/8][64][8] bf16, y[64] float, (1/2 - y)[64] half
  int NC;                       // number of chunks
  int DP;                       // D rounded up to 16
  int D;
  int n_pad;                    // zero rows appended to the last chunk (y = 0.5 there)
  unsigned chunk_bytes;
  int stages;                   // ring depth (shared memory permitting, <= LT_MAX_STAGES)
  int split;                    // 1, or 2: two CTAs per particle tile, each over half of the data rows,
                                // results combined with atomicAdd into zeroed outputs (0 + a + b is order independent)
  float inv_s2;
};

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
                                                               float* __restrict__ energy,
                                                               double* __restrict__ energy64, const LogisticTcArgs pa) {
  extern __shared__ __align__(128) unsigned char lt_smem[];
  const int DP = pa.DP, D = pa.D, NS = pa.stages;
  const int part = (int)(blockIdx.x % (unsigned)pa.split);
  const int cbeg = (int)((long long)pa.NC * part / pa.split);
  const int NC = (int)((long long)pa.NC * (part + 1) / pa.split) - cbeg;  // chunks of this CTA: cbeg .. cbeg + NC
  unsigned char* Xs0 = lt_smem;                                                     // NS stages of chunk_bytes
  double* xch = reinterpret_cast<double*>(Xs0 + (size_t)NS * pa.chunk_bytes);       // [2][128] energy exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * LT_M);
  uint64_t* x_full = bars;                      // [LT_MAX_STAGES]
  uint64_t* x_empty = bars + LT_MAX_STAGES;     // [LT_MAX_STAGES]
  uint64_t* s_full = bars + 2 * LT_MAX_STAGES;  // [2]
  uint64_t* r_full = s_full + 2;                // [2]
  uint64_t* g_done = r_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_done + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const long long p0 = (long long)(blockIdx.x / (unsigned)pa.split) * LT_M;

  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) {
        mbar_init(&x_full[i], 1);
        mbar_init(&x_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&r_full[i], LT_EPI_WARPS);
      }
      mbar_init(g_done, 1);
      fence_barrier_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t t_s = tmem_base + 128u, t_g = tmem_base + 256u;
  // a warp can only touch the TMEM lanes of ITS hardware quarter (warp index % 4)
  const int quarter = warp & 3, half = (warp - 2) >> 2;
  const int row = quarter * 32 + lane;
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;

  if (warp >= 2) {
    // Theta tile -> bf16 pairs in TMEM: this thread's particle row, dims [half * DP/2, (half + 1) * DP/2)
    const long long pi = p0 + row;
    const int d0 = half * (DP / 2);
    for (int b = 0; b < DP / 16; ++b) {  // 8 dims = 4 packed columns per batch (DP / 2 is a multiple of 8)
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = d0 + 8 * b + 2 * e;
        const float v0 = (d < D && pi < P) ? theta[d * t_ld + pi] : 0.f;
        const float v1 = (d + 1 < D && pi < P) ? theta[(d + 1) * t_ld + pi] : 0.f;
        w[e] = pack_bf16x2(v0, v1);
      }
      tmem_st4(tmem_base + lane_off + (uint32_t)(d0 / 2 + 4 * b), w);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int c = 0; c < NC; ++c) {
        const int s = c % NS;
        mbar_wait(&x_empty[s], (uint32_t)(((c / NS) & 1) ^ 1));
        mbar_expect_tx(&x_full[s], pa.chunk_bytes);
        tma_bulk_g2s(Xs0 + (size_t)s * pa.chunk_bytes, pa.chunks + (size_t)(cbeg + c) * pa.chunk_bytes, pa.chunk_bytes,
                     &x_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc1 = umma_idesc_bf16(LT_M, LT_NB, 0);
      const uint32_t idesc2 = umma_idesc_bf16(LT_M, DP, 1);
      const uint64_t k_step = (2u * LT_NB * 16u) >> 4;  // GEMM1: two 16-byte K chunks (of 64 rows) per MMA
      auto gemm2 = [&](int cc) {
        const int s = cc % NS, b = cc & 1;
        mbar_wait(&r_full[b], (uint32_t)((cc >> 1) & 1));
        tc_fence_after();
        // B = X chunk read MN-major: N = d (LT_NB*16 B between 8-d groups = SBO),
        // K = data rows (16 B apart, groups of 8 rows 128 B apart = LBO)
        const uint64_t dB2 = umma_desc2(smem_u32(Xs0 + (size_t)s * pa.chunk_bytes), 128, LT_NB * 16);
        // A = R: bf16 pairs of data rows (2k, 2k+1); half h of the epilogue wrote its 32 rows into the
        // first 16 columns of ITS 32-column half of the S buffer
#pragma unroll
        for (int j = 0; j < LT_NB / 16; ++j)
          umma_f16_ts(t_g, t_s + (uint32_t)(b * 64 + (j >> 1) * 32 + (j & 1) * 8), dB2 + j * ((16u * 16u) >> 4), idesc2,
                       (cc > 0 || j > 0) ? 1u : 0u);
        umma_commit(&x_empty[s]);  // the X stage is free once these MMAs have executed
      };
      for (int c = 0; c < NC; ++c) {
        const int s = c % NS;
        mbar_wait(&x_full[s], (uint32_t)((c / NS) & 1));
        tc_fence_after();
        // S buffer c & 1 was last read (as R) by GEMM2 of chunk c - 2, issued before this point: the
        // tensor pipe executes one thread's MMAs in order, no barrier needed
        const uint64_t dB1 = umma_desc2(smem_u32(Xs0 + (size_t)s * pa.chunk_bytes), LT_NB * 16, 128);
        for (int j = 0; j < DP / 16; ++j)
          umma_f16_ts(t_s + (uint32_t)((c & 1) * 64), tmem_base + 8u * j, dB1 + j * k_step, idesc1, j > 0 ? 1u : 0u);
        umma_commit(&s_full[c & 1]);
        if (c >= 1) gemm2(c - 1);
      }
      gemm2(NC - 1);
      umma_commit(g_done);
    }
  } else {
    // ===== epilogue warps: sigmoid-residual between the two GEMMs =====
    double Uacc = 0.0;  // 32-term float sums per chunk, double across the chunks
    const uint32_t t_mine = t_s + lane_off + (uint32_t)(half * 32);
    for (int c = 0; c < NC; ++c) {
      const int s = c % NS, b = c & 1;
      mbar_wait(&s_full[b], (uint32_t)((c >> 1) & 1));
      tc_fence_after();
      uint32_t sv[2][16];
      tmem_ld16_issue(t_mine + (uint32_t)(b * 64), sv[0]);
      tmem_ld16_issue(t_mine + (uint32_t)(b * 64 + 16), sv[1]);
      tmem_wait_ld16(sv[0]);
      tmem_wait_ld16(sv[1]);
      const unsigned char* tail = Xs0 + (size_t)s * pa.chunk_bytes + (size_t)DP * LT_NB * 2;
      const float* yv = reinterpret_cast<const float*>(tail) + half * 32;
      uint32_t rp[16];
      float Uc = 0.f;
      if (WITH_E) {
        // energy evaluations (first / last gradient of a trajectory): exact exp / log per logit
#pragma unroll
        for (int bb = 0; bb < 2; ++bb)
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float r2[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const float sgm = __uint_as_float(sv[bb][i + t]);
              const float yn = yv[16 * bb + i + t];
              const float sig = fmaf(0.5f, tanh_approx(0.5f * sgm), 0.5f);
              r2[t] = sig - yn;
              Uc += fmaxf(sgm, 0.f) + __logf(1.f + __expf(-fabsf(sgm))) - yn * sgm;
            }
            rp[(16 * bb + i) / 2] = pack_bf16x2(r2[0], r2[1]);
          }
      } else {
        // gradient only: two logits per MUFU op.  r = 1/2 tanh(s / 2) + (1/2 - y) in half2 (tanh.approx.f16x2,
        // HFMA2); r is rounded to bf16 (8 significant bits) for GEMM2 anyway, fp16 (11 bits) loses nothing.
        // (1/2 - y) comes pre-packed as half2 with the chunk.
        const uint4* cv = reinterpret_cast<const uint4*>(tail + LT_NB * 4) + half * 4;  // 32 halves = 4 x 16 B
        const __half2 half_ = __float2half2_rn(0.5f);
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          const uint4 c4 = cv[g4];
          const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 4 * g4 + e;  // pair index: logits 2k, 2k + 1 of my 32
            const float s0 = __uint_as_float(sv[k >> 3][(2 * k) & 15]), s1 = __uint_as_float(sv[k >> 3][(2 * k + 1) & 15]);
            const __half2 x = __hmul2(__floats2half2_rn(s0, s1), half_);
            uint32_t xb = *reinterpret_cast<const uint32_t*>(&x), tb;
            asm("tanh.approx.f16x2 %0, %1;" : "=r"(tb) : "r"(xb));
            const __half2 r = __hfma2(*reinterpret_cast<const __half2*>(&tb), half_, *reinterpret_cast<const __half2*>(&cw[e]));
            const float2 rf = __half22float2(r);
            rp[k] = pack_bf16x2(rf.x, rf.y);
          }
        }
      }
      if (WITH_E) Uacc += (double)Uc;
      tmem_st16(t_mine + (uint32_t)(b * 64), rp);  // R over the first 16 of my 32 S columns
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&r_full[b]);
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
          const float th = part == 0 ? theta[d * t_ld + pi] : 0.f;  // prior term: once per particle
          t2 = fmaf(th, th, t2);
          const float gd = __uint_as_float(gv[i]) + th * pa.inv_s2;
          if (grad) {
            if (pa.split == 1)
              grad[d * g_ld + pi] = gd;
            else
              atomicAdd(&grad[d * g_ld + pi], gd);
          }
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
          const float th = part == 0 ? theta[d * t_ld + pi] : 0.f;
          t2 = fmaf(th, th, t2);
          const float gd = __uint_as_float(gv[i]) + th * pa.inv_s2;
          if (grad) {
            if (pa.split == 1)
              grad[d * g_ld + pi] = gd;
            else
              atomicAdd(&grad[d * g_ld + pi], gd);
          }
        }
      }
    }
    if (WITH_E) xch[half * LT_M + row] = Uacc + (double)(0.5f * t2 * pa.inv_s2);
  }
  tc_fence_before();
  __syncthreads();
  if (WITH_E && tid < LT_M && p0 + tid < P) {
    // the zero rows padding the LAST chunk each contributed softplus(0) = ln 2
    const double pad = part == pa.split - 1 ? (double)pa.n_pad * 0.6931471805599453 : 0.0;
    const double ev = xch[tid] + xch[LT_M + tid] - pad;
    if (pa.split == 1) {
      if (energy) energy[p0 + tid] = (float)ev;
      if (energy64) energy64[p0 + tid] = ev;
    } else {
      if (energy) atomicAdd(&energy[p0 + tid], (float)ev);
      if (energy64) atomicAdd(&energy64[p0 + tid], ev);
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace ehmc
