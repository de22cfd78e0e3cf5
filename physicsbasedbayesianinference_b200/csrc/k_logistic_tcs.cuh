// K4-TCS: Bayesian logistic-regression gradient (and energy) on the tensor cores at FLOAT32 accuracy --
// BASELINE config 3 (X 100k x 256, 65 536 particles) inside the north_star's 1e-5 trajectory tolerance.
//
// Same flash-attention-shaped chain as k_logistic_tc.cuh (S = Theta Xc^T -> r = sigmoid(S) - y -> G += R Xc, logits
// never leave the SM), but every operand is a round-to-nearest fp16 pair, v = hi + lo with |v - hi - lo| <= 2^-24 |v|,
// and each GEMM runs the three products that matter (the lo x lo term is 2^-22 relative):
//
//   GEMM1  S  = Th_hi Xl^T + Th_lo Xh^T + Th_hi Xh^T        48 MMAs (M 128, N 64, K 16) per chunk at D = 256
//   GEMM2  G += R_hi Xl    + R_lo Xh    + R_hi Xh           12 MMAs (M 128, N 256, K 16)
//
// fp16 has 5 exponent bits, so everything is scaled by exact powers of two that are undone in fp32:
//   X       by 2^a (host, max |X| -> [2^7, 2^8)),
//   Theta   per particle row by 2^e (prologue, max |theta_row| -> [2^7, 2^8)),
//   r       by 2^10 (r in (-1, 1); lo keeps absolute precision 2^-24 * 2^-10 down to r = 0).
//
// The tensor core TRUNCATES when it adds into its fp32 accumulator (measured, profiles/r02_logistic_split_probe.txt:
// a gradient accumulated over all 1563 chunks in tensor memory came out 1.6e-4 low at N = 100 000, linear in N,
// -3.3e-8 of the running sum per MMA).  Hence (i) the small products run first in each chain, (ii) G is accumulated
// in tensor memory over windows of LTS_FLUSH = 4 chunks only (48 MMAs from zero: -8e-7) and each window is added,
// round-to-nearest, to the float32 master copy that the epilogue threads keep in REGISTERS (128 per thread; the
// epilogue warpgroups take 224 registers with setmaxnreg).  The read of a finished window (tcgen05.ld, 128 KB) runs
// under GEMM1 of the next chunk; GEMM2 of the next window waits for it.
//
// Tensor memory (512 columns, all used at D = 256):
//   Theta_hi [0, DP/2) | S0 [128, 192) | S1 [192, 256) | G window [256, 256 + DP)
// Theta_hi, Theta_lo, S and G would need 640 columns, so Theta_lo lives in SHARED memory (canonical K-major layout
// [DP/8][128][8], 64 KB) and its product is the one SS-mode pass (A and B both from shared memory: 192 B/clk against
// the 128 B/clk the SM delivers, i.e. that pass runs at 2/3 rate; 1/6 of the MMAs).  R_hi / R_lo (fp16 pairs, 2 x 16
// columns per half) are written back over the 32 S columns each epilogue thread read them from.
//
// Shared memory: Theta_lo + a ring of HALF chunks (entry 2c = Xl of chunk c, entry 2c + 1 = Xh + 2^10 y), one
// cp.async.bulk each.  At D = 256 five entries fit (33 KB each): GEMM1 of chunk c and GEMM2 of chunk c - 1 hold
// four, the fifth is the prefetch.  GEMM2 retires its Xl pass first, which frees the slot the ring order gives to
// Xh of chunk c + 1 about 1500 cycles before GEMM1's second pass needs it.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (warps 2, 3 idle: setmaxnreg works per
// warpgroup), warps 4..11 = epilogue (thread <-> particle row and one half of the 64 logit columns / of G's columns).
// sigmoid = rcp(1 + ex2(-s log2 e)): two MUFU operations per logit (|error| ~ 2e-7; tanh.approx, which the bf16
// kernel uses, is good to 2^-11 only).
//
// Reference arithmetic replaced: the gradient callable of src/integrator.py:61-73 inside the loop of :105-120, and
// the potential of src/HMC.py:106-116, for the build-defined logistic-regression model.
#pragma once

#include "k_logistic_tc.cuh"  // LT_* constants

namespace ehmc {

constexpr int LTS_MIN_STAGES = 4;  // GEMM1(c) + GEMM2(c - 1) hold four half-chunk entries
constexpr int LTS_FLUSH = 4;       // chunks per tensor-memory accumulation window of G
constexpr int LTS_THREADS = 384;   // warpgroup 0: TMA + MMA, warpgroups 1, 2: epilogue
constexpr float LTS_RSCALE = 1024.f;

struct LogisticTcsArgs {
  const unsigned char* entries;  // [2 NC] blocks of entry_bytes: [DP/8][64][8] fp16, then 64 floats (2^10 y; Xh entries)
  int NC;                        // chunks of 64 data rows
  int DP;                        // D rounded up to 16
  int D;
  int n_pad;                     // zero rows appended to the last chunk (y = 0.5 there)
  unsigned entry_bytes;
  int stages;                    // ring depth in entries
  int split;                     // 1, or 2: two CTAs per particle tile, each over half of the data rows
  float inv_s2;
  float x_iscale;                // 2^-a
};

// 16 logits of one epilogue thread: S (tensor memory) -> 2^10 (sigmoid - y) as fp16 (hi, lo) pair words
template <bool WITH_E>
__device__ __forceinline__ void lts_residual16(const uint32_t (&sv)[16], const float4* ky4, float c_row, float c2,
                                               uint32_t (&rh)[8], uint32_t (&rl)[8], float& Uc) {
#pragma unroll
  for (int g4 = 0; g4 < 4; ++g4) {
    const float4 k4 = ky4[g4];  // 2^10 y of 4 consecutive logits (warp-wide broadcast)
    const float ky[4] = {k4.x, k4.y, k4.z, k4.w};
    float rk[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float S = __uint_as_float(sv[4 * g4 + t]);
      if (WITH_E) {
        // energy evaluations (first / last gradient of a trajectory): softplus(s) - y s with
        // t = e^{-|s|}, softplus = max(s, 0) + ln(1 + t), sigmoid = s >= 0 ? 1 / (1 + t) : t / (1 + t)
        const float s = S * c_row;
        float tt, inv, lg;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(tt) : "f"(-1.4426950408889634f * fabsf(s)));
        const float opt = 1.f + tt;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(opt));
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(opt));
        const float sig = s >= 0.f ? inv : tt * inv;
        rk[t] = fmaf(LTS_RSCALE, sig, -ky[t]);
        Uc += fmaf(0.6931471805599453f, lg, fmaxf(s, 0.f)) - (ky[t] * (1.f / LTS_RSCALE)) * s;
      } else {
        float ex, sig;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(S * c2));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sig) : "f"(1.f + ex));
        rk[t] = fmaf(LTS_RSCALE, sig, -ky[t]);
      }
    }
    split16(rk[0], rk[1], rh[2 * g4], rl[2 * g4]);
    split16(rk[2], rk[3], rh[2 * g4 + 1], rl[2 * g4 + 1]);
  }
}

// grad[D,P] (and energy when WITH_E) at theta[D,P].  energy (float) and energy64 (double) are each optional.
template <bool WITH_E>
__global__ void __launch_bounds__(LTS_THREADS, 1) k_logistic_tcs(const float* __restrict__ theta, long long t_ld,
                                                                 long long P, float* __restrict__ grad, long long g_ld,
                                                                 float* __restrict__ energy,
                                                                 double* __restrict__ energy64,
                                                                 const LogisticTcsArgs pa) {
  extern __shared__ __align__(128) unsigned char lts_smem[];
  const int DP = pa.DP, D = pa.D, NS = pa.stages;
  const int part = (int)(blockIdx.x % (unsigned)pa.split);
  const int cbeg = (int)((long long)pa.NC * part / pa.split);
  const int NC = (int)((long long)pa.NC * (part + 1) / pa.split) - cbeg;  // chunks of this CTA: cbeg .. cbeg + NC
  unsigned char* Tlo = lts_smem;                               // Theta_lo [DP/8][128][8] fp16
  unsigned char* ring = Tlo + (size_t)DP * LT_M * 2;           // NS entries
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)NS * pa.entry_bytes);
  uint64_t* x_full = bars;                      // [LT_MAX_STAGES]
  uint64_t* x_empty = bars + LT_MAX_STAGES;     // [LT_MAX_STAGES]
  uint64_t* s_full = bars + 2 * LT_MAX_STAGES;  // [2]
  uint64_t* r_full = s_full + 2;                // [2]
  uint64_t* g_ready = r_full + 2;               // a window of G is complete in tensor memory
  uint64_t* g_free = g_ready + 1;               // ... and has been read by all epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_free + 1);
  double* xch = reinterpret_cast<double*>(Tlo);  // [2][128] energy exchange; Theta_lo is dead by then

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
      mbar_init(g_ready, 1);
      mbar_init(g_free, LT_EPI_WARPS);
      fence_barrier_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t t_s = tmem_base + 128u, t_g = tmem_base + 256u;

  if (warp < 4) {
    // ===== warpgroup 0: TMA producer (warp 0) and MMA issuer (warp 1) =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    auto produce = [&](int e0, int e1) {  // ring entries e0 .. e1 - 1
      for (int e = e0; e < e1; ++e) {
        const int s = e % NS;
        mbar_wait(&x_empty[s], (uint32_t)(((e / NS) & 1) ^ 1));  // first pass: a fresh barrier reads as "free"
        mbar_expect_tx(&x_full[s], pa.entry_bytes);
        tma_bulk_g2s(ring + (size_t)s * pa.entry_bytes, pa.entries + (size_t)(2 * cbeg + e) * pa.entry_bytes,
                     pa.entry_bytes, &x_full[s]);
      }
    };
    // the ring fills while the epilogue warps build the Theta operands
    const int e_pre = 2 * NC < NS ? 2 * NC : NS;
    if (warp == 0 && elect_one()) produce(0, e_pre);
    tc_fence_before();
    __syncthreads();  // Theta_hi / Theta_lo are in place
    tc_fence_after();
    if (warp == 0) {
      if (elect_one()) produce(e_pre, 2 * NC);
    } else if (warp == 1) {
      if (elect_one()) {
        const uint32_t idesc1 = umma_idesc_16(LT_M, LT_NB, 0, 0);
        const uint32_t idesc2 = umma_idesc_16(LT_M, DP, 0, 1);
        const uint64_t k_step = (2u * LT_NB * 16u) >> 4;  // GEMM1 B: two 16-byte K chunks (of 64 rows) per MMA
        const uint64_t a_step = (2u * LT_M * 16u) >> 4;   // Theta_lo: two K chunks of 128 rows
        const uint64_t r_step = (16u * 16u) >> 4;         // GEMM2 B (MN-major): 16 data rows
        const uint64_t dA = umma_desc2(smem_u32(Tlo), LT_M * 16, 128);
        const int K16 = DP / 16;
        auto entry = [&](int e) { return smem_u32(ring + (size_t)(e % NS) * pa.entry_bytes); };
        auto gemm2 = [&](int cc) {
          const int b = cc & 1, w = cc / LTS_FLUSH;
          const bool w_first = cc % LTS_FLUSH == 0, w_last = cc % LTS_FLUSH == LTS_FLUSH - 1 || cc == NC - 1;
          mbar_wait(&r_full[b], (uint32_t)((cc >> 1) & 1));
          if (w_first && w > 0) mbar_wait(g_free, (uint32_t)((w - 1) & 1));  // the previous window has been read out
          tc_fence_after();
          // B = X half chunk read MN-major: N = d (LT_NB*16 B between 8-d groups = SBO),
          // K = data rows (16 B apart, groups of 8 rows 128 B apart = LBO)
          const uint64_t dBl = umma_desc2(entry(2 * cc), 128, LT_NB * 16);
          const uint64_t dBh = umma_desc2(entry(2 * cc + 1), 128, LT_NB * 16);
          // A = R: fp16 pairs of data rows (2k, 2k+1); half h of the epilogue wrote R_hi of its 32 rows into the
          // first 16 columns of ITS 32-column half of the S buffer and R_lo into the other 16
          const uint32_t t_r = t_s + (uint32_t)(b * 64);
#pragma unroll
          for (int j = 0; j < LT_NB / 16; ++j)
            umma_f16_ts(t_g, t_r + (uint32_t)((j >> 1) * 32 + (j & 1) * 8), dBl + j * r_step, idesc2,
                        (!w_first || j > 0) ? 1u : 0u);
          umma_commit(&x_empty[(2 * cc) % NS]);  // Xl of this chunk is free once the MMAs above have executed
#pragma unroll
          for (int j = 0; j < LT_NB / 16; ++j)
            umma_f16_ts(t_g, t_r + (uint32_t)((j >> 1) * 32 + (j & 1) * 8 + 16), dBh + j * r_step, idesc2, 1u);
#pragma unroll
          for (int j = 0; j < LT_NB / 16; ++j)
            umma_f16_ts(t_g, t_r + (uint32_t)((j >> 1) * 32 + (j & 1) * 8), dBh + j * r_step, idesc2, 1u);
          umma_commit(&x_empty[(2 * cc + 1) % NS]);
          if (w_last) umma_commit(g_ready);
        };
        for (int c = 0; c < NC; ++c) {
          const uint32_t t_sc = t_s + (uint32_t)((c & 1) * 64);
          mbar_wait(&x_full[(2 * c) % NS], (uint32_t)(((2 * c) / NS) & 1));
          tc_fence_after();
          // S buffer c & 1 was last read (as R) by GEMM2 of chunk c - 2, issued before this point: the
          // tensor pipe executes one thread's MMAs in order, no barrier needed
          const uint64_t dBl = umma_desc2(entry(2 * c), LT_NB * 16, 128);
          for (int j = 0; j < K16; ++j) umma_f16_ts(t_sc, tmem_base + 8u * j, dBl + j * k_step, idesc1, j > 0 ? 1u : 0u);
          mbar_wait(&x_full[(2 * c + 1) % NS], (uint32_t)(((2 * c + 1) / NS) & 1));
          tc_fence_after();
          const uint64_t dBh = umma_desc2(entry(2 * c + 1), LT_NB * 16, 128);
          for (int j = 0; j < K16; ++j) umma_f16_ss(t_sc, dA + j * a_step, dBh + j * k_step, idesc1, 1u);
          for (int j = 0; j < K16; ++j) umma_f16_ts(t_sc, tmem_base + 8u * j, dBh + j * k_step, idesc1, 1u);
          umma_commit(&s_full[c & 1]);
          if (c >= 1) gemm2(c - 1);
        }
        gemm2(NC - 1);
      }
    }
  } else {
    // ===== warpgroups 1, 2: Theta operands, sigmoid-residual between the two GEMMs, float32 master copy of G =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // a warp can only touch the TMEM lanes of ITS hardware quarter (warp index % 4)
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const long long pi = p0 + row;
    const bool valid = pi < P;
    // Theta tile -> (hi, lo) fp16 pairs of theta * 2^e: hi into TMEM, lo into shared memory.  The row scale needs
    // the row's max over ALL dims, so both threads of a row scan the whole row first (L1/L2 hits the second time).
    float amax = 0.f;
    if (valid)
      for (int d = 0; d < D; ++d) amax = fmaxf(amax, fabsf(theta[d * t_ld + pi]));
    uint32_t eb = (__float_as_uint(amax) >> 23) & 0xffu;
    eb = (eb < 16u || eb > 240u) ? 134u : eb;                      // zero / tiny / non-finite rows: no scaling
    const float sc = __uint_as_float((261u - eb) << 23);           // 2^(7 - (eb - 127))
    const float c_row = __uint_as_float((eb - 7u) << 23) * pa.x_iscale;  // logit = S * c_row  (2^-e 2^-a)
    {
      const int d0 = half * (DP / 2);
      uint4* lo4 = reinterpret_cast<uint4*>(Tlo);
      for (int b = 0; b < DP / 16; ++b) {  // 8 dims = 4 packed columns per batch (DP / 2 is a multiple of 8)
        uint32_t wh[4], wl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int d = d0 + 8 * b + 2 * e;
          const float v0 = (d < D && valid) ? theta[d * t_ld + pi] * sc : 0.f;
          const float v1 = (d + 1 < D && valid) ? theta[(d + 1) * t_ld + pi] * sc : 0.f;
          split16_sat(v0, v1, wh[e], wl[e]);
        }
        tmem_st4(tmem_base + lane_off + (uint32_t)(d0 / 2 + 4 * b), wh);
        lo4[(size_t)(d0 / 8 + b) * LT_M + row] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
      }
      tmem_wait_st();
      fence_proxy_async();  // Theta_lo was written through the generic proxy, the tensor core reads it
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int dh = DP / 2;  // G columns of this half: [half * dh, (half + 1) * dh)
    float gacc[128];        // float32 master copy of this thread's G columns (scaled by 2^(10 + a))
#pragma unroll
    for (int i = 0; i < 128; ++i) gacc[i] = 0.f;
    const uint32_t t_gm = t_g + lane_off + (uint32_t)(half * dh);
    auto flush = [&](int w) {  // gacc += window w of G
      mbar_wait(g_ready, (uint32_t)(w & 1));
      tc_fence_after();
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (16 * b + 16 <= dh) {
          uint32_t gv[16];
          tmem_ld16_issue(t_gm + (uint32_t)(16 * b), gv);
          tmem_wait_ld16(gv);
#pragma unroll
          for (int i = 0; i < 16; ++i) gacc[16 * b + i] += __uint_as_float(gv[i]);
        } else if (16 * b + 8 <= dh) {
          uint32_t gv[8];
          tmem_ld8_issue2(t_gm + (uint32_t)(16 * b), gv);
          tmem_wait_ld8(gv);
#pragma unroll
          for (int i = 0; i < 8; ++i) gacc[16 * b + i] += __uint_as_float(gv[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_free);
    };

    double Uacc = 0.0;
    const uint32_t t_mine = t_s + lane_off + (uint32_t)(half * 32);
    const float c2 = -1.4426950408889634f * c_row;  // ex2(S * c2) = exp(-logit)
    for (int c = 0; c < NC; ++c) {
      const int b = c & 1;
      mbar_wait(&s_full[b], (uint32_t)((c >> 1) & 1));
      // the 2^10 y tail of this chunk's Xh entry was written by the TMA unit: observe its barrier ourselves
      mbar_wait(&x_full[(2 * c + 1) % NS], (uint32_t)(((2 * c + 1) / NS) & 1));
      tc_fence_after();
      const float4* ky4 = reinterpret_cast<const float4*>(ring + (size_t)((2 * c + 1) % NS) * pa.entry_bytes +
                                                          (size_t)DP * LT_NB * 2) + half * 8;
      uint32_t sv[2][16], rh[8], rl[8];
      float Uc = 0.f;
      tmem_ld16_issue(t_mine + (uint32_t)(b * 64), sv[0]);
      tmem_ld16_issue(t_mine + (uint32_t)(b * 64 + 16), sv[1]);
      tmem_wait_ld16(sv[0]);
      tmem_wait_ld16(sv[1]);  // all 32 of my S columns are in registers before R overwrites them
      // R_hi over the first 16 of my 32 S columns (8 per 16 logits), R_lo over the other 16
      lts_residual16<WITH_E>(sv[0], ky4, c_row, c2, rh, rl, Uc);
      tmem_st8(t_mine + (uint32_t)(b * 64), rh);
      tmem_st8(t_mine + (uint32_t)(b * 64 + 16), rl);
      lts_residual16<WITH_E>(sv[1], ky4 + 4, c_row, c2, rh, rl, Uc);
      tmem_st8(t_mine + (uint32_t)(b * 64 + 8), rh);
      tmem_st8(t_mine + (uint32_t)(b * 64 + 24), rl);
      if (WITH_E) Uacc += (double)Uc;  // 32-term float partial sums, accumulated in double over the chunks
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&r_full[b]);
      // GEMM2 of chunk c - 1 was issued behind GEMM1 of this chunk: if it closed a window, read that window now
      if (c >= 1 && (c - 1) % LTS_FLUSH == LTS_FLUSH - 1) flush((c - 1) / LTS_FLUSH);
    }
    flush((NC - 1) / LTS_FLUSH);  // the last window (closed by GEMM2 of chunk NC - 1)

    // ---- prior term, stores --------------------------------------------------------------------
    const float gscale = pa.x_iscale * (1.f / LTS_RSCALE);
    float t2 = 0.f;
#pragma unroll
    for (int i = 0; i < 128; ++i) {
      const int d = half * dh + i;
      if (i < dh && d < D && valid) {
        const float th = part == 0 ? theta[d * t_ld + pi] : 0.f;  // prior term: once per particle
        t2 = fmaf(th, th, t2);
        const float gd = fmaf(gacc[i], gscale, th * pa.inv_s2);
        if (grad) {
          if (pa.split == 1)
            grad[d * g_ld + pi] = gd;
          else
            atomicAdd(&grad[d * g_ld + pi], gd);
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
