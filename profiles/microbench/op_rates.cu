// Issue-rate microbenchmark for the conversion / mixed-precision instructions the fp16-split
// epilogue of k_dense_tc3 depends on (sm_100a).  Each thread runs 8 independent dependency chains
// of one instruction; 1024 threads per CTA, one CTA per SM.  Prints warp-instructions per clock
// per SM sub-partition (1.0 = full issue rate).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o op_rates op_rates.cu && ./op_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, float seed) {
  float a[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + threadIdx.x * 1e-3f + i;
    u[i] = 0x3c003c00u + threadIdx.x + i;
  }
  const unsigned short one = 0x3c00, m1 = 0xbc00;
  unsigned long long w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f);
  double dd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dd[i] = (double)a[i];
  const double dseed = (double)seed;
  extern __shared__ double sm64[];  // [8][1025]
  if (OP >= 30 && OP <= 31) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm64[i * 1025 + threadIdx.x] = dd[i];
  }
  const unsigned long long wseed = ((unsigned long long)__float_as_uint(seed) << 32) | __float_as_uint(seed);
  long long t0, t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(seed));
      if (OP == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 2) asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(a[i]) : "h"(one));                 // FHADD
      if (OP == 3) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(one), "h"(m1));   // FHFMA
      if (OP == 4) asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %0; cvt.f32.f16 %1, lo;}" : "+r"(u[i]), "+f"(a[i]));  // HADD2.F32 unpack
      if (OP == 5) asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(a[i]));             // F2FP pack (independent)
      if (OP == 6) asm volatile("{.reg .b32 t; cvt.rn.f16x2.f32 t, %0, %0; mov.b32 %0, t;}" : "+f"(a[i]));  // F2FP chain
      if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(u[i]) : "r"(0x1234567u));
      if (OP == 8) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u[i]) : "r"(0x1234567u));
      if (OP == 9) asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(u[i]));
      if (OP == 10) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(0x1234567u));
      if (OP == 11) asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(u[i]) : "r"(0x3c003c00u));    // HFMA2
      if (OP == 12) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 13) asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(u[i]) : "f"(a[i]));
      if (OP == 14) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[i]) : "l"(wseed));  // FFMA2 (two FMAs per lane)
      if (OP == 16) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(wseed));           // FADD2
      if (OP == 17) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(wseed));           // FMUL2
      if (OP == 18) asm volatile("rsqrt.approx.f32 %0, %0;" : "+f"(a[i]));                         // MUFU.RSQ
      if (OP == 19) { asm volatile("rsqrt.approx.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[i]) : "l"(wseed)); asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[(i + 4) & 7]) : "l"(wseed)); }
      if (OP == 15) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 20) asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; xor.b32 lo, lo, hi; mul.wide.u32 %0, lo, 0xD2511F53;}" : "+l"(w[i]));  // IMAD.WIDE.U32 (+ LOP3)
      if (OP == 21) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w[i]) : "r"(u[i]));  // IMAD.WIDE.U32 independent
      if (OP == 22) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(u[i]) : "r"(0xD2511F53u));  // IMAD
      if (OP == 23) asm volatile("lg2.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 24) asm volatile("sin.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 25) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(a[i]) : "r"(u[i] + it));  // I2FP (+ IADD)
      if (OP == 27) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(dd[i]) : "d"(dseed) : "memory");               // DADD
      if (OP == 28) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(dd[i]) : "d"(dseed) : "memory");           // DFMA
      if (OP == 29) asm volatile("{.reg .f64 t; cvt.f64.f32 t, %0; cvt.rn.f32.f64 %0, t;}" : "+f"(a[i]));  // F2F.F64.F32 + F2F.F32.F64
      if (OP == 35) asm volatile("{.reg .f64 t; cvt.f64.f32 t, %1; add.rn.f64 %0, %0, t;}" : "+d"(dd[i]) : "f"(a[(i + it) & 7]));  // F2F.F64.F32 + DADD
      if (OP == 36) asm volatile("redux.sync.add.s32 %0, %0, 0xffffffff;" : "+r"(u[i]));  // REDUX
      if (OP == 37) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u[i]));  // SHFL
      if (OP == 30) asm volatile("{.reg .f64 t; ld.shared.f64 t, [%0]; add.rn.f64 t, t, %1; st.shared.f64 [%0], t;}" ::"r"((unsigned)__cvta_generic_to_shared(&sm64[i * 1025 + threadIdx.x])), "d"(dseed) : "memory");  // LDS.64 + DADD + STS.64
      if (OP == 31) asm volatile("{.reg .f64 t, c; cvt.f64.f32 c, %2; ld.shared.f64 t, [%0]; add.rn.f64 t, t, c; st.shared.f64 [%0], t;}" ::"r"((unsigned)__cvta_generic_to_shared(&sm64[i * 1025 + threadIdx.x])), "d"(dseed), "f"(a[i]) : "memory");  // the statistics update: F2F + LDS.64 + DADD + STS.64
      if (OP == 32) asm volatile("{add.u32 %0, %0, %1; shf.l.wrap.b32 %1, %1, %1, 13; xor.b32 %1, %1, %0;}" : "+r"(u[i]), "+r"(u[(i + 1) & 7]));  // Threefry mix: IADD + SHF + LOP
      if (OP == 33) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));              // IADD (register operands)
      if (OP == 34) asm volatile("{.reg .b32 e; shr.u32 e, %1, 3; add.u32 e, e, 0x38000000; mov.b64 %0, {%2, e};}" : "=d"(dd[i]) : "r"(u[i]), "r"(u[(i + 1) & 7]));  // float -> double bits by integer ops (2 ALU + move)
      if (OP == 26) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(0xD2511F53u));  // IMAD.HI.U32
    }
  }
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)dd[i] + a[i] + __uint_as_float(u[i]) + __uint_as_float((uint32_t)w[i]) + __uint_as_float((uint32_t)(w[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, float* out, long long* cyc) {
  const size_t smb = 8 * 1025 * sizeof(double);
  cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
  k<OP><<<148, 1024, smb>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  k<OP><<<148, 1024, smb>>>(out, cyc, 1.0f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  // 32 warps per SM = 8 per sub-partition, each issuing 8 * ITER instructions
  const double per_smsp = 8.0 * 8 * ITER / c;
  printf("%-28s %8.3f warp-instr/clk/SMSP  (%.0f cycles)  %s\n", name, per_smsp, c, cudaGetErrorString(e));
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("FFMA", out, cyc);
  run<1>("FADD", out, cyc);
  run<12>("FMUL", out, cyc);
  run<15>("FMNMX", out, cyc);
  run<2>("FHADD (add.f32.f16)", out, cyc);
  run<3>("FHFMA (fma.f32.f16)", out, cyc);
  run<4>("HADD2.F32 (cvt.f32.f16)", out, cyc);
  run<5>("F2FP.F16.F32.PACK_AB", out, cyc);
  run<6>("F2FP chain", out, cyc);
  run<11>("HFMA2", out, cyc);
  run<14>("FFMA2 (fma.f32x2)", out, cyc);
  run<16>("FADD2 (add.f32x2)", out, cyc);
  run<17>("FMUL2 (mul.f32x2)", out, cyc);
  run<18>("MUFU.RSQ", out, cyc);
  run<19>("MUFU.RSQ + 2 FFMA2 (per 3 instr)", out, cyc);
  run<13>("cvt.rna.tf32", out, cyc);
  run<7>("LOP3", out, cyc);
  run<8>("PRMT", out, cyc);
  run<9>("SHF", out, cyc);
  run<10>("IADD", out, cyc);
  run<20>("LOP3 + IMAD.WIDE.U32 (per pair)", out, cyc);
  run<21>("IMAD.WIDE.U32 independent", out, cyc);
  run<22>("IMAD (mad.lo.u32)", out, cyc);
  run<26>("IMAD.HI.U32", out, cyc);
  run<23>("MUFU.LG2", out, cyc);
  run<24>("MUFU.SIN", out, cyc);
  run<25>("I2FP.F32.U32 (+ IADD)", out, cyc);
  run<27>("DADD", out, cyc);
  run<28>("DFMA", out, cyc);
  run<29>("F2F.F64.F32 + F2F.F32.F64 (per pair)", out, cyc);
  run<35>("F2F.F64.F32 + DADD (per pair)", out, cyc);
  run<36>("REDUX.SUM.S32", out, cyc);
  run<37>("SHFL.BFLY", out, cyc);
  run<30>("LDS.64 + DADD + STS.64 (per group)", out, cyc);
  run<31>("F2F + LDS.64 + DADD + STS.64 (per group)", out, cyc);
  run<32>("IADD + SHF + LOP3 (per group)", out, cyc);
  run<33>("IADD (register operands)", out, cyc);
  run<34>("f32->f64 bits, integer ops (per group)", out, cyc);
  return 0;
}
