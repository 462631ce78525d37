// Throughput microbenchmark of the GELU-epilogue instruction mix on sm_100a: warp-instructions per clock per SMSP
// for MUFU.TANH (f32 / f16x2), MUFU.EX2, FFMA, HFMA2, the f32<->f16/bf16 pack / unpack conversions, and mixes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#define OPS 8
template <int OP>
__device__ __forceinline__ void op(uint32_t& r, uint32_t& q) {
  if (OP == 0) { float x = __uint_as_float(r), y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); r = __float_as_uint(y); }
  if (OP == 1) { uint32_t y; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(r)); r = y; }
  if (OP == 2) { float x = __uint_as_float(r), y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); r = __float_as_uint(y); }
  if (OP == 3) { uint32_t y; asm volatile("fma.rn.f16x2 %0, %1, %1, %2;" : "=r"(y) : "r"(r), "r"(q)); r = y; }
  if (OP == 4) { float x = __uint_as_float(r), y; asm volatile("fma.rn.f32 %0, %1, %1, %2;" : "=f"(y) : "f"(x), "f"(__uint_as_float(q))); r = __float_as_uint(y); }
  if (OP == 5) { float x = __uint_as_float(r), y; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(__uint_as_float(q))); r = __float_as_uint(y); }
  if (OP == 6) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(__uint_as_float(r)), "f"(__uint_as_float(q))); r = y; }
  if (OP == 7) { uint32_t y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(__uint_as_float(r)), "f"(__uint_as_float(q))); r = y; }
  if (OP == 8) { float y; asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi;}" : "=f"(y) : "r"(r)); r = __float_as_uint(y); }
  if (OP == 9) { uint32_t y; asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(y) : "r"(r), "r"(q)); r = y; }
  if (OP == 10) { uint32_t y; asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(r), "r"(q)); r = y; }
  if (OP == 11) { uint32_t y; asm volatile("fma.rn.bf16x2 %0, %1, %1, %2;" : "=r"(y) : "r"(r), "r"(q)); r = y; }
}
// A: 8 chains of OPA; B (optional, -1 = none): 8 more chains of OPB interleaved
template <int OPA, int OPB>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t r[OPS], s[OPS], q = seed * 3 + 1;
#pragma unroll
  for (int j = 0; j < OPS; ++j) { r[j] = seed + threadIdx.x * 8 + j; s[j] = r[j] ^ 0x1234u; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < OPS; ++j) {
      op<OPA>(r[j], q);
      if (OPB >= 0) op<(OPB < 0 ? 0 : OPB)>(s[j], q);
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < OPS; ++j) x ^= r[j] ^ s[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int OPA, int OPB>
void run(const char* name) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t* out; cudaMalloc(&out, sms * 1024 * 4);
  const int iters = 4096;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<OPA, OPB><<<sms, 1024>>>(out, 64, 1);
  cudaEventRecord(a);
  k<OPA, OPB><<<sms, 1024>>>(out, iters, 1);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double winstr = (double)sms * 32 * iters * OPS * (OPB >= 0 ? 2 : 1);      // warp instructions
  printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SMSP (at %d MHz nominal)\n", name, ms, winstr / (ms * 1e-3) / (sms * 4) / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  run<0, -1>("tanh.approx.f32 (1 MUFU)");
  run<1, -1>("tanh.approx.f16x2 (2 MUFU + PRMT)");
  run<2, -1>("ex2.approx.f32");
  run<4, -1>("fma.f32");
  run<5, -1>("mul.f32");
  run<3, -1>("fma.f16x2");
  run<10, -1>("mul.f16x2");
  run<11, -1>("fma.bf16x2");
  run<6, -1>("cvt.rn.f16x2.f32 (F2FP)");
  run<7, -1>("cvt.rn.bf16x2.f32 (F2FP)");
  run<8, -1>("cvt.f32.f16 (hi half)");
  run<9, -1>("prmt");
  run<4, 3>("fma.f32 + fma.f16x2");
  run<4, 0>("fma.f32 + tanh.f32");
  run<3, 0>("fma.f16x2 + tanh.f32");
  run<4, 6>("fma.f32 + F2FP");
  run<3, 6>("fma.f16x2 + F2FP");
  run<4, 9>("fma.f32 + prmt");
  return 0;
}
