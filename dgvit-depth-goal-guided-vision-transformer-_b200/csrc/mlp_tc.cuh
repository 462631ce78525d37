// mlp_tc.cuh — fused GELU-MLP for sm_100a, forward and backward (FeedForward.forward + residual,
// vn/GoalFormer.py:39-50,104):
//
//     out = x + W2 gelu(W1 LN(x) + b1) + b2
//
// The 2048-wide hidden activation never leaves the SM in either direction.
//
// forward (mlp_fwd_tc_kernel<SPLIT, FRONT>): one CTA per 128-token tile.  For each 128-column hidden chunk GEMM1
// (tcgen05, K = 64) lands in TMEM, two ping-pong groups of 8 epilogue warps add the bias, apply GELU and write the
// chunk straight into a 128B-swizzled shared-memory tile that is the A operand of GEMM2, whose [128 x 64]
// fp32 accumulator stays in TMEM across all chunks.  Weight chunks stream through TMA rings (L2
// resident: 512 KB per layer).  Nothing but x and the output touches HBM: the backward recomputes
// the pre-activation instead of reading a saved copy (a K = 64 GEMM is cheaper than 4 KB/token).
// The kernel also absorbs what surrounds the MLP in a transformer block (vn/GoalFormer.py:81-82,103-104):
//   FRONT  prologue: x_m = to_out(o) + x_a ; LN2(x_m) written straight into GEMM1's A tile (no LN(x) input needed)
//   output stage   : the next block's LayerNorm-1 of the result (fp32 x, bf16 LN(x), mean, rstd)
//   SPLIT          : few token tiles -> a cluster of 8 CTAs shares a tile's hidden columns, reduce-scatter over DSMEM
//
// backward (mlp_bwd_tc_kernel<MODE>): two launches of one kernel over (token tile, hidden chunk) pairs.
// Per pair, from the smem images of X (LN output), dY, W1[chunk], W2[:, chunk]:
//     Hpre = X W1c^T (+b1)        dG = dY W2c           (both into TMEM)
//     G = gelu(Hpre)              dH = dG * gelu'(Hpre) (epilogue warps -> swizzled smem tiles)
//   MODE 0 (CTA = token tile, loops over chunks):  dX += dH W1c          (+ colsum(dY) for db2)
//   MODE 1 (CTA = hidden chunk x token range):     dW1c += dH^T X ; dW2c^T += G^T dY ; db1c += dH^T 1
// so neither dX nor dW needs a cross-CTA reduction beyond the few token ranges of MODE 1, and the
// [tokens x 2048] tensors G / dH never exist in HBM.  One smem image serves as K-major operand of
// one GEMM and MN-major operand of another (same trick as attn_tc.cuh).
//
// The GELU epilogue is the bound of all three kernels (MUFU + FMA pipe + issue slots), so it uses the cheapest
// form that stays below bf16 resolution: 0.5x(1 + tanh(x(a + b x^2))) with one MUFU.TANH per element and a, b
// fitted to the exact erf GELU (|err| <= 2.7e-4), evaluated in packed f16x2 (the f16 pipeline's rms error is 4x
// below that of rounding the exact GELU to bf16).  Measured on B200: MUFU.TANH is 16 results/clk/SM in every
// variant and HFMA2 issues at half the FFMA rate (profiles/micro/mufu_rate.cu), yet the same epilogue written in
// fp32 runs 10-13 % slower end to end (general three-register FFMAs do not sustain the full rate), so f16x2 stays.
// tcgen05 kind::f16 rejects mixed f16 / bf16 operands (illegal instruction): the result is re-packed to bf16.
#pragma once
#include <cuda_fp16.h>

#include "attn_tc.cuh"

namespace dgvit {
namespace mlp {

using namespace tc;
using attn::fence_async_smem;
using attn::sw128_off;
using attn::tmem_ld16;

constexpr int HC = 128;            // hidden chunk (columns of GEMM1 / K of GEMM2)
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int TILE16 = 16384;      // [128 rows][64 x 16-bit] swizzled tile
constexpr int MAX_HID = 4096;      // bias copy in smem

constexpr float GELU_A = 0.80015708f, GELU_B = 0.03470089f;

// kind::f16 instruction descriptor with explicit operand formats (0 = f16, 1 = bf16), D = f32
__host__ __device__ constexpr uint32_t make_idesc_fmt(int M, int N, bool a_mn, bool b_mn, uint32_t afmt, uint32_t bfmt) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ __half2 h2_tanh(__half2 x) {
  uint32_t r;
  const uint32_t a = *reinterpret_cast<const uint32_t*>(&x);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ uint32_t h2_to_bf2_bits(__half2 x) {
  const float2 f = __half22float2(x);
  const __nv_bfloat162 p = __floats2bfloat162_rn(f.x, f.y);
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ __half2 gelu_h2(__half2 x) {
  const __half2 A = __float2half2_rn(GELU_A), B = __float2half2_rn(GELU_B), hf = __float2half2_rn(0.5f);
  const __half2 x2 = __hmul2(x, x);
  const __half2 t = h2_tanh(__hmul2(x, __hfma2(x2, B, A)));
  const __half2 hx = __hmul2(x, hf);
  return __hfma2(hx, t, hx);
}
// the same function of xh = x / 2 (the forward epilogue forms xh = 0.5 acc + 0.5 b1 with ONE HFMA2, which also replaces
// the separate 0.5 x product): gelu(x) = xh + xh tanh(xh (2a + 8b xh^2))  — 4 FMA-pipe instructions + 1 MUFU per pair
__device__ __forceinline__ __half2 gelu_half_h2(__half2 xh) {
  const __half2 A2 = __float2half2_rn(2.0f * GELU_A), B8 = __float2half2_rn(8.0f * GELU_B);
  const __half2 s = __hmul2(xh, xh);
  const __half2 t = h2_tanh(__hmul2(xh, __hfma2(s, B8, A2)));
  return __hfma2(xh, t, xh);
}
// g = gelu(x), d = d gelu / dx of the same approximant
__device__ __forceinline__ void gelu_grad_h2(__half2 x, __half2& g, __half2& d) {
  const __half2 A = __float2half2_rn(GELU_A), B = __float2half2_rn(GELU_B), hf = __float2half2_rn(0.5f);
  const __half2 B3 = __float2half2_rn(3.0f * GELU_B), one = __float2half2_rn(1.0f);
  const __half2 x2 = __hmin2(__hmul2(x, x), __float2half2_rn(60000.0f));   // keeps (1 - t^2) * q finite
  const __half2 t = h2_tanh(__hmul2(x, __hfma2(x2, B, A)));
  const __half2 hx = __hmul2(x, hf);
  g = __hfma2(hx, t, hx);
  const __half2 q = __hfma2(x2, B3, A);
  const __half2 s = __hfma2(__hneg2(t), t, one);
  d = __hfma2(__hmul2(hx, s), q, __hfma2(t, hf, hf));
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&p);
}

// 32 lanes x 32 columns, no wait (pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int NWARPS = EPI_WARPS>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NWARPS * 32) : "memory"); }

// fp32 bias -> f16 copy in shared memory (epilogue threads only), then a barrier among them
template <int NWARPS = EPI_WARPS>
__device__ __forceinline__ void stage_bias(const float* b, __half* dst, int n, int tid, float scale = 1.0f) {
  for (int i = tid * 2; i < n; i += NWARPS * 32 * 2)
    *reinterpret_cast<__half2*>(dst + i) = __floats2half2_rn(scale * __ldg(b + i), scale * __ldg(b + i + 1));
  epi_bar_sync<NWARPS>();
}

// thread-block-cluster helpers (split-hidden forward)
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, int cta_rank) {
  uint32_t r;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t addr) {
  float2 v;
  asm("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  // not volatile / no memory clobber: the caller orders it after cluster_sync(), and independent loads must overlap
  // (a distributed-shared-memory load is ~750 cycles)
  asm("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// =====================================================================================
// forward
// =====================================================================================
namespace f {
constexpr int NSPLIT = 8;          // CTAs per token tile in the split-hidden variant (one cluster)
constexpr int NST = 3;             // weight ring depth
constexpr int NG = 2;              // epilogue groups = accumulator / H buffers in flight (3 measured slower: 30.7 vs 28.7 us)
constexpr int GW = 8;              // warps per group (2 per TMEM lane quadrant, 64 chunk columns each)
constexpr int FWD_EPI_WARPS = NG * GW;
constexpr int FWD_THREADS = 64 + FWD_EPI_WARPS * 32;
constexpr int X_BYTES = TILE16, W_BYTES = TILE16, H_BYTES = 2 * TILE16;
constexpr int OFF_W1 = X_BYTES;
constexpr int OFF_W2 = OFF_W1 + NST * W_BYTES;
constexpr int OFF_H = OFF_W2 + NST * W_BYTES;
constexpr int OFF_BIAS = OFF_H + NG * H_BYTES;
constexpr int OFF_BAR = OFF_BIAS + MAX_HID * 2;
constexpr int OFF_LN = OFF_BAR + 256;          // [128 rows][4 column groups] partial sums of the fused LayerNorm
constexpr int OFF_WO = (OFF_LN + 2048 + 1023) / 1024 * 1024;   // FRONT: W_out [64][256] bf16 (4 k-blocks of 8 KB), then the x_m tile [128][64] fp32
constexpr int SMEM_TOTAL = OFF_LN + 2048 + 1024;
constexpr int SMEM_TOTAL_FRONT = OFF_WO + 32768 + 1024;
static_assert(SMEM_TOTAL_FRONT <= 232448 && OFF_WO % 1024 == 0, "smem budget");
static_assert(SMEM_TOTAL <= 232448, "smem budget");
}  // namespace f

// optional timeline instrumentation (DGVIT_MLP_TRACE builds only): CTA 0 records clock64() at pipeline events
#ifdef DGVIT_MLP_TRACE
__device__ __forceinline__ long long gtime_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MLP_GT(idx) do { if (a.trace && threadIdx.x == 64) a.trace[2048 + (blockIdx.x & 255) * 4 + (idx)] = gtime_ns(); } while (0)
#define MLP_TRACE(slot, idx) do { if (a.trace && blockIdx.x == 0 && (lane == 0 || warp < 2)) a.trace[(slot) * 64 + (idx)] = clock64(); } while (0)
#else
#define MLP_TRACE(slot, idx) do { } while (0)
#define MLP_GT(idx) do { } while (0)
#endif

struct MlpArgs {
  long long* trace;           // [16 slots][64] (null unless tracing)
  int M, HID;                 // token rows, hidden width
  const float* b1; const float* b2;
  const float* resid; int64_t ldr;
  float* out; int64_t ldc;
  // optional fused LayerNorm of the output rows (the next block's PreNorm, vn/GoalFormer.py:31-37,103):
  const float* ln_gamma; const float* ln_beta;   // null = off
  bf16* ln_out; float* ln_mean; float* ln_rstd;  // [M,64] bf16, [M], [M]
  // FRONT: the attention out-projection + residual + this block's LayerNorm-2 run as the kernel's prologue
  // (x = to_out(o) + x ; ff input = LN(x), vn/GoalFormer.py:81-82,103-104): the kernel reads the attention output o
  // (tmO) instead of a ready-made LN(x) tile and also emits x_m, LN(x_m), mean, rstd for the backward.
  const float* ob;                               // to_out.0.bias [64]
  const float* xa; int64_t ldxa;                 // residual input rows (fp32)
  float* xm;                                     // [M,64] fp32 out (= resid of the MLP)
  const float* ln2_gamma; const float* ln2_beta;
  bf16* xn2; float* mean2; float* rstd2;
};

// SPLIT: few token tiles (the pruned last block, batch-1 act): a cluster of NSPLIT CTAs shares one tile, each taking
// HID / NSPLIT hidden columns; the partial outputs meet in CTA 0 of the cluster through distributed shared memory.
// FRONT: see MlpArgs.  tmX then maps the attention output o ([M rows, 256] bf16) and tmWo the out-projection weight.
// H16: the hidden activation tile and W2 are f16 (tmW2 maps the f16 copy of net.3.weight): the GELU result leaves the
// epilogue as it is computed (f16x2) instead of being re-packed to bf16 (2 cvt + 1 F2FP per pair); kind::f16 takes either
// format but not a mix, so GEMM2 runs f16 x f16 while GEMM1 and the prologue stay bf16 x bf16.
template <bool SPLIT, bool FRONT, bool H16>
__global__ void __launch_bounds__(f::FWD_THREADS, 1)
mlp_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmWo, const MlpArgs a) {
  using namespace f;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* x_full = bars;                 // 1
  uint64_t* w1_full = x_full + 1;          // NST
  uint64_t* w1_empty = w1_full + NST;      // NST
  uint64_t* w2_full = w1_empty + NST;      // NST
  uint64_t* w2_empty = w2_full + NST;      // NST
  uint64_t* acc_full = w2_empty + NST;     // NG
  uint64_t* acc_free = acc_full + NG;      // NG (8 warps each)
  uint64_t* h_ready = acc_free + NG;       // NG (8 warps each)
  uint64_t* h_free = h_ready + NG;         // NG
  uint64_t* y_full = h_free + NG;          // 1
  uint64_t* front_done = y_full + 1;       // 1: out-projection accumulator complete (FRONT)
  uint64_t* x_ready = front_done + 1;      // 1 (16 warps): LN(x_m) tile written to smem (FRONT)
  uint32_t* tmem_slot = (uint32_t*)(x_ready + 1);
  __half* bias_s = (__half*)(smem + OFF_BIAS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = SPLIT ? blockIdx.x / NSPLIT : blockIdx.x;
  const int rank = SPLIT ? blockIdx.x % NSPLIT : 0;                 // == %cluster_ctarank (cluster = NSPLIT CTAs along x)
  const int NC = SPLIT ? a.HID / HC / NSPLIT : a.HID / HC;          // chunks of this CTA
  const int cbase = rank * NC;                                      // first hidden chunk of this CTA
  if (warp == 2) MLP_TRACE(12, 0);
  MLP_GT(0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    mbar_init(x_full, 1);
    for (int i = 0; i < NST; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1); }
    for (int i = 0; i < NG; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], GW); mbar_init(&h_ready[i], GW); mbar_init(&h_free[i], 1); }
    mbar_init(y_full, 1);
    mbar_init(front_done, 1); mbar_init(x_ready, FWD_EPI_WARPS);
    if (FRONT) tma_prefetch_desc(&tmWo);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;     // accumulators: cols [0, NG*128) ; Y: cols [NG*128, NG*128+64)
  constexpr uint32_t T_Y = NG * HC;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail
  if (warp == 2) MLP_TRACE(12, 1);
  pdl_wait();
  pdl_launch();
  if (warp == 2) MLP_TRACE(12, 2);
  MLP_GT(1);

  if (warp == 0) {
    if (elect_one_sync()) {
      if (FRONT) {      // o tile: 4 k-blocks of [128 rows][64] into the (still unused) H buffers; W_out: 4 k-blocks of [64][64]
        mbar_expect_tx(x_full, 4 * TILE16 + 4 * 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tma_load_2d(smem + OFF_H + k * TILE16, &tmX, x_full, k * 64, mt * 128);
          tma_load_2d(smem + OFF_WO + k * 8192, &tmWo, x_full, k * 64, 0);
        }
      } else {
        mbar_expect_tx(x_full, X_BYTES);
        tma_load_2d(smem, &tmX, x_full, 0, mt * 128);
      }
      for (int c = 0; c < NC; ++c) {
        const int s = c % NST; const uint32_t ph = (c / NST) & 1;
        mbar_wait(&w1_empty[s], ph ^ 1);
        mbar_expect_tx(&w1_full[s], W_BYTES);
        tma_load_2d(smem + OFF_W1 + s * W_BYTES, &tmW1, &w1_full[s], 0, (cbase + c) * HC);          // [128 hidden rows][64 k]
        mbar_wait(&w2_empty[s], ph ^ 1);
        mbar_expect_tx(&w2_full[s], W_BYTES);
        tma_load_2d(smem + OFF_W2 + s * W_BYTES, &tmW2, &w2_full[s], (cbase + c) * HC, 0);          // [64 d rows][64 k] k-block 0
        tma_load_2d(smem + OFF_W2 + s * W_BYTES + 8192, &tmW2, &w2_full[s], (cbase + c) * HC + 64, 0);  // k-block 1
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc1 = make_idesc(128, HC, false, false);
      constexpr uint32_t idesc_front = make_idesc(128, 64, false, false);
      constexpr uint32_t idesc2 = make_idesc_fmt(128, 64, false, false, H16 ? 0u : 1u, H16 ? 0u : 1u);
      const uint32_t sx = smem_u32(smem);
      auto gemm1 = [&](int c) {
        const int s = c % NST; const uint32_t ph = (c / NST) & 1;
        const int ab = c % NG; const uint32_t aph = (c / NG) & 1;
        mbar_wait(&acc_free[ab], aph ^ 1);
        mbar_wait(&w1_full[s], ph);
        tc_fence_after();
        const uint64_t xd = make_smem_desc(sx, 16, 1024);
        const uint64_t wd = make_smem_desc(sx + OFF_W1 + s * W_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + ab * HC, xd + (uint64_t)(k * 2), wd + (uint64_t)(k * 2), idesc1, k > 0);
        umma_commit(&w1_empty[s]);
        umma_commit(&acc_full[ab]);
      };
      mbar_wait(x_full, 0);
      MLP_TRACE(12, 8);
      if (FRONT) {      // x_m accumulator = o W_out^T  (K = 256) into the Y columns, then wait for the epilogue's LN(x_m) tile
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint64_t od = make_smem_desc(sx + OFF_H + (k >> 2) * TILE16 + (k & 3) * 32, 16, 1024);
          const uint64_t wd = make_smem_desc(sx + OFF_WO + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024);
          umma_bf16(tmem_base + T_Y, od, wd, idesc_front, k > 0);
        }
        umma_commit(front_done);
        mbar_wait(x_ready, 0);
        tc_fence_after();
      }
      gemm1(0);
      MLP_TRACE(12, 9);
      for (int c = 1; c < NG && c < NC; ++c) gemm1(c);
      for (int c = 0; c < NC; ++c) {
        // accumulator buffer c % NG is free as soon as the epilogue of chunk c has pulled it into registers, long before
        // that epilogue finishes: GEMM1(c+NG) is issued now so its result is waiting when the group comes back
        if (c + NG < NC) gemm1(c + NG);
        const int s = c % NST; const uint32_t ph = (c / NST) & 1;
        const int hb = c % NG; const uint32_t hph = (c / NG) & 1;
        MLP_TRACE(8, c);
        mbar_wait(&h_ready[hb], hph);
        MLP_TRACE(9, c);
        mbar_wait(&w2_full[s], ph);
        MLP_TRACE(10, c);
        tc_fence_after();
        const uint32_t sh = sx + OFF_H + hb * H_BYTES, sw = sx + OFF_W2 + s * W_BYTES;
#pragma unroll
        for (int k = 0; k < HC / 16; ++k) {
          const uint64_t hd = make_smem_desc(sh + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          const uint64_t wd = make_smem_desc(sw + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024);
          umma_bf16(tmem_base + T_Y, hd, wd, idesc2, (c > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&w2_empty[s]);
        umma_commit(&h_free[hb]);
        MLP_TRACE(11, c);
      }
      umma_commit(y_full);
    }
  } else {
    // NG groups of 8 warps: group g owns the chunks c = g (mod NG), i.e. accumulator buffer g and H buffer g, so one
    // group's TMEM-load / smem-store / barrier phases overlap the other groups' MUFU + FMA phases (measured: a third group does not help, the MUFU / FMA / issue mix is the limit, not latency).
    const int ew = warp - 2;
    const int pg = ew / GW, half = (ew >> 2) & 1;        // group ; 64-column half of the chunk
    const int quad = warp & 3, grp = ew >> 2;            // TMEM lane quadrant ; 16-column group of the final output (grp < 4)
    const int r = quad * 32 + lane;
    const int row0 = mt * 128 + quad * 32;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    stage_bias<FWD_EPI_WARPS>(a.b1, bias_s, a.HID, threadIdx.x - 64, 0.5f);      // 0.5 b1: see gelu_half_h2
    if (warp == 2) MLP_TRACE(12, 3);
    if (FRONT) {
      static_assert(!FRONT || FWD_EPI_WARPS == 16, "front epilogue: 4 lane quadrants x 4 column groups");
      // x_m = o W_out^T + b_out + x_a (fp32) ; LayerNorm-2 -> bf16 tile that is the A operand of every GEMM1.
      // x_m also stays in shared memory (over the dead W_out tile) as the residual of the output stage.
      // residual row and bias are fetched BEFORE the wait for the out-projection accumulator: their global-load latency
      // runs under the TMA of the attention-output tile and the 16 MMAs instead of after them
      const int row = row0 + lane;
      const bool ok = row < a.M;
      float4 rr4[4], bb4[4];
      if (ok) {
        const float* R = a.xa + (int64_t)row * a.ldxa + grp * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) { rr4[i] = reinterpret_cast<const float4*>(R)[i]; bb4[i] = __ldg(reinterpret_cast<const float4*>(a.ob + grp * 16) + i); }
      }
      mbar_wait(front_done, 0);
      tc_fence_after();
      float y[16];
      tmem_ld16(tmem_base + T_Y + grp * 16 + lane_off, y);
      if (ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 r4 = rr4[i], b4 = bb4[i];
          y[4 * i] += b4.x + r4.x; y[4 * i + 1] += b4.y + r4.y; y[4 * i + 2] += b4.z + r4.z; y[4 * i + 3] += b4.w + r4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = 0.f;
      }
      float4* xs = reinterpret_cast<float4*>(smem + OFF_WO + (r * 64 + grp * 16) * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) xs[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
      const bool wr = ok && rank == 0;            // (SPLIT: every CTA of the cluster computes the same rows)
      if (wr) {
        float4* O = reinterpret_cast<float4*>(a.xm + (int64_t)row * 64 + grp * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) O[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
      }
      float* part = reinterpret_cast<float*>(smem + OFF_LN);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) sum += y[i];
      part[r * 4 + grp] = sum;
      epi_bar_sync<16>();
      const float4 p4 = *reinterpret_cast<const float4*>(part + r * 4);
      const float mu = ((p4.x + p4.y) + (p4.z + p4.w)) * (1.0f / 64.0f);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float c = y[i] - mu; q = fmaf(c, c, q); }
      epi_bar_sync<16>();
      part[r * 4 + grp] = q;
      epi_bar_sync<16>();
      const float4 q4 = *reinterpret_cast<const float4*>(part + r * 4);
      const float rs = 1.0f / sqrtf(((q4.x + q4.y) + (q4.z + q4.w)) * (1.0f / 64.0f) + 1e-5f);
      epi_bar_sync<16>();                               // table free again (the output stage reuses it)
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = grp * 16 + 2 * i;
        const float2 g2 = __ldg(reinterpret_cast<const float2*>(a.ln2_gamma + c)), b2 = __ldg(reinterpret_cast<const float2*>(a.ln2_beta + c));
        w[i] = ok ? pack_bf2((y[2 * i] - mu) * rs * g2.x + b2.x, (y[2 * i + 1] - mu) * rs * g2.y + b2.y) : 0u;
      }
      *reinterpret_cast<uint4*>(smem + sw128_off(r, grp * 2)) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(smem + sw128_off(r, grp * 2 + 1)) = make_uint4(w[4], w[5], w[6], w[7]);
      if (wr) {
        uint4* dst = reinterpret_cast<uint4*>(a.xn2 + (int64_t)row * 64 + grp * 16);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        if (grp == 0 && a.mean2) { a.mean2[row] = mu; a.rstd2[row] = rs; }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_ready);
    }
    for (int c = pg; c < NC; c += NG) {
      const int ab = pg; const uint32_t aph = (c / NG) & 1;
      if ((ew & 7) == 0) MLP_TRACE(0, c);
      mbar_wait(&acc_full[ab], aph);
      if ((ew & 7) == 0) MLP_TRACE(1, c);
      tc_fence_after();
      uint8_t* hb = smem + OFF_H + ab * H_BYTES + half * 16384;
      // the second 32 columns are requested while the first 32 are being turned into GELU values: one TMEM round trip per
      // chunk on the critical path instead of two
      uint32_t vv[2][32];
      tmem_ld32_nowait(tmem_base + ab * HC + half * 64 + lane_off, vv[0]);
      tmem_wait_ld();
      tmem_ld32_nowait(tmem_base + ab * HC + half * 64 + 32 + lane_off, vv[1]);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t (&v)[32] = vv[hh];
        if (hh == 1) {
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_free[ab]);      // TMEM chunk drained: GEMM1(c+2) may start
        }
        const uint4* bsm = reinterpret_cast<const uint4*>(bias_s + (cbase + c) * HC + half * 64 + hh * 32);
        uint4 ov[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 b4 = bsm[i];
          const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
          uint32_t ow[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __half2 acc = __floats2half2_rn(__uint_as_float(v[8 * i + 2 * j]), __uint_as_float(v[8 * i + 2 * j + 1]));
            const __half2 xh = __hfma2(acc, __float2half2_rn(0.5f), *reinterpret_cast<const __half2*>(&bw[j]));
            const __half2 g = gelu_half_h2(xh);
            ow[j] = H16 ? *reinterpret_cast<const uint32_t*>(&g) : h2_to_bf2_bits(g);
          }
          ov[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        // GEMM2(c-2) must have consumed this H buffer before the first store.  Its MMAs are issued only after this
        // group's previous chunk is complete, so they retire ~500 cycles into this chunk: the wait sits behind the first
        // half's arithmetic (results parked in registers) instead of in front of it (ncu: 10 % of all stall samples there).
        if (hh == 0) mbar_wait(&h_free[ab], aph ^ 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(hb + sw128_off(r, hh * 4 + i)) = ov[i];
      }
      if ((ew & 7) == 0) MLP_TRACE(2, c);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_ready[ab]);
      if ((ew & 7) == 0) MLP_TRACE(4, c);
    }
    if (SPLIT) {   // partial Y of this CTA's hidden columns -> own shared memory (the H buffers are dead by now)
      if (grp < 4) {
        mbar_wait(y_full, 0);
        tc_fence_after();
        float y[16];
        tmem_ld16(tmem_base + T_Y + grp * 16 + lane_off, y);
        float4* yb = reinterpret_cast<float4*>(smem + OFF_H + (r * 64 + grp * 16) * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) yb[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
      }
    }
  }
  if (warp == 2) MLP_TRACE(12, 10);
  if (SPLIT) {
    // Reduce-scatter over distributed shared memory (a single CTA pulling all partials would be bound by the
    // ~20 B/clk DSMEM port): CTA `rank` finishes rows [16 rank, 16 rank + 16) of the tile, one row per epilogue warp,
    // two columns per lane, partials summed in rank order (deterministic); LayerNorm statistics by warp shuffle.
    static_assert(FWD_EPI_WARPS * NSPLIT == 128, "one tile row per epilogue warp and cluster rank");
    cluster_sync();                     // every CTA's partial is in place (all threads of all CTAs)
    if (warp == 2) MLP_TRACE(12, 11);
    if (warp >= 2) {
      const int rt = rank * FWD_EPI_WARPS + (warp - 2);
      const int row = mt * 128 + rt, col = lane * 2;
      const uint32_t local = smem_u32(smem + OFF_H + (rt * 64 + col) * 4);
      float2 v[NSPLIT];
#pragma unroll
      for (int cr = 0; cr < NSPLIT; ++cr) v[cr] = ld_dsmem_f2(mapa_shared(local, cr));
      float y0 = v[0].x, y1 = v[0].y;
#pragma unroll
      for (int cr = 1; cr < NSPLIT; ++cr) { y0 += v[cr].x; y1 += v[cr].y; }
      if (warp == 2) MLP_TRACE(12, 12);
      const bool ok = row < a.M;
      if (ok) {
        const float2 b2 = __ldg(reinterpret_cast<const float2*>(a.b2 + col));
        const float2 rr = FRONT ? *reinterpret_cast<const float2*>(smem + OFF_WO + (rt * 64 + col) * 4)
                                : *reinterpret_cast<const float2*>(a.resid + (int64_t)row * a.ldr + col);
        y0 += b2.x + rr.x; y1 += b2.y + rr.y;
        *reinterpret_cast<float2*>(a.out + (int64_t)row * a.ldc + col) = make_float2(y0, y1);
      }
      if (a.ln_gamma) {
        const float mu = warp_sum(y0 + y1) * (1.0f / 64.0f);
        const float c0 = y0 - mu, c1 = y1 - mu;
        const float rs = 1.0f / sqrtf(warp_sum(fmaf(c0, c0, c1 * c1)) * (1.0f / 64.0f) + 1e-5f);
        if (ok) {
          const float2 g2 = __ldg(reinterpret_cast<const float2*>(a.ln_gamma + col)), be = __ldg(reinterpret_cast<const float2*>(a.ln_beta + col));
          *reinterpret_cast<uint32_t*>(a.ln_out + (int64_t)row * 64 + col) = pack_bf2(c0 * rs * g2.x + be.x, c1 * rs * g2.y + be.y);
          if (lane == 0 && a.ln_mean) { a.ln_mean[row] = mu; a.ln_rstd[row] = rs; }
        }
      }
    }
  } else if (warp >= 2) {
    const int ew = warp - 2;
    const int quad = warp & 3, grp = ew >> 2;
    const int r = quad * 32 + lane;
    const int row0 = mt * 128 + quad * 32;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    // ---- output: y + b2 + residual  (16 columns per warp, the first 16 epilogue warps); the residual loads overlap
    //      the last GEMM2
    if (grp < 4) {
    const int row = row0 + lane;
    float4 r4[4], b4[4];
    if (row < a.M) {
      const float* R = FRONT ? reinterpret_cast<const float*>(smem + OFF_WO) + r * 64 + grp * 16
                             : a.resid + (int64_t)row * a.ldr + grp * 16;
#pragma unroll
      for (int i = 0; i < 4; ++i) { r4[i] = reinterpret_cast<const float4*>(R)[i]; b4[i] = __ldg(reinterpret_cast<const float4*>(a.b2 + grp * 16) + i); }
    }
    float y[16];
    {
      if (warp == 2) MLP_TRACE(12, 4);
      mbar_wait(y_full, 0);
      if (warp == 2) MLP_TRACE(12, 5);
      tc_fence_after();
      tmem_ld16(tmem_base + T_Y + grp * 16 + lane_off, y);
    }
    if (row < a.M) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        y[4 * i] += b4[i].x + r4[i].x; y[4 * i + 1] += b4[i].y + r4[i].y;
        y[4 * i + 2] += b4[i].z + r4[i].z; y[4 * i + 3] += b4[i].w + r4[i].w;
      }
      float* O = a.out + (int64_t)row * a.ldc + grp * 16;
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(O)[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
    }
    if (a.ln_gamma) {
      // LayerNorm over the 64 columns of a row = this thread's 16 + three other warps' (same quadrant, other groups):
      // two-pass statistics through a [128][4] table in shared memory
      float* part = reinterpret_cast<float*>(smem + OFF_LN);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += y[i];
      part[r * 4 + grp] = s;
      epi_bar_sync<16>();
      const float4 p4 = *reinterpret_cast<const float4*>(part + r * 4);
      const float mu = ((p4.x + p4.y) + (p4.z + p4.w)) * (1.0f / 64.0f);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float c = y[i] - mu; q = fmaf(c, c, q); }
      epi_bar_sync<16>();                               // everyone has read the sums
      part[r * 4 + grp] = q;
      epi_bar_sync<16>();
      const float4 q4 = *reinterpret_cast<const float4*>(part + r * 4);
      const float rs = 1.0f / sqrtf(((q4.x + q4.y) + (q4.z + q4.w)) * (1.0f / 64.0f) + 1e-5f);
      if (row < a.M) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = grp * 16 + 2 * i;
          const float2 g2 = __ldg(reinterpret_cast<const float2*>(a.ln_gamma + c)), b2 = __ldg(reinterpret_cast<const float2*>(a.ln_beta + c));
          w[i] = pack_bf2((y[2 * i] - mu) * rs * g2.x + b2.x, (y[2 * i + 1] - mu) * rs * g2.y + b2.y);
        }
        uint4* dst = reinterpret_cast<uint4*>(a.ln_out + (int64_t)row * 64 + grp * 16);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        if (grp == 0 && a.ln_mean) { a.ln_mean[row] = mu; a.ln_rstd[row] = rs; }
      }
    }
    }
  }
  if (warp == 2) MLP_TRACE(12, 6);
  if (SPLIT) cluster_sync();            // every CTA has read its slice of everyone's partial: shared memory may go away
  tc_fence_before();
  __syncthreads();
  if (warp == 2) MLP_TRACE(12, 7);
  MLP_GT(2);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================
// backward
// =====================================================================================
namespace b {
constexpr int NST = 3;                       // ring depth of the streamed operand pair
constexpr int OFF_FIX = 0;                   // fixed pair:    MODE 0: X | dY      MODE 1: W1c | W2c
constexpr int OFF_RING = 2 * TILE16;         // streamed pair: MODE 0: W1c | W2c   MODE 1: X | dY
// 4 KB of bf16 1.0 (operand of the column-sum MMAs).  As a 128-row A operand only its first 8-row
// group is ones; the other rows alias the dH tile that follows and produce accumulator rows nobody reads.
constexpr int OFF_ONES = OFF_RING + NST * 2 * TILE16;
constexpr int OFF_DH = OFF_ONES + 4096;               // dH tile [128 tok][128 hid] bf16 (two 64-column blocks)
constexpr int OFF_G = OFF_DH + 2 * TILE16;            // G tile (MODE 1)
constexpr int OFF_BIAS = OFF_G + 2 * TILE16;
constexpr int OFF_BAR = OFF_BIAS + MAX_HID * 2;
constexpr int SMEM_TOTAL = OFF_BAR + 256 + 1024;
static_assert(SMEM_TOTAL <= 232448, "smem budget");
static_assert(OFF_DH % 1024 == 0 && OFF_ONES % 1024 == 0, "swizzle atoms are 1024-byte aligned");
// TMEM columns
constexpr uint32_t T_HPRE = 0, T_DG = 128, T_ACC0 = 256, T_ACC1 = 320, T_ACC2 = 384;
}  // namespace b

struct MlpBwdArgs {
  int M, HID;                 // token rows, hidden width
  int tiles_per_split;        // MODE 1: token tiles per CTA
  const float* b1;
  float* dX; int64_t lddx;    // MODE 0: [M, 64] fp32
  float* db2_part;            // MODE 0: [tiles][64] column sums of dY (or null)
  float* dW1_part;            // MODE 1: [split][HID][64]
  float* db1_part;            // MODE 1: [split][HID]
  float* dW2_part;            // MODE 1: [split][64][HID]
  int64_t split_stride;       // floats between consecutive splits (same for the three partials)
};

// SPLIT (MODE 0 only, few token tiles): a cluster of f::NSPLIT CTAs shares one tile, each taking HID / NSPLIT hidden
// columns; the partial dX tiles are reduce-scattered over distributed shared memory as in the forward.
template <int MODE, bool SPLIT = false>
__global__ void __launch_bounds__(THREADS, 1)
mlp_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const MlpBwdArgs a) {
  using namespace b;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* fix_full = bars;                // 1
  uint64_t* ring_full = fix_full + 1;       // NST
  uint64_t* ring_empty = ring_full + NST;   // NST
  uint64_t* acc_full = ring_empty + NST;    // 1: Hpre and dG of this step are in TMEM
  uint64_t* acc_free = acc_full + 1;        // 1 (16 warps): both drained into registers
  uint64_t* h_ready = acc_free + 1;         // 1 (16 warps): dH (and G) tiles written
  uint64_t* h_free = h_ready + 1;           // 1: the MMAs reading the tiles have retired
  uint64_t* out_full = h_free + 1;          // 1
  uint32_t* tmem_slot = (uint32_t*)(out_full + 1);
  __half* bias_s = (__half*)(smem + OFF_BIAS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NC = a.HID / HC;
  const int tiles = (a.M + 127) / 128;
  // step i of this CTA = (token tile t_of(i), hidden chunk c_of(i))
  int nsteps, t0, c0;
  const int rank = SPLIT ? blockIdx.x % f::NSPLIT : 0;
  if (MODE == 0 && SPLIT) { t0 = blockIdx.x / f::NSPLIT; nsteps = NC / f::NSPLIT; c0 = rank * nsteps; }
  else if (MODE == 0) { t0 = blockIdx.x; c0 = 0; nsteps = NC; }
  else {
    c0 = blockIdx.x % NC;
    t0 = (blockIdx.x / NC) * a.tiles_per_split;
    nsteps = min(a.tiles_per_split, tiles - t0);
  }
  const int split = (MODE == 0) ? 0 : blockIdx.x / NC;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    mbar_init(fix_full, 1);
    for (int i = 0; i < NST; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_free, EPI_WARPS); mbar_init(h_ready, EPI_WARPS); mbar_init(h_free, 1);
    mbar_init(out_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // constant ones tile (generic-proxy writes, made visible to the tensor core by the fence below)
  for (int i = threadIdx.x; i < 4096 / 4; i += THREADS) reinterpret_cast<uint32_t*>(smem + OFF_ONES)[i] = 0x3F803F80u;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  // smem images
  uint8_t* sX = smem + (MODE == 0 ? OFF_FIX : OFF_RING);
  uint8_t* sW = smem + (MODE == 0 ? OFF_RING : OFF_FIX);
  // (X, dY) and (W1c, W2c) are adjacent 16 KB tiles; the streamed pair advances by 2*TILE16 per stage

  if (warp == 0) {
    if (elect_one_sync()) {
      auto load_tok = [&](uint8_t* dst, uint64_t* bar, int t) {   // X tile | dY tile, rows t*128..
        tma_load_2d(dst, &tmX, bar, 0, t * 128);
        tma_load_2d(dst + TILE16, &tmDY, bar, 0, t * 128);
      };
      auto load_w = [&](uint8_t* dst, uint64_t* bar, int c) {     // W1[c*128.., :] | W2[:, c*128..] (two 64-col blocks)
        tma_load_2d(dst, &tmW1, bar, 0, c * HC);
        tma_load_2d(dst + TILE16, &tmW2, bar, c * HC, 0);
        tma_load_2d(dst + TILE16 + 8192, &tmW2, bar, c * HC + 64, 0);
      };
      mbar_expect_tx(fix_full, 2 * TILE16);
      if (MODE == 0) load_tok(sX, fix_full, t0); else load_w(sW, fix_full, c0);
      for (int i = 0; i < nsteps; ++i) {
        const int s = i % NST; const uint32_t ph = (i / NST) & 1;
        mbar_wait(&ring_empty[s], ph ^ 1);
        mbar_expect_tx(&ring_full[s], 2 * TILE16);
        if (MODE == 0) load_w(sW + s * 2 * TILE16, &ring_full[s], c0 + i);
        else load_tok(sX + s * 2 * TILE16, &ring_full[s], t0 + i);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t id_hpre = make_idesc(128, HC, false, false);                       // X (K) x W1c (K)
      constexpr uint32_t id_dg = make_idesc(128, HC, false, true);                          // dY (K) x W2c (MN)
      constexpr uint32_t id_dx = make_idesc(128, 64, false, true);                          // dH (K) x W1c (MN)
      constexpr uint32_t id_dw1 = make_idesc(128, 64, true, true);                          // dH^T (MN) x X (MN)
      constexpr uint32_t id_dw2 = make_idesc(128, 64, true, true);                          // G^T (MN) x dY (MN)
      constexpr uint32_t id_db1 = make_idesc(128, 16, true, false);                         // dH^T (MN) x ones
      constexpr uint32_t id_db2 = make_idesc(128, 64, false, true);                         // ones (K) x dY (MN)
      const uint32_t s_dh = smem_u32(smem + OFF_DH), s_g = smem_u32(smem + OFF_G), s_one = smem_u32(smem + OFF_ONES);
      auto stage_x = [&](int i) { return smem_u32(sX) + (MODE == 0 ? 0 : (i % NST) * 2 * TILE16); };
      auto stage_w = [&](int i) { return smem_u32(sW) + (MODE == 0 ? (i % NST) * 2 * TILE16 : 0); };
      auto issue_ab = [&](int i) {
        const int s = i % NST; const uint32_t ph = (i / NST) & 1;
        if (i > 0) mbar_wait(acc_free, (uint32_t)((i - 1) & 1));
        mbar_wait(&ring_full[s], ph);
        tc_fence_after();
        const uint32_t x = stage_x(i), dy = x + TILE16, w1 = stage_w(i), w2 = w1 + TILE16;
        const uint64_t xd = make_smem_desc(x, 16, 1024), w1d = make_smem_desc(w1, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + T_HPRE, xd + (uint64_t)(k * 2), w1d + (uint64_t)(k * 2), id_hpre, k > 0);
        const uint64_t dyd = make_smem_desc(dy, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // W2c image: rows = 64 out (K), two 64-hid blocks 8192 B apart
          umma_bf16(tmem_base + T_DG, dyd + (uint64_t)(k * 2), make_smem_desc(w2 + k * 2048, 8192, 1024), id_dg, k > 0);
        umma_commit(acc_full);
      };
      mbar_wait(fix_full, 0);
      issue_ab(0);
      for (int i = 0; i < nsteps; ++i) {
        if (i + 1 < nsteps) issue_ab(i + 1);
        const int s = i % NST;
        mbar_wait(h_ready, (uint32_t)(i & 1));
        tc_fence_after();
        const uint32_t x = stage_x(i), dy = x + TILE16, w1 = stage_w(i);
        if (MODE == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dX[tok, in] += dH[tok, hid] W1c[hid, in]
            umma_bf16(tmem_base + T_ACC0, make_smem_desc(s_dh + (k >> 2) * TILE16 + (k & 3) * 32, 16, 1024),
                      make_smem_desc(w1 + k * 2048, TILE16, 1024), id_dx, (i > 0 || k > 0) ? 1u : 0u);
          if (i == 0 && a.db2_part && rank == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k)   // every row of the result = column sums of the dY tile
              umma_bf16(tmem_base + T_ACC1, make_smem_desc(s_one, 16, 1024), make_smem_desc(dy + k * 2048, TILE16, 1024),
                        id_db2, k > 0);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // K = 16 tokens per step = 16 rows of 128 B
            const uint64_t dhd = make_smem_desc(s_dh + k * 2048, TILE16, 1024);
            const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
            umma_bf16(tmem_base + T_ACC0, dhd, make_smem_desc(x + k * 2048, TILE16, 1024), id_dw1, acc);
            umma_bf16(tmem_base + T_ACC1, make_smem_desc(s_g + k * 2048, TILE16, 1024),
                      make_smem_desc(dy + k * 2048, TILE16, 1024), id_dw2, acc);
            umma_bf16(tmem_base + T_ACC2, dhd, make_smem_desc(s_one, 16, 1024), id_db1, acc);
          }
        }
        umma_commit(&ring_empty[s]);
        umma_commit(h_free);
      }
      umma_commit(out_full);
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3, grp = ew >> 2;            // rows quad*32.., hidden columns grp*32.. of the chunk
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    stage_bias(a.b1, bias_s, a.HID, threadIdx.x - 64);
    for (int i = 0; i < nsteps; ++i) {
      const int c = (MODE == 0) ? c0 + i : c0;
      mbar_wait(acc_full, (uint32_t)(i & 1));
      tc_fence_after();
      uint32_t hv[32], gv[32];
      tmem_ld32_nowait(tmem_base + T_HPRE + grp * 32 + lane_off, hv);
      tmem_ld32_nowait(tmem_base + T_DG + grp * 32 + lane_off, gv);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);               // TMEM drained: the next step's Hpre / dG MMAs may start
      const uint4* bsm = reinterpret_cast<const uint4*>(bias_s + c * HC + grp * 32);
      uint4 odh[4], og[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 b4 = bsm[q];
        const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
        uint32_t wd[4], wg[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = 8 * q + 2 * j;
          __half2 x = __floats2half2_rn(__uint_as_float(hv[e]), __uint_as_float(hv[e + 1]));
          x = __hadd2(x, *reinterpret_cast<const __half2*>(&bw[j]));
          __half2 g, d;
          gelu_grad_h2(x, g, d);
          const float2 df = __half22float2(d);       // dG stays fp32: gradients underflow f16
          wd[j] = pack_bf2(__uint_as_float(gv[e]) * df.x, __uint_as_float(gv[e + 1]) * df.y);
          wg[j] = h2_to_bf2_bits(g);
        }
        odh[q] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        og[q] = make_uint4(wg[0], wg[1], wg[2], wg[3]);
      }
      if (i > 0) mbar_wait(h_free, (uint32_t)((i - 1) & 1));   // the previous step's MMAs have consumed the tiles
      uint8_t* tdh = smem + OFF_DH + (grp >> 1) * TILE16;
      uint8_t* tg = smem + OFF_G + (grp >> 1) * TILE16;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = sw128_off(r, (grp & 1) * 4 + q);
        *reinterpret_cast<uint4*>(tdh + off) = odh[q];
        if (MODE == 1) *reinterpret_cast<uint4*>(tg + off) = og[q];
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready);
    }
    // ---- outputs (16 columns per warp)
    mbar_wait(out_full, 0);
    tc_fence_after();
    float y[16];
    if (MODE == 0) {
      tmem_ld16(tmem_base + T_ACC0 + grp * 16 + lane_off, y);
      const int row = t0 * 128 + r;
      if (SPLIT) {        // partial dX of this CTA's hidden columns -> own shared memory (the dH tile is dead by now)
        float4* yb = reinterpret_cast<float4*>(smem + OFF_DH + (r * 64 + grp * 16) * 4);
#pragma unroll
        for (int q = 0; q < 4; ++q) yb[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
      } else if (row < a.M) {
        float* O = a.dX + (int64_t)row * a.lddx + grp * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(O)[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
      }
      if (a.db2_part && quad == 0 && rank == 0) {         // row 0 of the ones x dY product
        tmem_ld16(tmem_base + T_ACC1 + grp * 16 + lane_off, y);
        if (lane == 0) {
          float* O = a.db2_part + (int64_t)t0 * 64 + grp * 16;
#pragma unroll
          for (int q = 0; q < 16; ++q) O[q] = y[q];
        }
      }
    } else {
      const int h = c0 * HC + r;                          // hidden unit of this thread's accumulator row
      const int64_t so = (int64_t)split * a.split_stride;
      tmem_ld16(tmem_base + T_ACC0 + grp * 16 + lane_off, y);
      {
        float* O = a.dW1_part + so + (int64_t)h * 64 + grp * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(O)[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
      }
      tmem_ld16(tmem_base + T_ACC1 + grp * 16 + lane_off, y);
#pragma unroll
      for (int q = 0; q < 16; ++q) a.dW2_part[so + (int64_t)(grp * 16 + q) * a.HID + h] = y[q];   // lanes = consecutive h
      if (grp == 0) {
        tmem_ld16(tmem_base + T_ACC2 + lane_off, y);
        a.db1_part[so + h] = y[0];
      }
    }
  }
  if (MODE == 0 && SPLIT) {
    static_assert(EPI_WARPS * f::NSPLIT == 128, "one tile row per epilogue warp and cluster rank");
    cluster_sync();                     // every CTA's partial dX is in place
    if (warp >= 2) {                    // rows [16 rank, 16 rank + 16): one per warp, two columns per lane, rank order
      const int rt = rank * EPI_WARPS + (warp - 2);
      const int row = t0 * 128 + rt, col = lane * 2;
      const uint32_t local = smem_u32(smem + OFF_DH + (rt * 64 + col) * 4);
      float2 v[f::NSPLIT];
#pragma unroll
      for (int cr = 0; cr < f::NSPLIT; ++cr) v[cr] = ld_dsmem_f2(mapa_shared(local, cr));
      float y0 = v[0].x, y1 = v[0].y;
#pragma unroll
      for (int cr = 1; cr < f::NSPLIT; ++cr) { y0 += v[cr].x; y1 += v[cr].y; }
      if (row < a.M) *reinterpret_cast<float2*>(a.dX + (int64_t)row * a.lddx + col) = make_float2(y0, y1);
    }
    cluster_sync();                     // everyone has read its slice: shared memory may go away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ host
static bool g_split_enabled = true;     // split-hidden cluster variant of the forward (set_option "mlp_split")
static long long* g_trace = nullptr;   // device buffer for DGVIT_MLP_TRACE builds (dgvit_set_option_ptr)
static bool eligible(int D, int HID, int64_t M, const void* x, const void* w1, const void* w2, const float* resid,
                     int64_t ldr, const float* out, int64_t ldc) {
  return tc::g_tc_enabled && D == 64 && HID % HC == 0 && HID <= MAX_HID && M >= 1 && ((uintptr_t)x & 15) == 0 &&
         ((uintptr_t)w1 & 15) == 0 && ((uintptr_t)w2 & 15) == 0 && ((uintptr_t)resid & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
         ldr % 4 == 0 && ldc % 4 == 0;
}

struct LnFuse {   // LayerNorm of the output rows, fused into the forward's last stage (all null = off)
  const float* gamma = nullptr; const float* beta = nullptr;
  bf16* out = nullptr; float* mean = nullptr; float* rstd = nullptr;
};
struct FrontFuse {   // out-projection + residual + LayerNorm-2 as the kernel's prologue (o == null: off)
  const bf16* o = nullptr; int64_t ldo = 0;       // attention output rows [M, 256] (row pitch ldo elements)
  const bf16* Wo = nullptr; const float* ob = nullptr;
  const float* xa = nullptr; int64_t ldxa = 0;
  float* xm = nullptr;
  const float* gamma = nullptr; const float* beta = nullptr;
  bf16* xn2 = nullptr; float* mean = nullptr; float* rstd = nullptr;
};
static bool g_front_enabled = true;     // set_option "mlp_front"
static bool front_eligible(int inner, const FrontFuse& fr) {
  return g_front_enabled && inner == 256 && fr.o && fr.Wo && fr.ob && fr.xa && fr.xm && fr.gamma && fr.beta && fr.xn2 &&
         ((((uintptr_t)fr.o) | ((uintptr_t)fr.Wo) | ((uintptr_t)fr.xa) | ((uintptr_t)fr.xm) | ((uintptr_t)fr.xn2)) & 15) == 0 &&
         fr.ldo % 8 == 0 && fr.ldxa % 4 == 0;
}
// W2h: f16 copy of W2 (null = bf16 W2 and a bf16 hidden tile)
static bool g_h16_enabled = true;       // set_option "mlp_h16"
static void fwd(const bf16* x, const bf16* W1, const float* b1, const bf16* W2, const float* b2, const float* resid,
                int64_t ldr, float* out, int64_t ldc, int64_t M, int HID, cudaStream_t st, const LnFuse& ln = LnFuse(),
                const FrontFuse& fr = FrontFuse(), const void* W2h = nullptr) {
  if (!g_h16_enabled || ((uintptr_t)W2h & 15)) W2h = nullptr;
  const bool h16 = W2h != nullptr;
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  a.trace = g_trace;
  a.ln_gamma = ln.gamma; a.ln_beta = ln.beta; a.ln_out = ln.out; a.ln_mean = ln.mean; a.ln_rstd = ln.rstd;
  DG_REQUIRE(!ln.gamma || (ln.beta && ln.out && (((uintptr_t)ln.out) & 15) == 0), "mlp::fwd: fused LayerNorm needs beta and an aligned output");
  a.M = (int)M; a.HID = HID; a.b1 = b1; a.b2 = b2; a.resid = resid; a.ldr = ldr; a.out = out; a.ldc = ldc;
  const bool front = fr.o != nullptr;
  a.ob = fr.ob; a.xa = fr.xa; a.ldxa = fr.ldxa; a.xm = fr.xm; a.ln2_gamma = fr.gamma; a.ln2_beta = fr.beta;
  a.xn2 = fr.xn2; a.mean2 = fr.mean; a.rstd2 = fr.rstd;
  CUtensorMap tx = front ? make_map(fr.o, 256, M, fr.ldo, 64, 128) : make_map(x, 64, M, 64, 64, 128);
  CUtensorMap tw1 = make_map(W1, 64, HID, 64, 64, 128);
  CUtensorMap tw2 = make_map(h16 ? W2h : (const void*)W2, HID, 64, HID, 64, 64);      // (16-bit elements either way)
  CUtensorMap two = front ? make_map(fr.Wo, 256, 64, 256, 64, 64) : tw1;
  static DevOnce attr;
  if (attr.first()) {
#define DG_MLP_ATTR(S_, F_, H_, B_) DG_CUDA(cudaFuncSetAttribute((mlp_fwd_tc_kernel<S_, F_, H_>), cudaFuncAttributeMaxDynamicSharedMemorySize, B_))
    DG_MLP_ATTR(false, false, false, f::SMEM_TOTAL); DG_MLP_ATTR(true, false, false, f::SMEM_TOTAL);
    DG_MLP_ATTR(false, true, false, f::SMEM_TOTAL_FRONT); DG_MLP_ATTR(true, true, false, f::SMEM_TOTAL_FRONT);
    DG_MLP_ATTR(false, false, true, f::SMEM_TOTAL); DG_MLP_ATTR(true, false, true, f::SMEM_TOTAL);
    DG_MLP_ATTR(false, true, true, f::SMEM_TOTAL_FRONT); DG_MLP_ATTR(true, true, true, f::SMEM_TOTAL_FRONT);
#undef DG_MLP_ATTR
  }
  if (skip_mask() & SKIP_MLP_FWD) return;
  const int grid = (int)cdiv(M, 128);
  // few tiles (pruned last block, batch-1 act): one 16-chunk CTA per tile is pure latency, so a cluster of 8 CTAs
  // shares each tile's hidden columns
  const bool split = g_split_enabled && grid * f::NSPLIT <= sm_count() && (HID / HC) % f::NSPLIT == 0 && (HID / HC) / f::NSPLIT >= 1;
  const int smem = front ? f::SMEM_TOTAL_FRONT : f::SMEM_TOTAL;
#define DG_MLP_GO(S_, F_, H_)                                                                                              \
  do {                                                                                                                     \
    if (S_) launch_k_cluster((mlp_fwd_tc_kernel<S_, F_, H_>), grid * f::NSPLIT, f::FWD_THREADS, smem, st, f::NSPLIT, tx, tw1, tw2, two, a); \
    else launch_k((mlp_fwd_tc_kernel<S_, F_, H_>), grid, f::FWD_THREADS, smem, st, tx, tw1, tw2, two, a);                    \
  } while (0)
  if (split) {
    if (front) { if (h16) DG_MLP_GO(true, true, true); else DG_MLP_GO(true, true, false); }
    else { if (h16) DG_MLP_GO(true, false, true); else DG_MLP_GO(true, false, false); }
  } else {
    if (front) { if (h16) DG_MLP_GO(false, true, true); else DG_MLP_GO(false, true, false); }
    else { if (h16) DG_MLP_GO(false, false, true); else DG_MLP_GO(false, false, false); }
  }
#undef DG_MLP_GO
  DG_LAUNCH_CHECK();
}

// number of token ranges of the weight-gradient launch (grid = HID/128 x splits ~ one wave)
static int bwd_splits(int64_t M, int HID) {
  const int tiles = (int)cdiv(M, 128), NC = HID / HC;
  int S = std::max(1, std::min(tiles, sm_count() / NC));
  const int per = (int)cdiv(tiles, S);
  return (int)cdiv(tiles, per);          // no empty range
}
// floats of partial storage needed by bwd()
static size_t bwd_partial_floats(int64_t M, int HID) {
  return (size_t)bwd_splits(M, HID) * ((size_t)HID * 64 * 2 + HID) + (size_t)cdiv(M, 128) * 64;
}

// dXn[M,64] = d/dx ; dW1 [HID,64], db1 [HID], dW2 [64,HID], db2 [64] (fp32, overwritten; db2 may be null = not wanted).
// dW1 / db1 / dW2 must be adjacent in that order (the parameter arena's order) so that one pass reduces all three.
static void bwd(const bf16* x, const bf16* dy, const bf16* W1, const float* b1, const bf16* W2, float* dXn, float* dW1,
                float* db1, float* dW2, float* db2, float* partial, int64_t M, int HID, cudaStream_t st,
                ReduceList* defer = nullptr, cudaStream_t st_w = nullptr) {
  // st_w: stream of the weight-gradient launch (it feeds nothing but the deferred reduction, so the caller may run it
  // beside the dX chain); null = st
  if (!st_w || !defer) st_w = st;
  if (defer) partial = defer->alloc(bwd_partial_floats(M, HID));
  DG_REQUIRE(db1 == dW1 + (int64_t)HID * 64 && dW2 == db1 + HID, "mlp::bwd: dW1 | db1 | dW2 must be contiguous");
  const int tiles = (int)cdiv(M, 128), NC = HID / HC;
  const int S = bwd_splits(M, HID);
  const int64_t per_split = (int64_t)HID * 64 * 2 + HID;
  MlpBwdArgs a;
  a.M = (int)M; a.HID = HID; a.tiles_per_split = (int)cdiv(tiles, S); a.b1 = b1;
  a.dX = dXn; a.lddx = 64;
  a.dW1_part = partial; a.db1_part = partial + (int64_t)HID * 64; a.dW2_part = a.db1_part + HID;
  a.split_stride = per_split;
  a.db2_part = db2 ? partial + (int64_t)S * per_split : nullptr;
  CUtensorMap tx = make_map(x, 64, M, 64, 64, 128);
  CUtensorMap tdy = make_map(dy, 64, M, 64, 64, 128);
  CUtensorMap tw1 = make_map(W1, 64, HID, 64, 64, 128);
  CUtensorMap tw2 = make_map(W2, HID, 64, HID, 64, 64);
  static DevOnce attr;
  if (attr.first()) {
    DG_CUDA(cudaFuncSetAttribute(mlp_bwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, b::SMEM_TOTAL));
    DG_CUDA(cudaFuncSetAttribute((mlp_bwd_tc_kernel<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, b::SMEM_TOTAL));
    DG_CUDA(cudaFuncSetAttribute(mlp_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, b::SMEM_TOTAL));
  }
  const bool split = g_split_enabled && tiles * f::NSPLIT <= sm_count() && NC % f::NSPLIT == 0;
  if (skip_mask() & SKIP_MLP_BWD_X) {}
  else if (split) launch_k_cluster((mlp_bwd_tc_kernel<0, true>), tiles * f::NSPLIT, THREADS, b::SMEM_TOTAL, st, f::NSPLIT, tx, tdy, tw1, tw2, a);
  else launch_k(mlp_bwd_tc_kernel<0>, tiles, THREADS, b::SMEM_TOTAL, st, tx, tdy, tw1, tw2, a);
  DG_LAUNCH_CHECK();
  if (!(skip_mask() & SKIP_MLP_BWD_W)) launch_k(mlp_bwd_tc_kernel<1>, NC * S, THREADS, b::SMEM_TOTAL, st_w, tx, tdy, tw1, tw2, a);
  DG_LAUNCH_CHECK();
  if (defer) {       // the caller reduces these together with the block's other partial sums
    defer->add(partial, dW1, S, per_split, per_split);
    if (db2) defer->add(a.db2_part, db2, tiles, 64, 64);
    return;
  }
  launch_k(reduce_partials_kernel, reduce_grid(per_split), 256, 0, st, (const float*)partial, dW1, S, per_split);
  DG_LAUNCH_CHECK();
  if (db2) {
    launch_k(reduce_partials_kernel, 1, 64, 0, st, (const float*)a.db2_part, db2, tiles, (int64_t)64);
    DG_LAUNCH_CHECK();
  }
}

}  // namespace mlp
}  // namespace dgvit
