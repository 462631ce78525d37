// mlp_tc.cuh — fused GELU-MLP forward for sm_100a (FeedForward.forward + residual,
// vn/GoalFormer.py:39-50,104):
//
//     out = x + W2 gelu(W1 LN(x) + b1) + b2        (LN applied by the preceding kernel)
//
// One CTA per 128-token tile.  The 2048-wide hidden activation never leaves the SM: for each
// 128-column hidden chunk, GEMM1 (tcgen05, K = 64) lands in TMEM, 16 epilogue warps add the bias,
// apply GELU in fp32 and write the bf16 chunk straight into a 128B-swizzled shared-memory tile that
// is the A operand of GEMM2, whose [128 x 64] fp32 accumulator stays in TMEM across all chunks.
// Weight chunks stream through TMA rings (they are L2 resident: 512 KB per layer).  In training
// passes the pre-activation is additionally written out (bf16, coalesced) for the backward.
//
//   HBM traffic per token: read 128 B (bf16 LN output) + 256 B (fp32 residual), write 256 B
//   (+ 4 KB pre-activation when saved) instead of 8-12 KB for the unfused pair of GEMMs.
#pragma once
#include "attn_tc.cuh"

namespace dgvit {
namespace mlp {

using namespace tc;
using attn::fence_async_smem;
using attn::sw128_off;
using attn::tmem_ld16;

constexpr int HC = 128;            // hidden chunk (columns of GEMM1 / K of GEMM2)
constexpr int NST = 3;             // weight ring depth
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int X_BYTES = 16384, W_BYTES = 16384, H_BYTES = 32768;
constexpr int STAGE_PITCH = 80;
constexpr int OFF_W1 = X_BYTES;
constexpr int OFF_W2 = OFF_W1 + NST * W_BYTES;
constexpr int OFF_H = OFF_W2 + NST * W_BYTES;
constexpr int OFF_BAR = OFF_H + 2 * H_BYTES;
constexpr int OFF_STG = OFF_BAR + 256;
constexpr int SMEM_TOTAL = OFF_STG + EPI_WARPS * 32 * STAGE_PITCH + 1024;
static_assert(SMEM_TOTAL <= 232448, "smem budget");

struct MlpArgs {
  int M, HID;                 // token rows, hidden width
  const float* b1; const float* b2;
  const float* resid; int64_t ldr;
  float* out; int64_t ldc;
  bf16* hpre;                 // [M, HID] or null
};

template <bool SAVE_PRE>
__global__ void __launch_bounds__(THREADS, 1)
mlp_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const MlpArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* x_full = bars;                 // 1
  uint64_t* w1_full = x_full + 1;          // NST
  uint64_t* w1_empty = w1_full + NST;      // NST
  uint64_t* w2_full = w1_empty + NST;      // NST
  uint64_t* w2_empty = w2_full + NST;      // NST
  uint64_t* acc_full = w2_empty + NST;     // 2
  uint64_t* acc_free = acc_full + 2;       // 2 (16 warps)
  uint64_t* h_ready = acc_free + 2;        // 2 (16 warps)
  uint64_t* h_free = h_ready + 2;          // 2
  uint64_t* y_full = h_free + 2;           // 1
  uint32_t* tmem_slot = (uint32_t*)(y_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x;
  const int NC = a.HID / HC;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    mbar_init(x_full, 1);
    for (int i = 0; i < NST; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], EPI_WARPS); mbar_init(&h_ready[i], EPI_WARPS); mbar_init(&h_free[i], 1); }
    mbar_init(y_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;     // acc1: cols [0,256) ; Y: cols [256,320)
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(x_full, X_BYTES);
      tma_load_2d(smem, &tmX, x_full, 0, mt * 128);
      for (int c = 0; c < NC; ++c) {
        const int s = c % NST; const uint32_t ph = (c / NST) & 1;
        mbar_wait(&w1_empty[s], ph ^ 1);
        mbar_expect_tx(&w1_full[s], W_BYTES);
        tma_load_2d(smem + OFF_W1 + s * W_BYTES, &tmW1, &w1_full[s], 0, c * HC);          // [128 hidden rows][64 k]
        mbar_wait(&w2_empty[s], ph ^ 1);
        mbar_expect_tx(&w2_full[s], W_BYTES);
        tma_load_2d(smem + OFF_W2 + s * W_BYTES, &tmW2, &w2_full[s], c * HC, 0);          // [64 d rows][64 k] k-block 0
        tma_load_2d(smem + OFF_W2 + s * W_BYTES + 8192, &tmW2, &w2_full[s], c * HC + 64, 0);  // k-block 1
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(128, HC, false, false);
      constexpr uint32_t idesc2 = make_idesc(128, 64, false, false);
      const uint32_t sx = smem_u32(smem);
      auto gemm1 = [&](int c) {
        const int s = c % NST; const uint32_t ph = (c / NST) & 1;
        const int ab = c & 1; const uint32_t aph = (c >> 1) & 1;
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
      gemm1(0);
      for (int c = 0; c < NC; ++c) {
        if (c + 1 < NC) gemm1(c + 1);
        const int s = c % NST; const uint32_t ph = (c / NST) & 1;
        const int hb = c & 1; const uint32_t hph = (c >> 1) & 1;
        mbar_wait(&h_ready[hb], hph);
        mbar_wait(&w2_full[s], ph);
        tc_fence_after();
        const uint32_t sh = sx + OFF_H + hb * H_BYTES, sw = sx + OFF_W2 + s * W_BYTES;
#pragma unroll
        for (int k = 0; k < HC / 16; ++k) {
          const uint64_t hd = make_smem_desc(sh + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          const uint64_t wd = make_smem_desc(sw + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024);
          umma_bf16(tmem_base + 256, hd, wd, idesc2, (c > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&w2_empty[s]);
        umma_commit(&h_free[hb]);
      }
      umma_commit(y_full);
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3, grp = ew >> 2;            // rows quad*32.., hidden columns grp*32.. of the chunk
    const int r = quad * 32 + lane;
    const int row0 = mt * 128 + quad * 32;
    const bool rows_full = row0 + 32 <= a.M;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    uint8_t* stage_buf = smem + OFF_STG + ew * (32 * STAGE_PITCH);
    const int rr0 = lane >> 2, c16 = lane & 3;
    for (int c = 0; c < NC; ++c) {
      const int ab = c & 1; const uint32_t aph = (c >> 1) & 1;
      mbar_wait(&acc_full[ab], aph);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ab * HC + grp * 32 + lane_off, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free[ab]);          // TMEM chunk drained: GEMM1(c+2) may start
      const int col = c * HC + grp * 32;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.b1 + col) + i);
        v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
      }
      if constexpr (SAVE_PRE) {
        uint8_t* dst = stage_buf + lane * STAGE_PITCH;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(dst + i * 16) = attn_pack8(v + 8 * i);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 8 + rr0;
          const uint4 u = *reinterpret_cast<const uint4*>(stage_buf + rr * STAGE_PITCH + c16 * 16);
          if (rows_full || row0 + rr < a.M)
            *reinterpret_cast<uint4*>(a.hpre + (int64_t)(row0 + rr) * a.HID + col + c16 * 8) = u;
        }
        __syncwarp();
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu3(v[i]);
      mbar_wait(&h_free[ab], aph ^ 1);                   // GEMM2(c-2) has consumed this buffer
      uint8_t* hb = smem + OFF_H + ab * H_BYTES + (grp >> 1) * 16384;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(hb + sw128_off(r, (grp & 1) * 4 + i)) = attn_pack8(v + 8 * i);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_ready[ab]);
    }
    // ---- output: y + b2 + residual  (16 columns per warp)
    mbar_wait(y_full, 0);
    tc_fence_after();
    float y[16];
    tmem_ld16(tmem_base + 256 + grp * 16 + lane_off, y);
    const int row = row0 + lane;
    if (row < a.M) {
      const float* R = a.resid + (int64_t)row * a.ldr + grp * 16;
      float* O = a.out + (int64_t)row * a.ldc + grp * 16;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.b2 + grp * 16) + i);
        const float4 r4 = reinterpret_cast<const float4*>(R)[i];
        reinterpret_cast<float4*>(O)[i] = make_float4(y[4 * i] + b4.x + r4.x, y[4 * i + 1] + b4.y + r4.y,
                                                      y[4 * i + 2] + b4.z + r4.z, y[4 * i + 3] + b4.w + r4.w);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ host
static bool eligible(int D, int HID, int64_t M, const void* x, const void* w1, const void* w2, const float* resid,
                     int64_t ldr, const float* out, int64_t ldc) {
  return tc::g_tc_enabled && D == 64 && HID % HC == 0 && M >= 1 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w1 & 15) == 0 &&
         ((uintptr_t)w2 & 15) == 0 && ((uintptr_t)resid & 15) == 0 && ((uintptr_t)out & 15) == 0 && ldr % 4 == 0 && ldc % 4 == 0;
}

static void fwd(const bf16* x, const bf16* W1, const float* b1, const bf16* W2, const float* b2, const float* resid,
                int64_t ldr, float* out, int64_t ldc, bf16* hpre, int64_t M, int HID, cudaStream_t st) {
  MlpArgs a;
  a.M = (int)M; a.HID = HID; a.b1 = b1; a.b2 = b2; a.resid = resid; a.ldr = ldr; a.out = out; a.ldc = ldc; a.hpre = hpre;
  CUtensorMap tx = make_map(x, 64, M, 64, 64, 128);
  CUtensorMap tw1 = make_map(W1, 64, HID, 64, 64, 128);
  CUtensorMap tw2 = make_map(W2, HID, 64, HID, 64, 64);
  static bool attr = false;
  if (!attr) {
    DG_CUDA(cudaFuncSetAttribute(mlp_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    DG_CUDA(cudaFuncSetAttribute(mlp_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr = true;
  }
  const int grid = (int)cdiv(M, 128);
  if (hpre) launch_k(mlp_fwd_tc_kernel<true>, grid, THREADS, SMEM_TOTAL, st, tx, tw1, tw2, a);
  else launch_k(mlp_fwd_tc_kernel<false>, grid, THREADS, SMEM_TOTAL, st, tx, tw1, tw2, a);
  DG_LAUNCH_CHECK();
}

}  // namespace mlp
}  // namespace dgvit
