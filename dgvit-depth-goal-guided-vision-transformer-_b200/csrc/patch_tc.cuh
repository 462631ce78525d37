// patch_tc.cuh — im2col-free patch embedding for sm_100a (vn/GoalFormer.py:137-139,156-163; D = 64, 16 x 20 patches):
//
//     x = cat(goal_token, Linear(320 -> 64)(Rearrange('b (h p1) (w p2) -> b (h w) (p1 p2)')(img))) + pos ; dropout ; LayerNorm-1
//
// in ONE launch and without a patch matrix in HBM.  A CTA owns 128 consecutive patch tokens (two 128x160 frames): its
// threads read the fp32 frame rows with 16-byte loads (every pixel row of a patch is 80 contiguous bytes), convert to bf16
// and write straight into the 128B-swizzled K-major A tiles of the UMMA ([128 tokens][320] = five [128][64] k-blocks:
// the rearrangement is the shared-memory address), the 40 KB weight arrives by TMA, 20 tcgen05.mma steps form the
// [128 x 64] product in TMEM, and the epilogue adds bias and position embedding, applies the embedding dropout, writes the
// fp32 residual stream and the first block's LayerNorm-1 (bf16 + mean / rstd).  The goal-token rows (fc_embed, ReLU in the
// critic) are computed by two otherwise idle warps while the tensor core runs.  Replaces patchify_kernel + the patch GEMM
// + embed_ln_kernel (3 launches, 10.5 MB of patch matrix written and read back per pass at B = 256).
#pragma once
#include "attn_tc.cuh"

namespace dgvit {
namespace patch {

using namespace tc;
using attn::fence_async_smem;
using attn::sw128_off;

constexpr int D = 64, PH = 16, PW = 20, PD = PH * PW, KBLK = PD / 64;   // 320 = 5 k-blocks of 64
constexpr int THREADS = 256;
constexpr int TILE = 16384;                  // [128 tokens][64] bf16
constexpr int OFF_W = KBLK * TILE;           // W k-blocks: [64 d][64 k] bf16, 8 KB each
constexpr int OFF_LN = OFF_W + KBLK * 8192;  // [128][2] row partials of the fused LayerNorm
constexpr int OFF_BAR = OFF_LN + 1024;
constexpr int SMEM_TOTAL = OFF_BAR + 64 + 1024;

struct PatchArgs {
  const float* img;        // [B, img_h, img_w] fp32 frames, contiguous
  int64_t n_tok;           // B * P patch tokens
  int P, N;                // patches per frame, tokens per sample (P + 1)
  const float* bias;       // to_patch_embedding.1.bias [64]
  const float* pos;        // pos_embedding [N, 64]
  GoalTok gt; float* tok;  // goal token (fc_embed) -> tok [B, 64]
  DropDev drop;
  float* X0;               // [B*N, 64] fp32 residual stream
  bf16* Y;                 // [B*N, 64] LayerNorm-1 output
  const float* gamma; const float* beta;
  float* mean; float* rstd;
};

// frame -> 128B-swizzled K-major tiles [128 tokens][320] (five [128][64] k-blocks at `tiles`) for the 128 patch tokens
// starting at t0.  The tokens are 128 / GW patch rows = 16 * 128 / GW pixel rows of GW * 20 pixels, contiguous in memory
// (frames back to back).  One thread-item = 4 consecutive pixels of one patch row = 4 consecutive k; eight 16-byte loads in
// flight per thread.
template <int GW>
__device__ __forceinline__ void frames_to_tiles(uint8_t* tiles, const float* img, int64_t t0, int64_t n_tok, int tid) {
  constexpr int ROW4 = GW * 5;                           // 16-byte groups per pixel row
  constexpr int ITEMS = (128 / GW) * PH * ROW4;          // 10240
  const float4* src = reinterpret_cast<const float4*>(img) + (t0 / GW) * PH * ROW4;
  constexpr int U = 8;
  for (int base = tid; base < ITEMS; base += THREADS * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * THREADS;
      const int yy = i / ROW4, q = i % ROW4;
      const int r = (yy / PH) * GW + (4 * q) / PW;
      v[u] = (i < ITEMS && t0 + r < n_tok) ? __ldcg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);      // (frames: written by the gather right before)
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * THREADS;
      if (i >= ITEMS) continue;
      const int yy = i / ROW4, q = i % ROW4;
      const int r = (yy / PH) * GW + (4 * q) / PW;       // token row of the tile
      const int k = (yy % PH) * PW + (4 * q) % PW;       // k = p1 * 20 + p2, a multiple of 4
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v[u].x, v[u].y), hi = __floats2bfloat162_rn(v[u].z, v[u].w);
      uint8_t* dst = tiles + (k >> 6) * TILE + sw128_off(r, (k & 63) >> 3) + (k & 4) * 2;
      *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
}

template <int GW>      // patches per frame row (img_w / 20): 8 at 160 pixels, 16 at 320
__global__ void __launch_bounds__(THREADS, 1)
patch_embed_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ PatchArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* w_full = (uint64_t*)(smem + OFF_BAR);
  uint64_t* mma_done = w_full + 1;
  uint32_t* tmem_slot = (uint32_t*)(mma_done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.x * 128;            // first patch token of this CTA

  if (tid == 0) {
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1); mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  if (warp == 0 && elect_one_sync()) {
    mbar_expect_tx(w_full, KBLK * 8192);
#pragma unroll
    for (int kb = 0; kb < KBLK; ++kb) tma_load_2d(smem + OFF_W + kb * 8192, &tmW, w_full, kb * 64, 0);
  }
  frames_to_tiles<GW>(smem, a.img, t0, a.n_tok, tid);
  fence_async_smem();
  __syncthreads();
  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_wait(w_full, 0);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc(128, D, false, false);
      const uint32_t sa = smem_u32(smem), sw = smem_u32(smem + OFF_W);
#pragma unroll
      for (int kb = 0; kb < KBLK; ++kb) {
        const uint64_t ad = make_smem_desc(sa + kb * TILE, 16, 1024), wd = make_smem_desc(sw + kb * 8192, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + (uint64_t)(k * 2), wd + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(mma_done);
    }
    __syncwarp();
  } else if (warp == 2 || warp == 3) {
    // goal-token rows of the samples that start inside this tile (two at P = 64), while the tensor core runs
    const int64_t b_first = (t0 + a.P - 1) / a.P;
    const int64_t b_last = (t0 + 127) / a.P;               // inclusive
    for (int64_t b = b_first + (warp - 2); b <= b_last && b * a.P < a.n_tok; b += 2) {
      const int64_t row = b * a.N;
      float x[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int d = lane + 32 * i;
        float v;
        if (a.gt.direct) {
          v = a.gt.direct[b * D + d];
        } else {
          v = a.gt.b[d];
          for (int j = 0; j < a.gt.nps; ++j) v = fmaf(a.gt.W[d * a.gt.nps + j], a.gt.ps[b * a.gt.nps + j], v);
          if (a.gt.relu) v = fmaxf(v, 0.f);
        }
        a.tok[b * D + d] = v;
        v = (v + a.pos[d]) * drop_factor(a.drop, row * D + d);
        a.X0[row * D + d] = v;
        x[i] = v;
      }
      const float mu = warp_sum(x[0] + x[1]) * (1.0f / D);
      const float c0 = x[0] - mu, c1 = x[1] - mu;
      const float rs = 1.0f / sqrtf(warp_sum(fmaf(c0, c0, c1 * c1)) * (1.0f / D) + 1e-5f);
      a.Y[row * D + lane] = __float2bfloat16_rn(c0 * rs * a.gamma[lane] + a.beta[lane]);
      a.Y[row * D + lane + 32] = __float2bfloat16_rn(c1 * rs * a.gamma[lane + 32] + a.beta[lane + 32]);
      if (lane == 0 && a.mean) { a.mean[row] = mu; a.rstd[row] = rs; }
    }
  }
  // ---- epilogue: 8 warps = 4 TMEM lane quadrants x 2 column halves
  mbar_wait(mma_done, 0);
  tc_fence_after();
  {
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;
    const int64_t t = t0 + r;
    const bool ok = t < a.n_tok;
    float v[32];
    tmem_ld32(tmem_base + half * 32 + ((uint32_t)(quad * 32) << 16), v);
    const int64_t b = ok ? t / a.P : 0;
    const int p = ok ? (int)(t % a.P) : 0;
    const int64_t row = b * a.N + 1 + p;
    const int c0 = half * 32;
    float s = 0.f;
    if (ok) {
      const float* pos = a.pos + (int64_t)(p + 1) * D + c0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + c0) + i);
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(pos) + i);
        const float4 f4 = drop_factor4(a.drop, row * D + c0 + 4 * i);
        v[4 * i] = (v[4 * i] + b4.x + p4.x) * f4.x; v[4 * i + 1] = (v[4 * i + 1] + b4.y + p4.y) * f4.y;
        v[4 * i + 2] = (v[4 * i + 2] + b4.z + p4.z) * f4.z; v[4 * i + 3] = (v[4 * i + 3] + b4.w + p4.w) * f4.w;
        reinterpret_cast<float4*>(a.X0 + row * D + c0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) s += v[i];
    }
    // LayerNorm over the 64 columns of a row: this warp holds 32, the warp of the other half (same quadrant) the rest
    float* part = reinterpret_cast<float*>(smem + OFF_LN);
    part[r * 2 + half] = s;
    __syncthreads();
    const float mu = (part[r * 2] + part[r * 2 + 1]) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float c = v[i] - mu; q = fmaf(c, c, q); }
    __syncthreads();
    part[r * 2 + half] = q;
    __syncthreads();
    const float rs = 1.0f / sqrtf((part[r * 2] + part[r * 2 + 1]) * (1.0f / D) + 1e-5f);
    if (ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + 8 * i + 2 * j;
          const float2 g2 = __ldg(reinterpret_cast<const float2*>(a.gamma + c)), b2 = __ldg(reinterpret_cast<const float2*>(a.beta + c));
          const __nv_bfloat162 y = __floats2bfloat162_rn((v[8 * i + 2 * j] - mu) * rs * g2.x + b2.x, (v[8 * i + 2 * j + 1] - mu) * rs * g2.y + b2.y);
          w[j] = *reinterpret_cast<const uint32_t*>(&y);
        }
        reinterpret_cast<uint4*>(a.Y + row * D + c0)[i] = make_uint4(w[0], w[1], w[2], w[3]);
      }
      if (half == 0 && a.mean) { a.mean[row] = mu; a.rstd[row] = rs; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// ------------------------------------------------------------------ host
static bool g_enabled = true;       // set_option "patch_fused"
static bool g_dw_enabled = true;    // set_option "patch_dw": patch-weight gradient from the frames (no patch matrix)
static bool eligible(const dgvit_cfg& cfg, const void* img, const void* w, const void* y) {
  const int gw = cfg.img_w / PW;
  const int P = (cfg.img_h / PH) * gw;
  return tc::g_tc_enabled && g_enabled && cfg.dim == D && cfg.patch_h == PH && cfg.patch_w == PW && (gw == 8 || gw == 16) &&
         cfg.img_w % PW == 0 && cfg.img_h % PH == 0 && (128 % P == 0 || P % 128 == 0) &&
         ((((uintptr_t)img) | ((uintptr_t)w) | ((uintptr_t)y)) & 15) == 0;
}

static void fwd(const dgvit_cfg& cfg, const PatchArgs& a, const bf16* W, cudaStream_t st) {
  CUtensorMap tw = make_map(W, PD, D, PD, 64, 64);
  static DevOnce attr;
  if (attr.first()) {
    DG_CUDA(cudaFuncSetAttribute(patch_embed_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    DG_CUDA(cudaFuncSetAttribute(patch_embed_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  }
  const unsigned grid = (unsigned)cdiv(a.n_tok, 128);
  if (cfg.img_w / PW == 8) launch_k(patch_embed_tc_kernel<8>, grid, THREADS, SMEM_TOTAL, st, tw, a);
  else launch_k(patch_embed_tc_kernel<16>, grid, THREADS, SMEM_TOTAL, st, tw, a);
  DG_LAUNCH_CHECK();
}

// ------------------------------------------------------------------ patch-weight gradient, no patch matrix in HBM
// dW[64][320] = sum over tokens dXp[t][:] (x) patch[t][:]  (vn/GoalFormer.py:139 backward).  Computed transposed so that the
// wide dimension rides on UMMA M: dW^T[320 x 64] = tiles^T . dXp.  A CTA walks its token tiles: the threads rebuild the
// [128 tokens][320] bf16 tiles from the frames in shared memory (same image as the forward's A operand, read MN-major
// here: three M = 128 blocks, the sixth 64-column block is padding whose accumulator rows are never stored), the dXp tile
// [128 tokens][64] arrives by TMA and is the MN-major B operand; the three [128 x 64] accumulators stay in TMEM across
// the CTA's tiles.  One partial per CTA, summed by the caller's deferred reduction (fixed order).
namespace dw {
constexpr int OFF_DX = 6 * TILE;             // dXp ring: 2 x [128][64] bf16
constexpr int OFF_BAR = OFF_DX + 2 * TILE;
constexpr int SMEM_TOTAL = OFF_BAR + 64 + 1024;
}  // namespace dw

struct PatchDwArgs {
  const float* img; int64_t n_tok;
  int tiles_per_cta;
  float* partial;          // [gridDim.x][64][320]
};

template <int GW>
__global__ void __launch_bounds__(THREADS, 1)
patch_dw_tc_kernel(const __grid_constant__ CUtensorMap tmDX, const __grid_constant__ PatchDwArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* dx_full = (uint64_t*)(smem + dw::OFF_BAR);    // [2]
  uint64_t* mma_done = dx_full + 2;                        // 1
  uint32_t* tmem_slot = (uint32_t*)(mma_done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int64_t tiles = (a.n_tok + 127) / 128;
  const int64_t tile0 = (int64_t)blockIdx.x * a.tiles_per_cta;
  const int nt = (int)min((int64_t)a.tiles_per_cta, tiles - tile0);

  if (tid == 0) {
    tma_prefetch_desc(&tmDX);
    mbar_init(&dx_full[0], 1); mbar_init(&dx_full[1], 1); mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  // the padding block (columns 320..383 of the tiles) is never written by frames_to_tiles: zero it once
  for (int i = tid; i < TILE / 16; i += THREADS) reinterpret_cast<uint4*>(smem + 5 * TILE)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  if (warp == 0 && elect_one_sync() && nt > 0) {
    mbar_expect_tx(&dx_full[0], TILE);
    tma_load_2d(smem + dw::OFF_DX, &tmDX, &dx_full[0], 0, (int)(tile0 * 128));
  }
  for (int t = 0; t < nt; ++t) {
    if (t > 0) mbar_wait(mma_done, (uint32_t)((t - 1) & 1));      // the MMAs of the previous tile have read the tiles
    if (warp == 0 && elect_one_sync() && t + 1 < nt) {              // next dXp tile (its buffer was read two tiles ago)
      mbar_expect_tx(&dx_full[(t + 1) & 1], TILE);
      tma_load_2d(smem + dw::OFF_DX + ((t + 1) & 1) * TILE, &tmDX, &dx_full[(t + 1) & 1], 0, (int)((tile0 + t + 1) * 128));
    }
    frames_to_tiles<GW>(smem, a.img, (tile0 + t) * 128, a.n_tok, tid);
    fence_async_smem();
    __syncthreads();
    if (warp == 0 && elect_one_sync()) {
      mbar_wait(&dx_full[t & 1], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc(128, D, true, true);
      const uint32_t sa = smem_u32(smem), sdx = smem_u32(smem + dw::OFF_DX + (t & 1) * TILE);
#pragma unroll
      for (int mb = 0; mb < 3; ++mb) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {          // 16 tokens per step
          const uint64_t ad = make_smem_desc(sa + mb * 2 * TILE + k * 2048, TILE, 1024);     // tiles^T: 64-wide M blocks TILE apart
          const uint64_t bd = make_smem_desc(sdx + k * 2048, 8192, 1024);
          umma_bf16(tmem_base + mb * 64, ad, bd, idesc, (t > 0 || k > 0) ? 1u : 0u);
        }
      }
      umma_commit(mma_done);
    }
    __syncwarp();
  }
  if (nt > 0) {
    mbar_wait(mma_done, (uint32_t)((nt - 1) & 1));
    tc_fence_after();
    const int quad = warp & 3, half = warp >> 2;
    float* out = a.partial + (int64_t)blockIdx.x * D * PD;
#pragma unroll 1
    for (int mb = 0; mb < 3; ++mb) {
      float v[32];
      tmem_ld32(tmem_base + mb * 64 + half * 32 + ((uint32_t)(quad * 32) << 16), v);
      const int k = mb * 128 + quad * 32 + lane;             // row of dW^T = column of dW
      if (k < PD) {
#pragma unroll
        for (int i = 0; i < 32; ++i) out[(int64_t)(half * 32 + i) * PD + k] = v[i];      // lanes = consecutive k: coalesced
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// number of CTAs / partials of the weight-gradient launch
static int dw_ctas(int64_t n_tok) {
  const int64_t tiles = cdiv(n_tok, 128);
  const int per = (int)cdiv(tiles, 74);          // about two tiles per CTA at B = 256, half a wave
  return (int)cdiv(tiles, per);
}
static size_t dw_partial_floats(int64_t n_tok) { return (size_t)dw_ctas(n_tok) * D * PD; }

// dW_patch [64][320] (fp32, overwritten through the deferred reduction) from the frames and dXp [n_tok][64] bf16
static void bwd_w(const dgvit_cfg& cfg, const float* img, const bf16* dXp, int64_t n_tok, float* dW, ReduceList& rl, cudaStream_t st) {
  const int S = dw_ctas(n_tok);
  PatchDwArgs a;
  a.img = img; a.n_tok = n_tok; a.tiles_per_cta = (int)cdiv(cdiv(n_tok, 128), S);
  a.partial = rl.alloc((size_t)S * D * PD);
  CUtensorMap tdx = make_map(dXp, D, n_tok, D, 64, 128);
  static DevOnce attr;
  if (attr.first()) {
    DG_CUDA(cudaFuncSetAttribute(patch_dw_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw::SMEM_TOTAL));
    DG_CUDA(cudaFuncSetAttribute(patch_dw_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw::SMEM_TOTAL));
  }
  if (cfg.img_w / PW == 8) launch_k(patch_dw_tc_kernel<8>, S, THREADS, dw::SMEM_TOTAL, st, tdx, a);
  else launch_k(patch_dw_tc_kernel<16>, S, THREADS, dw::SMEM_TOTAL, st, tdx, a);
  DG_LAUNCH_CHECK();
  rl.add(a.partial, dW, S, (int64_t)D * PD, (int64_t)D * PD);
}

}  // namespace patch
}  // namespace dgvit
