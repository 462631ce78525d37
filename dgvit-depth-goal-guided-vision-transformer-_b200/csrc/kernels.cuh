// kernels.cuh — non-GEMM kernels of the DGViT hot path (embedding, LayerNorm, attention,
// RMSNorm pooling, tanh-Gaussian head, SAC losses, Adam/Polyak, replay gather).
// Reference lines are cited per kernel; "vn/" = src/vis_nav/vis_nav/.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace dgvit {

// =====================================================================================
// Embedding  (vn/GoalFormer.py:137-139,156-163 ; vn/got_sac_network.py:111,226)
// =====================================================================================

// Rearrange 'b (h p1) (w p2) -> b (h w) (p1 p2)' materialised once per forward in the
// activation dtype (the fp32->bf16 cast of the frame is fused here).
template <typename A>
__global__ void patchify_kernel(const float* __restrict__ img, A* __restrict__ out, int64_t total4,
                                int img_h, int img_w, int ph, int pw) {
  pdl_wait();
  pdl_launch();
  // one thread = 4 consecutive pixels of one patch row (pw % 4 == 0, img_w % 4 == 0)
  const int gw = img_w / pw, gh = img_h / ph;
  const int q4 = pw / 4, pd4 = ph * q4, P = gh * gw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int k4 = (int)(i % pd4);
    const int64_t bp = i / pd4;
    const int p = (int)(bp % P);
    const int64_t b = bp / P;
    const int p1 = k4 / q4, q = k4 % q4;
    const int y = (p / gw) * ph + p1, x = (p % gw) * pw + 4 * q;
    const float4 v = *reinterpret_cast<const float4*>(img + (b * img_h + y) * img_w + x);
    A* o = out + i * 4;
    stf(o, v.x); stf(o + 1, v.y); stf(o + 2, v.z); stf(o + 3, v.w);
  }
}

// goal token: tok[b,:] = act(W_e pstate[b] + b_e); ReLU only in the critic
__global__ void goal_embed_kernel(const float* __restrict__ ps, const float* __restrict__ W,
                                  const float* __restrict__ bias, float* __restrict__ tok, int B, int D,
                                  int npst, int relu) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i % D;
  float v = bias[d];
  for (int j = 0; j < npst; ++j) v = fmaf(W[d * npst + j], ps[b * npst + j], v);
  tok[i] = relu ? fmaxf(v, 0.f) : v;
}

// x = cat(tok, patches) + pos ; x = dropout(x)     (GoalFormer.py:160-163)
__global__ void embed_assemble_kernel(const float* __restrict__ tok, const float* __restrict__ Xp,
                                      const float* __restrict__ pos, float* __restrict__ X0, DropDev drop,
                                      int64_t total, int N, int D) {
  pdl_wait();
  pdl_launch();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const int64_t bn = i / D;
    const int n = (int)(bn % N);
    const int64_t b = bn / N;
    const float v = (n == 0 ? tok[b * D + d] : Xp[(b * (N - 1) + (n - 1)) * D + d]) + pos[n * D + d];
    X0[i] = v * drop_factor(drop, i);
  }
}

// The same fused with the goal-token embedding (fc_embed, optional ReLU; vn/got_sac_network.py:111,224) and the first
// block's LayerNorm-1: one warp per token row (D = 32 * VPL).  Writes tok [B,D], X0 (fp32 residual stream), LN(X0) in the
// operand dtype, mean, rstd.
struct GoalTok {
  const float* ps; const float* W; const float* b;   // pstate [B,nps], fc_embed.weight [D,nps], .bias [D]
  int nps, relu;
  const float* direct = nullptr;                     // GoT.forward(img, goal): the goal token itself [B,D] (ps / W / b unused)
};
template <typename A, int VPL>
__global__ void embed_ln_kernel(GoalTok gt, float* __restrict__ tok, const float* __restrict__ Xp,
                                const float* __restrict__ pos, float* __restrict__ X0, DropDev drop,
                                const float* __restrict__ gamma, const float* __restrict__ beta, A* __restrict__ Y,
                                float* __restrict__ mean, float* __restrict__ rstd, int64_t T, int N) {
  pdl_wait();
  pdl_launch();
  constexpr int D = 32 * VPL;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T) return;
  const int n = (int)(row % N);
  const int64_t b = row / N;
  float v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int d = lane + 32 * i;
    float x;
    if (n == 0) {
      if (gt.direct) {
        x = gt.direct[b * D + d];
      } else {
        x = gt.b[d];
        for (int j = 0; j < gt.nps; ++j) x = fmaf(gt.W[d * gt.nps + j], gt.ps[b * gt.nps + j], x);
        if (gt.relu) x = fmaxf(x, 0.f);
      }
      tok[b * D + d] = x;
    } else {
      x = Xp[(b * (N - 1) + (n - 1)) * D + d];
    }
    x = (x + pos[n * D + d]) * drop_factor(drop, row * D + d);
    X0[row * D + d] = x;
    v[i] = x;
    s += x;
  }
  const float mu = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) { const float c = v[i] - mu; q = fmaf(c, c, q); }
  const float rs = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int d = lane + 32 * i;
    stf(Y + row * D + d, (v[i] - mu) * rs * gamma[d] + beta[d]);
  }
  if (lane == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
}

// backward of the above: dXp (compact, activation dtype), dtok (through the optional ReLU)
template <typename A>
__global__ void embed_bwd_kernel(const float* __restrict__ dX0, const float* __restrict__ tok,
                                 A* __restrict__ dXp, float* __restrict__ dtok, DropDev drop, int64_t total,
                                 int N, int D, int relu) {
  pdl_wait();
  pdl_launch();
  // one thread = 4 consecutive features of one token (D % 4 == 0): one 16-byte load, one Philox call; 32-bit index
  // arithmetic when the tensor allows it (two 64-bit divisions per element cost more than the copy)
  const int64_t total4 = total >> 2;
  const int D4 = D >> 2;
  for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < total4; i4 += (int64_t)gridDim.x * blockDim.x) {
    int d4, n;
    int64_t b;
    if (total4 < ((int64_t)1 << 31)) {
      const unsigned u = (unsigned)i4, bn = u / D4, bb = bn / N;
      d4 = u - bn * D4; n = bn - bb * N; b = bb;
    } else {
      d4 = (int)(i4 % D4);
      const int64_t bn = i4 / D4;
      n = (int)(bn % N); b = bn / N;
    }
    const float4 f = drop_factor4(drop, i4 * 4);
    float4 g = reinterpret_cast<const float4*>(dX0)[i4];
    g.x *= f.x; g.y *= f.y; g.z *= f.z; g.w *= f.w;
    if (n == 0) {
      float* o = dtok + b * D + 4 * d4;
      if (relu) {
        const float4 t = *reinterpret_cast<const float4*>(tok + b * D + 4 * d4);
        g.x = t.x > 0.f ? g.x : 0.f; g.y = t.y > 0.f ? g.y : 0.f; g.z = t.z > 0.f ? g.z : 0.f; g.w = t.w > 0.f ? g.w : 0.f;
      }
      *reinterpret_cast<float4*>(o) = g;
    } else {
      A* o = dXp + (b * (N - 1) + (n - 1)) * D + 4 * d4;
      stf(o, g.x); stf(o + 1, g.y); stf(o + 2, g.z); stf(o + 3, g.w);
    }
  }
}

// dpos partials: part[s][n,d] = sum over the s-th batch chunk of dX0[b,n,d]*keep (fixed order);
// grid (N, S); a reduce_partials pass finishes the sum deterministically.
__global__ void dpos_kernel(const float* __restrict__ dX0, float* __restrict__ part, DropDev drop, int B,
                            int N, int D) {
  pdl_wait();
  pdl_launch();
  const int n = blockIdx.x, s = blockIdx.y, S = gridDim.y;
  const int per = (B + S - 1) / S;
  const int b0 = s * per, b1 = min(B, b0 + per);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) {
      const int64_t i = ((int64_t)b * N + n) * D + d;
      acc += dX0[i] * drop_factor(drop, i);
    }
    part[((int64_t)s * N + n) * D + d] = acc;
  }
}

// the keep decisions of drop_factor as a {0,1} mask (dgvit_debug_drop_mask: tests)
__global__ void drop_mask_dump_kernel(DropDev drop, uint8_t* __restrict__ out, int64_t total) {
  pdl_wait();
  pdl_launch();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = drop_factor(drop, i) != 0.0f ? 1 : 0;
}

// =====================================================================================
// LayerNorm (vn/GoalFormer.py:31-37; eps 1e-5, affine) — one warp per token row
// =====================================================================================
template <typename A, int VPL>  // D = 32*VPL
__global__ void layernorm_fwd_kernel(const float* __restrict__ X, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, A* __restrict__ Y,
                                     float* __restrict__ mean, float* __restrict__ rstd, int64_t T) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T) return;
  constexpr int D = 32 * VPL;
  float v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) { v[i] = X[row * D + lane + 32 * i]; s += v[i]; }
  const float mu = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) { const float c = v[i] - mu; q = fmaf(c, c, q); }
  const float var = warp_sum(q) * (1.0f / D);
  const float rs = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int d = lane + 32 * i;
    stf(Y + row * D + d, (v[i] - mu) * rs * gamma[d] + beta[d]);
  }
  if (lane == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
}

// dX_io[row] += LN'(dY[row]); per-block partial sums of dgamma / dbeta / column sums of the updated dX_io
// -> part[block][3][D].  The third vector is the bias gradient of the linear layer that produced this residual
// stream (to_out.0.bias / net.3.bias): colsum(dL/dX) in fp32, for free instead of a separate pass over dX.
template <int VPL>
__global__ void __launch_bounds__(512) layernorm_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ X,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     const float* __restrict__ gamma, float* __restrict__ dX_io,
                                     bf16* __restrict__ dX_lp, float* __restrict__ part, int64_t T,
                                     const float* __restrict__ dX_row0, int Ntok) {
  // dX_row0 (optional, [T / Ntok, D]): the incoming residual gradient is zero except on token 0 of every sample, where it
  // is this compact array (last transformer block); dX_io is then write-only.
  pdl_wait();
  pdl_launch();
  constexpr int D = 32 * VPL;
  constexpr int RB = VPL <= 2 ? 4 : (VPL <= 4 ? 2 : 1);   // rows in flight per warp: the kernel is a chain of load latencies otherwise
  extern __shared__ float sm[];  // [warps][3][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float dg[VPL], db[VPL], dxs[VPL], gm[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) { dg[i] = 0.f; db[i] = 0.f; dxs[i] = 0.f; gm[i] = gamma[lane + 32 * i]; }
  const int64_t stride = (int64_t)gridDim.x * nw;
  for (int64_t row0 = (int64_t)blockIdx.x * nw + warp; row0 < T; row0 += RB * stride) {
    float mu[RB], rs[RB], xh[RB][VPL], dy[RB][VPL], dxin[RB][VPL];
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int64_t row = row0 + b * stride;
      const bool ok = row < T;
      mu[b] = ok ? mean[row] : 0.f; rs[b] = ok ? rstd[row] : 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int64_t e = row * D + lane + 32 * i;
        xh[b][i] = ok ? X[e] : 0.f; dy[b][i] = ok ? dY[e] : 0.f;
        if (dX_row0) dxin[b][i] = (ok && row % Ntok == 0) ? dX_row0[(row / Ntok) * D + lane + 32 * i] : 0.f;
        else dxin[b][i] = ok ? dX_io[e] : 0.f;
      }
    }
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int64_t row = row0 + b * stride;
      if (row >= T) break;                      // warp-uniform
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        xh[b][i] = (xh[b][i] - mu[b]) * rs[b];
        const float w = dy[b][i] * gm[i];
        s1 += w;
        s2 = fmaf(w, xh[b][i], s2);
        dg[i] = fmaf(dy[b][i], xh[b][i], dg[i]);
        db[i] += dy[b][i];
      }
      s1 = warp_sum(s1) * (1.0f / D);
      s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int d = lane + 32 * i;
        const float nv = dxin[b][i] + rs[b] * (dy[b][i] * gm[i] - s1 - xh[b][i] * s2);
        dX_io[row * D + d] = nv;
        dxs[i] += nv;
        if (dX_lp) dX_lp[row * D + d] = __float2bfloat16_rn(nv);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    sm[(warp * 3 + 0) * D + lane + 32 * i] = dg[i];
    sm[(warp * 3 + 1) * D + lane + 32 * i] = db[i];
    sm[(warp * 3 + 2) * D + lane + 32 * i] = dxs[i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 3 * D; j += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += sm[w * 3 * D + j];
    part[(int64_t)blockIdx.x * 3 * D + j] = s;
  }
}

// dgamma[d] = sum_blocks part[b][0][d], dbeta[d] = sum_blocks part[b][1][d], dxsum[d] = sum_blocks part[b][2][d]
// block = 32 columns x 32 row-groups (1024 threads); fixed summation order (deterministic)
__global__ void __launch_bounds__(1024) ln_param_reduce_kernel(const float* __restrict__ part,
                                                               float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta,
                                                               float* __restrict__ dxsum, int nblocks, int D) {
  pdl_wait();
  pdl_launch();
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cx;
  float s0 = 0.f, s1 = 0.f;
  if (j < 3 * D) {
    int b = gy;
    for (; b + 32 < nblocks; b += 64) {
      s0 += part[(int64_t)b * 3 * D + j];
      s1 += part[(int64_t)(b + 32) * 3 * D + j];
    }
    if (b < nblocks) s0 += part[(int64_t)b * 3 * D + j];
  }
  sm[gy][cx] = s0 + s1;
  __syncthreads();
  if (gy == 0 && j < 3 * D) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) t += sm[g][cx];
    if (j < D) dgamma[j] = t;
    else if (j < 2 * D) dbeta[j - D] = t;
    else if (dxsum) dxsum[j - 2 * D] = t;
  }
}

// column sums (bias gradients): part[s][n] = sum over this block's row range.
// thread = 2 adjacent columns (4-byte bf16x2 / 8-byte float2 loads), 4 row-interleaved accumulators.
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld2(const bf16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
template <typename T_>
__global__ void colsum_partial_kernel(const T_* __restrict__ A, int64_t lda, float* __restrict__ part,
                                      int64_t rows, int N, int64_t rows_per_block) {
  pdl_wait();
  pdl_launch();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (n >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  const bool pair = (n + 1 < N) && ((lda & 1) == 0) && ((((uintptr_t)A) & 7) == 0);
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f, c0 = 0.f, c1 = 0.f, d0 = 0.f, d1 = 0.f;
  if (pair) {
    int64_t r = r0;
    for (; r + 3 < r1; r += 4) {
      const float2 x0 = ld2(A + r * lda + n), x1 = ld2(A + (r + 1) * lda + n);
      const float2 x2 = ld2(A + (r + 2) * lda + n), x3 = ld2(A + (r + 3) * lda + n);
      a0 += x0.x; a1 += x0.y; b0 += x1.x; b1 += x1.y; c0 += x2.x; c1 += x2.y; d0 += x3.x; d1 += x3.y;
    }
    for (; r < r1; ++r) { const float2 x = ld2(A + r * lda + n); a0 += x.x; a1 += x.y; }
    part[(int64_t)blockIdx.y * N + n] = (a0 + b0) + (c0 + d0);
    part[(int64_t)blockIdx.y * N + n + 1] = (a1 + b1) + (c1 + d1);
  } else {
    for (int64_t r = r0; r < r1; ++r) {
      a0 += ldf(A + r * lda + n);
      if (n + 1 < N) a1 += ldf(A + r * lda + n + 1);
    }
    part[(int64_t)blockIdx.y * N + n] = a0;
    if (n + 1 < N) part[(int64_t)blockIdx.y * N + n + 1] = a1;
  }
}

// The same for long, narrow matrices (the CNN critic's conv bias gradients: 274k rows x 64 columns): with one thread per
// column pair the launch above has 32 active threads per block.  Here a block covers N/2 column pairs x (256 / (N/2)) row
// lanes (a warp reads one or more whole rows), four rows in flight per thread, row lanes summed through shared memory
// in a fixed order.  N even, N/2 a divisor of 256.
template <typename T_>
__global__ void __launch_bounds__(256) colsum_tall_kernel(const T_* __restrict__ A, int64_t lda, float* __restrict__ part,
                                                          int64_t rows, int N, int64_t rows_per_block) {
  pdl_wait();
  pdl_launch();
  __shared__ float2 sm[256];
  const int cp = N / 2, nrl = 256 / cp;
  const int c = threadIdx.x % cp, rl = threadIdx.x / cp;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
  int64_t r = r0 + rl;
  for (; r + 3 * nrl < r1; r += 4 * nrl) {
    const float2 x0 = ld2(A + r * lda + 2 * c), x1 = ld2(A + (r + nrl) * lda + 2 * c);
    const float2 x2 = ld2(A + (r + 2 * nrl) * lda + 2 * c), x3 = ld2(A + (r + 3 * nrl) * lda + 2 * c);
    s0.x += x0.x; s0.y += x0.y; s1.x += x1.x; s1.y += x1.y; s2.x += x2.x; s2.y += x2.y; s3.x += x3.x; s3.y += x3.y;
  }
  for (; r < r1; r += nrl) { const float2 x = ld2(A + r * lda + 2 * c); s0.x += x.x; s0.y += x.y; }
  sm[threadIdx.x] = make_float2((s0.x + s1.x) + (s2.x + s3.x), (s0.y + s1.y) + (s2.y + s3.y));
  __syncthreads();
  if (rl == 0) {
    float2 t = sm[c];
    for (int k = 1; k < nrl; ++k) { const float2 v = sm[k * cp + c]; t.x += v.x; t.y += v.y; }
    part[(int64_t)blockIdx.x * N + 2 * c] = t.x;
    part[(int64_t)blockIdx.x * N + 2 * c + 1] = t.y;
  }
}

// =====================================================================================
// Last-block attention.  Only token 0 of the last block's output is consumed (x[:, 0], vn/GoalFormer.py:167), so
// there the attention needs a single query row per (sample, head): o0 = softmax(q0 K^T * dh^-0.5) V.  K and V
// still come from every token, so gradients reach all of them; dQ is zero except on row 0.  Exact, not an
// approximation.  One warp per (sample, head), dh = 64, up to 72 keys (the shipped model has 65):
// lane = (g, c): key slot g of 8 per pass, quarter c of the 64 head dims -> every load is a 32-byte piece of a row and
// all loads of a phase are independent (the kernel is a latency chain: q, K, softmax, V).
// =====================================================================================
constexpr int R0_DH = 64, R0_NP = 9, R0_MAXN = 8 * R0_NP;
template <typename A> __device__ __forceinline__ void r0_ld16(const A* p, float (&v)[16]);
template <> __device__ __forceinline__ void r0_ld16<float>(const float* p, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 f = reinterpret_cast<const float4*>(p)[i];
    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
  }
}
template <> __device__ __forceinline__ void r0_ld16<bf16>(const bf16* p, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint4 u = reinterpret_cast<const uint4*>(p)[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[8 * i + 2 * j] = __uint_as_float(w[j] << 16); v[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
  }
}
template <typename A> __device__ __forceinline__ void r0_st16(A* p, const float (&v)[16]);
template <> __device__ __forceinline__ void r0_st16<float>(float* p, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <> __device__ __forceinline__ void r0_st16<bf16>(bf16* p, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(p)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// x[t] = <vec quarter, row (8t+g) quarter> summed over the 4 quarter lanes (rows >= N give 0)
template <typename A>
__device__ __forceinline__ void r0_dots(const float (&vec)[16], const A* rows, int ld, int N, int g, float (&x)[R0_NP]) {
#pragma unroll
  for (int t = 0; t < R0_NP; ++t) {
    const int j = 8 * t + g;
    float s = 0.f;
    if (j < N) {
      float k[16];
      r0_ld16<A>(rows + (int64_t)j * ld, k);
#pragma unroll
      for (int i = 0; i < 16; ++i) s = fmaf(vec[i], k[i], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    x[t] = s;
  }
}
__device__ __forceinline__ float r0_groups_sum(float v) {   // over the 8 key slots (each value is replicated on 4 lanes)
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
// p[t] = softmax weight of key 8t+g (unnormalised exp; returns 1 / sum)
__device__ __forceinline__ float r0_softmax(float (&p)[R0_NP], int N, int g, float scale) {
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < R0_NP; ++t) if (8 * t + g < N) mx = fmaxf(mx, p[t]);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < R0_NP; ++t) {
    p[t] = (8 * t + g < N) ? expf((p[t] - mx) * scale) : 0.f;
    sum += p[t];
  }
  return 1.0f / r0_groups_sum(sum);
}
template <typename A>
__global__ void __launch_bounds__(128) attention_row0_fwd_kernel(const A* __restrict__ QKV, A* __restrict__ O, int items, int N,
                                                                 int H, float scale) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (it >= items) return;
  const int b = it / H, h = it % H, inner = H * R0_DH, ld = 3 * inner;
  const A* base = QKV + (int64_t)b * N * ld + h * R0_DH;
  float q[16], p[R0_NP];
  r0_ld16<A>(base + 16 * c, q);
  r0_dots<A>(q, base + inner + 16 * c, ld, N, g, p);
  const float inv = r0_softmax(p, N, g, scale);
  // o0[d] = sum_j p_j V[j][d]: each lane sums its own keys over its quarter of the head dims, then the 8 key slots meet
  float o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = 0.f;
#pragma unroll
  for (int t = 0; t < R0_NP; ++t) {
    const int j = 8 * t + g;
    if (j < N) {
      float v[16];
      r0_ld16<A>(base + 2 * inner + (int64_t)j * ld + 16 * c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = fmaf(p[t], v[i], o[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = r0_groups_sum(o[i]) * inv;
  if (g == 0) r0_st16<A>(O + (int64_t)b * N * inner + h * R0_DH + 16 * c, o);
}
// dQKV of the (sample, head): dV[j] = p_j dO0, dK[j] = ds_j q0, dQ[0] = sum_j ds_j K[j], dQ[j>0] = 0
template <typename A>
__global__ void __launch_bounds__(128) attention_row0_bwd_kernel(const A* __restrict__ QKV, const A* __restrict__ dO,
                                                                 A* __restrict__ dQKV, int items, int N, int H, float scale) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (it >= items) return;
  const int b = it / H, h = it % H, inner = H * R0_DH, ld = 3 * inner;
  const A* base = QKV + (int64_t)b * N * ld + h * R0_DH;
  A* obase = dQKV + (int64_t)b * N * ld + h * R0_DH;
  float q[16], go[16], p[R0_NP], ds[R0_NP];
  r0_ld16<A>(base + 16 * c, q);
  r0_ld16<A>(dO + (int64_t)b * N * inner + h * R0_DH + 16 * c, go);
  r0_dots<A>(q, base + inner + 16 * c, ld, N, g, p);
  const float inv = r0_softmax(p, N, g, scale);
  r0_dots<A>(go, base + 2 * inner + 16 * c, ld, N, g, ds);          // dp_j = <dO0, V[j]>
  float delta = 0.f;
#pragma unroll
  for (int t = 0; t < R0_NP; ++t) { p[t] *= inv; delta = fmaf(p[t], ds[t], delta); }
  delta = r0_groups_sum(delta);
  float dq[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dq[i] = 0.f;
#pragma unroll
  for (int t = 0; t < R0_NP; ++t) {
    const int j = 8 * t + g;
    const float dsj = p[t] * (ds[t] - delta) * scale;
    if (j < N) {
      float k[16], w[16];
      r0_ld16<A>(base + inner + (int64_t)j * ld + 16 * c, k);
#pragma unroll
      for (int i = 0; i < 16; ++i) { dq[i] = fmaf(dsj, k[i], dq[i]); w[i] = dsj * q[i]; }
      r0_st16<A>(obase + inner + (int64_t)j * ld + 16 * c, w);       // dK[j]
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = p[t] * go[i];
      r0_st16<A>(obase + 2 * inner + (int64_t)j * ld + 16 * c, w);   // dV[j]
      if (j > 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0.f;
        r0_st16<A>(obase + (int64_t)j * ld + 16 * c, w);             // dQ[j] = 0
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) dq[i] = r0_groups_sum(dq[i]);
  if (g == 0) r0_st16<A>(obase + 16 * c, dq);                        // dQ[0]
}

// =====================================================================================
// Attention (vn/GoalFormer.py:71-81): softmax(q k^T * dh^-0.5) v per (sample, head).
// CUDA-core version: K,V (and Q,dO in backward) of one (b,h) staged in shared memory,
// one warp per query row, warp-shuffle softmax.
// QKV layout: [T, 3*inner], column = which*inner + h*dh + d  (chunk(3) then 'b n (h d)').
// =====================================================================================
template <typename A>
__global__ void attention_fwd_kernel(const A* __restrict__ QKV, A* __restrict__ O, int N, int H, int dh,
                                     float scale) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int inner = H * dh, ld = 3 * inner, kst = dh + 1;
  float* Ks = sm;                 // [N][dh+1]
  float* Vs = Ks + N * kst;       // [N][dh]
  float* Ps = Vs + N * dh;        // [warps][N]
  float* Qs = Ps + (blockDim.x >> 5) * N;  // [warps][dh]
  const A* base = QKV + (int64_t)b * N * ld + h * dh;
  for (int i = threadIdx.x; i < N * dh; i += blockDim.x) {
    const int j = i / dh, d = i % dh;
    Ks[j * kst + d] = ldf(base + (int64_t)j * ld + inner + d);
    Vs[j * dh + d] = ldf(base + (int64_t)j * ld + 2 * inner + d);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* P = Ps + warp * N;
  float* Q = Qs + warp * dh;
  for (int i = warp; i < N; i += nw) {
    for (int d = lane; d < dh; d += 32) Q[d] = ldf(base + (int64_t)i * ld + d);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      float s = 0.f;
      for (int d = 0; d < dh; ++d) s = fmaf(Q[d], Ks[j * kst + d], s);
      s *= scale;
      P[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) { const float e = expf(P[j] - mx); P[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float o = 0.f;
      for (int j = 0; j < N; ++j) o = fmaf(P[j], Vs[j * dh + d], o);
      stf(O + ((int64_t)b * N + i) * inner + h * dh + d, o * inv);
    }
    __syncwarp();
  }
}

// Backward: pass 1 (warp per query row) -> lse, delta, dQ ; pass 2 (warp per key row) -> dK, dV.
// K and V always live in shared memory; Q and dO too when they fit (QDO_SMEM), otherwise (long
// sequences, e.g. 257 tokens at 2x resolution) they are read through L1/L2.
template <typename A, bool QDO_SMEM>
__global__ void attention_bwd_kernel(const A* __restrict__ QKV, const A* __restrict__ O,
                                     const A* __restrict__ dO, A* __restrict__ dQKV, int N, int H, int dh,
                                     float scale) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int inner = H * dh, ld = 3 * inner, st = dh + 1, nw = blockDim.x >> 5;
  float* Ks = sm;                  // [N][dh+1]
  float* Vs = Ks + N * st;         // [N][dh+1]
  float* lse = Vs + N * st;        // [N]
  float* dlt = lse + N;            // [N]
  float* W1 = dlt + N;             // [warps][N]
  float* W2 = W1 + nw * N;         // [warps][N]
  float* Qw = W2 + nw * N;         // [warps][dh]   (pass 1: this warp's query row)
  float* dOw = Qw + nw * dh;       // [warps][dh]
  float* Qs = dOw + nw * dh;       // [N][dh+1]     (QDO_SMEM only)
  float* dOs = Qs + N * st;        // [N][dh+1]     (QDO_SMEM only)
  const A* base = QKV + (int64_t)b * N * ld + h * dh;
  const int64_t obase = (int64_t)b * N * inner + h * dh;
  for (int i = threadIdx.x; i < N * dh; i += blockDim.x) {
    const int j = i / dh, d = i % dh;
    Ks[j * st + d] = ldf(base + (int64_t)j * ld + inner + d);
    Vs[j * st + d] = ldf(base + (int64_t)j * ld + 2 * inner + d);
    if (QDO_SMEM) {
      Qs[j * st + d] = ldf(base + (int64_t)j * ld + d);
      dOs[j * st + d] = ldf(dO + obase + (int64_t)j * inner + d);
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* P = W1 + warp * N;
  float* DS = W2 + warp * N;
  float* q = Qw + warp * dh;
  float* go = dOw + warp * dh;
  // ---- pass 1: rows of the score matrix
  for (int i = warp; i < N; i += nw) {
    float dl = 0.f;
    for (int d = lane; d < dh; d += 32) {
      q[d] = ldf(base + (int64_t)i * ld + d);
      go[d] = ldf(dO + obase + (int64_t)i * inner + d);
      dl = fmaf(go[d], ldf(O + obase + (int64_t)i * inner + d), dl);
    }
    dl = warp_sum(dl);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < dh; ++d) {
        s = fmaf(q[d], Ks[j * st + d], s);
        dp = fmaf(go[d], Vs[j * st + d], dp);
      }
      s *= scale;
      P[j] = s;
      DS[j] = dp;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) sum += expf(P[j] - mx);
    sum = warp_sum(sum);
    const float l = mx + logf(sum);
    for (int j = lane; j < N; j += 32) {
      const float p = expf(P[j] - l);
      DS[j] = p * (DS[j] - dl) * scale;
    }
    if (lane == 0) { lse[i] = l; dlt[i] = dl; }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float a = 0.f;
      for (int j = 0; j < N; ++j) a = fmaf(DS[j], Ks[j * st + d], a);
      stf(dQKV + ((int64_t)b * N + i) * ld + h * dh + d, a);
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- pass 2: columns of the score matrix
  for (int j = warp; j < N; j += nw) {
    for (int i = lane; i < N; i += 32) {
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < dh; ++d) {
        const float qv = QDO_SMEM ? Qs[i * st + d] : ldf(base + (int64_t)i * ld + d);
        const float gv = QDO_SMEM ? dOs[i * st + d] : ldf(dO + obase + (int64_t)i * inner + d);
        s = fmaf(qv, Ks[j * st + d], s);
        dp = fmaf(gv, Vs[j * st + d], dp);
      }
      const float p = expf(s * scale - lse[i]);
      P[i] = p;
      DS[i] = p * (dp - dlt[i]) * scale;
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float dk = 0.f, dv = 0.f;
      for (int i = 0; i < N; ++i) {
        const float qv = QDO_SMEM ? Qs[i * st + d] : ldf(base + (int64_t)i * ld + d);
        const float gv = QDO_SMEM ? dOs[i * st + d] : ldf(dO + obase + (int64_t)i * inner + d);
        dk = fmaf(DS[i], qv, dk);
        dv = fmaf(P[i], gv, dv);
      }
      stf(dQKV + ((int64_t)b * N + j) * ld + inner + h * dh + d, dk);
      stf(dQKV + ((int64_t)b * N + j) * ld + 2 * inner + h * dh + d, dv);
    }
    __syncwarp();
  }
}

// =====================================================================================
// cls pooling + RMSNorm (vn/GoalFormer.py:167-170,120-122): z = x0/max(|x0|,1e-12)*sqrt(D)*g
// =====================================================================================
__global__ void pool_rmsnorm_fwd_kernel(const float* __restrict__ X, const float* __restrict__ g,
                                        float* __restrict__ z, int B, int N, int D, float sqrtD) {
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* x = X + (int64_t)b * N * D;
  float q = 0.f;
  for (int d = lane; d < D; d += 32) q = fmaf(x[d], x[d], q);
  const float nrm = fmaxf(sqrtf(warp_sum(q)), 1e-12f);
  for (int d = lane; d < D; d += 32) z[b * D + d] = x[d] / nrm * sqrtD * g[d];
}

// dX (full [B,N,D], zero except token 0) and per-sample dg contributions
__global__ void pool_rmsnorm_bwd_kernel(const float* __restrict__ X, const float* __restrict__ g,
                                        const float* __restrict__ dz, float* __restrict__ dX,
                                        bf16* __restrict__ dX_lp, float* __restrict__ dg_rows, int B, int N,
                                        int D, float sqrtD) {
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.x;
  const float* x = X + (int64_t)b * N * D;
  float* dx = dX + (int64_t)b * N * D;
  __shared__ float red[2];
  // zero rows 1..N-1
  bf16* dxl = dX_lp ? dX_lp + (int64_t)b * N * D : nullptr;
  for (int i = D + threadIdx.x; i < N * D; i += blockDim.x) {
    dx[i] = 0.f;
    if (dxl) dxl[i] = __float2bfloat16_rn(0.f);
  }
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float q = 0.f, dot = 0.f;
    for (int d = lane; d < D; d += 32) q = fmaf(x[d], x[d], q);
    const float nr = sqrtf(warp_sum(q));
    const float nrm = fmaxf(nr, 1e-12f);
    for (int d = lane; d < D; d += 32) dot = fmaf(dz[b * D + d] * g[d], x[d], dot);
    dot = warp_sum(dot);
    if (lane == 0) { red[0] = nrm; red[1] = (nr > 1e-12f) ? dot / (nrm * nrm) : 0.f; }
  }
  __syncthreads();
  const float nrm = red[0], proj = red[1];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float w = dz[b * D + d] * g[d];
    const float nv = sqrtD / nrm * (w - x[d] * proj);
    dx[d] = nv;
    if (dxl) dxl[d] = __float2bfloat16_rn(nv);
    dg_rows[b * D + d] = dz[b * D + d] * x[d] / nrm * sqrtD;
  }
}

// dX[b, n, :] = (n == 0) ? dXc[b, :] : 0     (last block: only token 0 carries gradient)
__global__ void scatter_row0_kernel(const float* __restrict__ dXc, float* __restrict__ dX, int64_t total, int N,
                                    int D) {
  pdl_wait();
  pdl_launch();
  // one thread = 4 consecutive features (D % 4 == 0)
  const int64_t total4 = total >> 2;
  const int D4 = D >> 2;
  for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < total4; i4 += (int64_t)gridDim.x * blockDim.x) {
    int64_t bn;
    int d4;
    if (total4 < ((int64_t)1 << 31)) { const unsigned u = (unsigned)i4, q = u / D4; d4 = u - q * D4; bn = q; }
    else { d4 = (int)(i4 % D4); bn = i4 / D4; }
    const bool row0 = total4 < ((int64_t)1 << 31) ? ((unsigned)bn % (unsigned)N == 0) : (bn % N == 0);
    const int64_t b = total4 < ((int64_t)1 << 31) ? (int64_t)((unsigned)bn / (unsigned)N) : bn / N;
    reinterpret_cast<float4*>(dX)[i4] =
        row0 ? *reinterpret_cast<const float4*>(dXc + b * D + 4 * d4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// =====================================================================================
// Actor head tail: tanh-Gaussian sample + log-prob (vn/got_sac_network.py:235,238-251)
// =====================================================================================
struct SampleArgs {
  const float* mean;      // [B,na]
  const float* lstd_raw;  // [B,na] unclamped
  const float* eps;       // [B,na] or null
  const float* scale;     // [na]
  const float* bias;      // [na]
  const uint64_t* rng;    // when eps == null
  uint32_t stream_id;
  int64_t sample_offset;
  float *mean_out, *log_std, *action, *log_prob, *mean_t, *eps_out;
  int B, na;
  uint64_t* rng_bump;     // non-null: advance the RNG counter once this (last) kernel of the call has made its draws
};
__global__ void actor_sample_kernel(SampleArgs a) {
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  float lp = 0.f;
  for (int j = 0; j < a.na; ++j) {
    const int i = b * a.na + j;
    const float mu = a.mean[i];
    const float ls = fminf(fmaxf(a.lstd_raw[i], -20.f), 2.f);
    if (a.log_std) a.log_std[i] = ls;
    if (a.mean_out) a.mean_out[i] = mu;
    float e;
    if (a.eps) {
      e = a.eps[i];
    } else {
      uint32_t r[4];
      const uint64_t gi = (uint64_t)(a.sample_offset + b) * a.na + j;
      philox4x32(a.rng[0], gi, ((uint64_t)a.stream_id << 32) | (a.rng[1] & 0xffffffffu), r);
      e = sqrtf(-2.0f * logf(u01(r[0]))) * cospif(2.0f * u01(r[1]));
    }
    if (a.eps_out) a.eps_out[i] = e;
    const float sd = expf(ls);
    const float x = mu + sd * e;
    const float y = tanhf(x);
    const float sc = a.scale[j];
    if (a.action) a.action[i] = y * sc + a.bias[j];
    if (a.mean_t) a.mean_t[i] = tanhf(mu) * sc + a.bias[j];
    const float dx = x - mu;
    lp += -(dx * dx) / (2.0f * sd * sd) - ls - 0.91893853320467274178f;
    lp -= logf(sc * (1.0f - y * y) + 1e-6f);
  }
  if (a.log_prob) a.log_prob[b] = lp;
  // (single-block launches only — the batch-1 act loop: every draw of this call has been made when thread 0 gets here)
  if (a.rng_bump && blockIdx.x == 0 && gridDim.x == 1) {
    __syncthreads();
    if (threadIdx.x == 0) a.rng_bump[1] += 1;
  }
}

struct SampleBwdArgs {
  const float *mean, *lstd_raw, *eps, *scale;
  const float *d_mean, *d_log_std, *d_action, *d_log_prob, *d_mean_t;  // any may be null
  float d_log_prob_const;          // host constant added to every d_log_prob
  const float* d_log_prob_dev;     // optional device scalar (alpha) ...
  float d_log_prob_dev_scale;      // ... times this (1/B_global)
  float *d_mean_out, *d_lstd_out;  // [B,na] each: d/d mean, d/d raw log_std
  int B, na;
};
__global__ void actor_sample_bwd_kernel(SampleBwdArgs a) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.na) return;
  const int b = i / a.na, j = i % a.na;
  const float mu = a.mean[i], raw = a.lstd_raw[i];
  const float ls = fminf(fmaxf(raw, -20.f), 2.f);
  const float sd = expf(ls), e = a.eps[i], sc = a.scale[j];
  const float y = tanhf(mu + sd * e);
  const float omy = 1.0f - y * y;
  const float glp = (a.d_log_prob ? a.d_log_prob[b] : 0.f) + a.d_log_prob_const +
                    (a.d_log_prob_dev ? (*a.d_log_prob_dev) * a.d_log_prob_dev_scale : 0.f);
  // d/dx of  action (= y*sc+bias) and of  -log(sc*(1-y^2)+1e-6); the Normal.log_prob terms in
  // mean cancel analytically (d/dmu = 0) and leave -1/sd in d/dsd.
  float gx = glp * (2.0f * sc * y * omy / (sc * omy + 1e-6f));
  if (a.d_action) gx += a.d_action[i] * sc * omy;
  float gmu = gx;
  float gsd = gx * e - glp / sd;
  if (a.d_mean) gmu += a.d_mean[i];
  if (a.d_mean_t) { const float t = tanhf(mu); gmu += a.d_mean_t[i] * sc * (1.0f - t * t); }
  float gls = gsd * sd;
  if (a.d_log_std) gls += a.d_log_std[i];
  if (!(raw >= -20.f && raw <= 2.f)) gls = 0.f;  // clamp backward
  a.d_mean_out[i] = gmu;
  a.d_lstd_out[i] = gls;
}

// g *= (h > 0)
__global__ void relu_mask_kernel(float* __restrict__ g, const float* __restrict__ h, int64_t n) {
  pdl_wait();
  pdl_launch();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(h[i] > 0.f)) g[i] = 0.f;
}

// cat([z, a]) (vn/got_sac_network.py:114) and its inverse for gradients
__global__ void concat_za_kernel(const float* __restrict__ z, const float* __restrict__ a,
                                 float* __restrict__ out, int B, int D, int na) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = D + na;
  if (i >= B * W) return;
  const int b = i / W, c = i % W;
  out[i] = c < D ? z[b * D + c] : a[b * na + (c - D)];
}
// dxcat = dx1 + dx2 ; split into dz and da
__global__ void split_dza_kernel(const float* __restrict__ dx1, const float* __restrict__ dx2,
                                 float* __restrict__ dz, float* __restrict__ da, int B, int D, int na) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = D + na;
  if (i >= B * W) return;
  const int b = i / W, c = i % W;
  const float v = dx1[i] + dx2[i];
  if (c < D) { if (dz) dz[b * D + c] = v; }
  else if (da) da[b * na + (c - D)] = v;
}

// =====================================================================================
// SAC losses (vn/DRL.py:388-399, 404-410, 416-424).  Single block, fixed-order reductions.
// =====================================================================================
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) { for (int w = 0; w < nw; ++w) s += red[w]; red[0] = s; }
  __syncthreads();
  s = red[0];
  return s;
}

// next_q = r + gamma*(min(q1t,q2t) - alpha*logp')   [B,na]   (no (1-done), DRL.py:393)
// then qf losses and their gradients wrt q1,q2 (mse mean over B_global*na elements).
__global__ void critic_loss_kernel(const float* __restrict__ q1, const float* __restrict__ q2,
                                   const float* __restrict__ q1t, const float* __restrict__ q2t,
                                   const float* __restrict__ logp2, const float* __restrict__ rew,
                                   const float* __restrict__ alpha, float gamma, int B, int na,
                                   int Bglobal, float* __restrict__ nq_out, float* __restrict__ dq1,
                                   float* __restrict__ dq2, float* __restrict__ losses) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const float al = *alpha;
  const float inv = 1.0f / ((float)Bglobal * na);
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < B * na; i += blockDim.x) {
    const int b = i / na;
    const float nq = rew[b] + gamma * (fminf(q1t[i], q2t[i]) - al * logp2[b]);
    if (nq_out) nq_out[i] = nq;
    const float e1 = q1[i] - nq, e2 = q2[i] - nq;
    s1 = fmaf(e1, e1, s1);
    s2 = fmaf(e2, e2, s2);
    dq1[i] = 2.0f * e1 * inv;
    dq2[i] = 2.0f * e2 * inv;
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) { losses[0] = s1 * inv; losses[2] = s2 * inv; }
}

// policy_loss = mean_{b,j}(alpha*log_pi[b] - min(q1pi,q2pi)[b,j]); alpha_loss and its gradient.
// galpha[0] receives d alpha_loss / d log_alpha (local shard share, SUM-reducible).
__global__ void policy_loss_kernel(const float* __restrict__ q1p, const float* __restrict__ q2p,
                                   const float* __restrict__ logpi, const float* __restrict__ alpha,
                                   const float* __restrict__ log_alpha, float target_entropy, int B, int na,
                                   int Bglobal, float* __restrict__ dq1, float* __restrict__ dq2,
                                   float* __restrict__ losses, float* __restrict__ galpha) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const float al = *alpha;
  const float inv = 1.0f / ((float)Bglobal * na);
  float sp = 0.f, sa = 0.f;
  for (int i = threadIdx.x; i < B * na; i += blockDim.x) {
    const int b = i / na;
    const float a1 = q1p[i], a2 = q2p[i];
    sp += al * logpi[b] - fminf(a1, a2);
    // torch.min(a,b) backward: ties split the gradient equally
    const float w1 = a1 < a2 ? 1.f : (a1 == a2 ? 0.5f : 0.f);
    dq1[i] = -inv * w1;
    dq2[i] = -inv * (1.f - w1);
    if (i % na == 0) sa += logpi[b] + target_entropy;
  }
  sp = block_sum(sp, red);
  sa = block_sum(sa, red);
  if (threadIdx.x == 0) {
    losses[1] = sp * inv;
    const float g = -sa / (float)Bglobal;      // d/dlog_alpha of -(log_alpha*(log_pi+H)).mean()
    losses[3] = (*log_alpha) * g;              // alpha_loss value (local share)
    galpha[0] = g;
  }
}

// learn_guidence (vn/DRL.py:257-278): per-row gradients of the actor pass over B policy rows followed by
// imitation rows.  rows < B: d log_pi = alpha / B_global, d mean_t = 0 (d action comes from the critic heads);
// rows >= B: d action = d log_pi = 0, d mean_t = 2 w_r (mean_t - target).  The imitation loss is added to losses[1].
__global__ void imitation_grad_kernel(const float* __restrict__ mean_t, const float* __restrict__ target,
                                      const float* __restrict__ weight, const float* __restrict__ alpha,
                                      float inv_bglobal, int B, int Ba, int na, float* __restrict__ dpi,
                                      float* __restrict__ dlogp, float* __restrict__ dmeant,
                                      float* __restrict__ losses) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const float al = *alpha;
  float acc = 0.f;
  for (int i = threadIdx.x; i < Ba * na; i += blockDim.x) {
    const int r = i / na;
    if (r < B) {
      dmeant[i] = 0.f;
      if (i % na == 0) dlogp[r] = al * inv_bglobal;
    } else {
      const float w = weight[r - B];
      const float e = mean_t[i] - target[(r - B) * na + i % na];
      acc = fmaf(w * e, e, acc);
      dmeant[i] = 2.0f * w * e;
      dpi[i] = 0.f;
      if (i % na == 0) dlogp[r] = 0.f;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) losses[1] += acc;
}

// Adam on the scalar log_alpha, then alpha = exp(log_alpha)  (DRL.py:419-423)
__global__ void alpha_step_kernel(float* log_alpha, float* alpha, float* m, float* v, int64_t* step,
                                  const float* g, float lr, float b1, float b2, float omb1, float omb2,
                                  float eps) {
  pdl_wait();
  pdl_launch();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int64_t t = ++(*step);
  const float gr = *g;
  *m = *m + omb1 * (gr - *m);
  *v = *v * b2 + omb2 * gr * gr;
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const double bc2 = 1.0 - pow((double)b2, (double)t);
  const float denom = sqrtf(*v) / (float)sqrt(bc2) + eps;
  *log_alpha = *log_alpha - (float)((double)lr / bc1) * (*m / denom);
  *alpha = expf(*log_alpha);
}

// =====================================================================================
// Adam (torch.optim.Adam defaults; vn/DRL.py:113,150) fused with the Polyak target update
// (vn/utils.py:31-33) and the bf16 shadow refresh.  One pass over the flat arenas.
// =====================================================================================
struct AdamArgs {
  float* p; const float* g; float* m; float* v; bf16* shadow;
  float* tgt; bf16* tgt_shadow; float tau;   // polyak target (may be null)
  int64_t n; int64_t* step;                  // device step counter (incremented by the bump kernel)
  float lr, b1, b2, eps;
  float omb1, omb2;                          // (float)(1 - beta) computed in double like torch
  int n_skip; int64_t skip_b[4], skip_e[4];
  const float* gscale;                       // optional device scalar multiplied into every gradient (clip_grad_norm_)
};
__global__ void step_bump_kernel(int64_t* step) {
  pdl_wait();
  pdl_launch(); if (threadIdx.x == 0 && blockIdx.x == 0) ++(*step); }

// 16-bit shadows of a parameter arena: bf16 copy at shadow[0, n), f16 copy at shadow[n, 2n) (operands of the tensor-core
// contractions; the f16 copy feeds the f16 x f16 second GEMM of the fused MLP forward).  Four elements per thread.
__device__ __forceinline__ void store_shadows4(bf16* shadow, int64_t n, int64_t i, const float4& p) {
  const __nv_bfloat162 b0 = __floats2bfloat162_rn(p.x, p.y), b1 = __floats2bfloat162_rn(p.z, p.w);
  const __half2 h0 = __floats2half2_rn(p.x, p.y), h1 = __floats2half2_rn(p.z, p.w);
  *reinterpret_cast<uint2*>(shadow + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&b0), *reinterpret_cast<const uint32_t*>(&b1));
  *reinterpret_cast<uint2*>(shadow + n + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
}

// ---- data parallel: gradient all-reduce fused into the optimizer pass (no NCCL call, no extra launch) -----------------
// Every rank's gradient arena lives in symmetric memory (torch.distributed._symmetric_memory): the kernel first runs a
// flag barrier over NVLink (each rank stores the Adam step number into its slot of every peer's signal pad and waits until
// all slots of its own pad show that number: all gradients of this step are written), then reads the SUM over ranks of
// every gradient element — one multimem.ld_reduce per 16 bytes, reduced inside the NVSwitch (NVLS), or, without multicast
// support, one peer load per rank in fixed rank order — and applies Adam.  All ranks read the same sums, so the replicas
// stay identical.  When a rank has read everything it raises its "done" slot on every peer; the next backward pass waits
// for those slots (dp_wait_done_kernel) before it overwrites the gradients.
struct DpDev {
  int world, rank;
  const float* mc;                  // multicast address of this arena (null: peer loads)
  const float* const* peers;        // [world] this arena on every rank (device array of device pointers)
  int64_t arena_off;                // float offset of this arena inside the symmetric buffer (peers[] point at the buffer)
  uint32_t* const* pads;            // [world] signal pads (uint32 words); slots: [which][0: ready, 1: done][rank]
  int which;                        // 0 critic, 1 actor
  float* tail_out; int64_t tail_begin, tail_n;   // reduced copy of a skipped range Adam does not touch (alpha gradient + loss sums)
  float* reduced_out;               // optional: the reduced gradients (tests)
  int* error_flag;                  // set when a flag wait times out
};
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t* dp_slot(const DpDev& dp, int on_rank, int kind, int of_rank) {
  return dp.pads[on_rank] + ((dp.which * 2 + kind) * dp.world + of_rank);
}
// all `world` slots of kind `kind` on this rank's pad have reached `epoch` (bounded spin: a dead peer must not hang the GPU)
__device__ __forceinline__ void dp_wait_all(const DpDev& dp, int kind, uint32_t epoch) {
  const long long t0 = clock64();
  for (int p = 0; p < dp.world; ++p) {
    const uint32_t* slot = dp_slot(dp, dp.rank, kind, p);
    while ((int32_t)(ld_acquire_sys(slot) - epoch) < 0) {
      if (clock64() - t0 > 6000000000ll) { if (dp.error_flag) *dp.error_flag = 1; return; }
      __nanosleep(64);
    }
  }
}
__device__ __forceinline__ float4 dp_reduced4(const DpDev& dp, int64_t i) {
  float4 g;
  if (dp.mc) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w) : "l"(dp.mc + i) : "memory");
  } else {
    g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < dp.world; ++p) {
      float4 v;
      asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(dp.peers[p] + dp.arena_off + i) : "memory");
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
    }
  }
  return g;
}
// the next writer of a gradient arena waits until every peer has finished reading the previous step's gradients
__global__ void dp_wait_done_kernel(const __grid_constant__ DpDev dp, const int64_t* step) {
  pdl_wait();
  pdl_launch();
  if (threadIdx.x == 0) dp_wait_all(dp, 1, (uint32_t)(*step));
}

// n is a multiple of 64 and every skip range starts / ends on a 64-float boundary (dgvit_param_layout), so a float4
// never straddles a range
template <bool DP>
__global__ void __launch_bounds__(256) adam_polyak_kernel(const __grid_constant__ AdamArgs a, const __grid_constant__ DpDev dp,
                                                          unsigned int* finished) {
  pdl_wait();
  pdl_launch();
  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    const double t = (double)(*a.step);
    const double bc1 = 1.0 - pow((double)a.b1, t);
    const double bc2 = 1.0 - pow((double)a.b2, t);
    sh[0] = (float)((double)a.lr / bc1);
    sh[1] = (float)sqrt(bc2);
    if (DP) {
      const uint32_t epoch = (uint32_t)(*a.step);
      if (blockIdx.x == 0) {                 // this rank's gradients are complete (stream order): tell every peer
        __threadfence_system();
        for (int p = 0; p < dp.world; ++p) st_release_sys(dp_slot(dp, p, 0, dp.rank), epoch);
      }
      dp_wait_all(dp, 0, epoch);
    }
  }
  __syncthreads();
  const float step_size = sh[0], bc2s = sh[1];
  const float gs = a.gscale ? __ldcg(a.gscale) : 1.0f;       // written by the launch right before this one
  const int64_t n4 = a.n >> 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q << 2;
    bool skip = false;
    for (int k = 0; k < a.n_skip; ++k) skip |= (i >= a.skip_b[k] && i < a.skip_e[k]);
    float4 p = *reinterpret_cast<const float4*>(a.p + i);
    if (DP && skip && i >= dp.tail_begin && i < dp.tail_begin + dp.tail_n)
      *reinterpret_cast<float4*>(dp.tail_out + (i - dp.tail_begin)) = dp_reduced4(dp, i);
    if (!skip) {
      float4 g = DP ? dp_reduced4(dp, i) : *reinterpret_cast<const float4*>(a.g + i);
      if (a.gscale) { g.x *= gs; g.y *= gs; g.z *= gs; g.w *= gs; }
      if (DP && dp.reduced_out) *reinterpret_cast<float4*>(dp.reduced_out + i) = g;
      float4 m = *reinterpret_cast<const float4*>(a.m + i), v = *reinterpret_cast<const float4*>(a.v + i);
#define DG_ADAM1(C_)                                                                                     \
      m.C_ = m.C_ + a.omb1 * (g.C_ - m.C_);              /* exp_avg.lerp_(grad, 1-beta1) */                 \
      v.C_ = v.C_ * a.b2 + a.omb2 * g.C_ * g.C_;         /* mul_(beta2).addcmul_(g,g,1-beta2) */            \
      p.C_ = p.C_ - step_size * (m.C_ / (sqrtf(v.C_) / bc2s + a.eps));
      DG_ADAM1(x) DG_ADAM1(y) DG_ADAM1(z) DG_ADAM1(w)
#undef DG_ADAM1
      *reinterpret_cast<float4*>(a.m + i) = m;
      *reinterpret_cast<float4*>(a.v + i) = v;
      *reinterpret_cast<float4*>(a.p + i) = p;
    }
    if (a.shadow) store_shadows4(a.shadow, a.n, i, p);
    if (a.tgt) {
      float4 t = *reinterpret_cast<const float4*>(a.tgt + i);
      t.x = t.x * (1.0f - a.tau) + p.x * a.tau; t.y = t.y * (1.0f - a.tau) + p.y * a.tau;
      t.z = t.z * (1.0f - a.tau) + p.z * a.tau; t.w = t.w * (1.0f - a.tau) + p.w * a.tau;
      *reinterpret_cast<float4*>(a.tgt + i) = t;
      if (a.tgt_shadow) store_shadows4(a.tgt_shadow, a.n, i, t);
    }
  }
  if (DP) {     // the last block to finish reading raises this rank's "done" slot on every peer
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int done = atomicAdd(finished, 1u);
      if (done == gridDim.x - 1) {
        *finished = 0;
        const uint32_t epoch = (uint32_t)(*a.step);
        for (int p = 0; p < dp.world; ++p) st_release_sys(dp_slot(dp, p, 1, dp.rank), epoch);
      }
    }
  }
}

// ------------------------------------------------------------------ behaviour-cloning step (vn/attention_imitating.py:58-64)
// loss = sqrt(mean((clip(tanh-mean, -max_action, max_action) - action)^2)) over B * na elements; d_mean_t = dloss / d tanh-mean.
// One block (B * na is a few hundred); fixed summation order.
__global__ void __launch_bounds__(1024) bc_loss_kernel(const float* __restrict__ mean_t, const float* __restrict__ target, int n,
                                                       float max_action, float* __restrict__ d_mean_t, float* __restrict__ loss) {
  pdl_wait();
  pdl_launch();
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = fminf(fmaxf(mean_t[i], -max_action), max_action) - target[i];
    s = fmaf(d, d, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) sh[0] = sqrtf(t / (float)n);
  }
  __syncthreads();
  const float l = sh[0];
  if (threadIdx.x == 0) *loss = l;
  const float inv = l > 0.f ? 1.0f / ((float)n * l) : 0.f;      // d sqrt(mean d^2) / d x = d / (n * loss)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float m = mean_t[i];
    d_mean_t[i] = (m >= -max_action && m <= max_action) ? (m - target[i]) * inv : 0.f;      // clip passes the gradient inside its range
  }
}
// torch.nn.utils.clip_grad_norm_: total_norm = ||g||_2 over every parameter that has a gradient (the ranges Adam skips have
// none); gradients are multiplied by min(1, max_norm / (total_norm + 1e-6)).  Two launches, fixed summation order:
// per-block sums of squares, then one block -> *scale (read by the Adam pass), *norm.
constexpr int GN_BLOCKS = 148 * 2;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const __grid_constant__ AdamArgs a, float* __restrict__ part) {
  pdl_wait();
  pdl_launch();
  __shared__ float sh[8];
  float s0 = 0.f, s1 = 0.f;
  const int64_t n4 = a.n >> 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q << 2;
    bool skip = false;
    for (int k = 0; k < a.n_skip; ++k) skip |= (i >= a.skip_b[k] && i < a.skip_e[k]);
    if (skip) continue;
    const float4 g = *reinterpret_cast<const float4*>(a.g + i);
    s0 = fmaf(g.x, g.x, fmaf(g.y, g.y, s0));
    s1 = fmaf(g.z, g.z, fmaf(g.w, g.w, s1));
  }
  float s = s0 + s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    part[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(32) clip_scale_kernel(const float* __restrict__ part, int nb, float max_norm, float* __restrict__ scale,
                                                        float* __restrict__ norm) {
  pdl_wait();
  pdl_launch();
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 32) s += (double)part[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) {
    const float total = (float)sqrt(s);
    *scale = fminf(max_norm / (total + 1e-6f), 1.0f);
    if (norm) *norm = total;
  }
}

__global__ void __launch_bounds__(256) polyak_kernel(float* __restrict__ tgt, const float* __restrict__ src, bf16* tgt_shadow,
                                                     float tau, int64_t n) {
  pdl_wait();
  pdl_launch();
  const int64_t n4 = n >> 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q << 2;
    const float4 sv = *reinterpret_cast<const float4*>(src + i);
    float4 t = sv;
    if (tau != 1.0f) {
      const float4 tv = *reinterpret_cast<const float4*>(tgt + i);
      t.x = tv.x * (1.0f - tau) + sv.x * tau; t.y = tv.y * (1.0f - tau) + sv.y * tau;
      t.z = tv.z * (1.0f - tau) + sv.z * tau; t.w = tv.w * (1.0f - tau) + sv.w * tau;
    }
    *reinterpret_cast<float4*>(tgt + i) = t;
    if (tgt_shadow) store_shadows4(tgt_shadow, n, i, t);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float t = (tau == 1.0f) ? src[i] : tgt[i] * (1.0f - tau) + src[i] * tau;      // tail (flat arenas of any length)
    tgt[i] = t;
    if (tgt_shadow) { tgt_shadow[i] = __float2bfloat16_rn(t); reinterpret_cast<__half*>(tgt_shadow)[n + i] = __float2half_rn(t); }
  }
}

__global__ void __launch_bounds__(256) shadow_refresh_kernel(const float* __restrict__ p, bf16* __restrict__ s, int64_t n) {
  pdl_wait();
  pdl_launch();
  const int64_t n4 = n >> 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x)
    store_shadows4(s, n, q << 2, *reinterpret_cast<const float4*>(p + (q << 2)));
}

__global__ void rng_advance_kernel(uint64_t* rng) {
  pdl_wait();
  pdl_launch(); if (threadIdx.x == 0 && blockIdx.x == 0) rng[1] += 1; }

// =====================================================================================
// Replay gather (cpprb sample(); vn/DRL.py:375-386): bit-exact 16-byte vectorised row copies, ONE launch.
// grid = (slices, B, 2): z=0 -> obs[idx], z=1 -> obs[(idx+1)%size]; every thread keeps GATHER_ILP independent
// 16-byte loads in flight (the copy is a pure latency x bandwidth problem).  The block (0, b, 0) also copies the
// per-transition scalar fields of row b.  The field table is a __grid_constant__ parameter: indexed straight out of
// the constant bank (a by-value copy indexed at run time lands in local memory).
// =====================================================================================
constexpr int GATHER_ILP = 4;
struct GatherArgs {
  const float4* store; const int64_t* idx; int64_t size, frame4;
  float4* obs; float4* next_obs;
  const float* src[5]; float* dst[5]; int width[5]; int n;
};
__global__ void __launch_bounds__(256) replay_gather_kernel(const __grid_constant__ GatherArgs a) {
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.y;
  int64_t r = a.idx[b];
  if (blockIdx.x == 0 && blockIdx.z == 0) {
    // scalar fields: thread t copies element t of the concatenated fields (<= 16 floats per transition)
    int t = threadIdx.x;
    for (int k = 0; k < a.n; ++k) {
      const int w = a.width[k];
      if (t < w) { if (a.dst[k] && a.src[k]) a.dst[k][(int64_t)b * w + t] = a.src[k][r * w + t]; break; }
      t -= w;
    }
  }
  float4* dst = a.obs;
  if (blockIdx.z == 1) { r = (r + 1) % a.size; dst = a.next_obs; }
  if (!dst) return;
  const float4* src = a.store + r * a.frame4;
  dst += (int64_t)b * a.frame4;
  const int64_t per = (int64_t)blockDim.x * GATHER_ILP;
  for (int64_t base = (int64_t)blockIdx.x * per; base < a.frame4; base += (int64_t)gridDim.x * per) {
    float4 v[GATHER_ILP];
#pragma unroll
    for (int u = 0; u < GATHER_ILP; ++u) {
      const int64_t i = base + u * blockDim.x + threadIdx.x;
      if (i < a.frame4)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(src + i));
    }
#pragma unroll
    for (int u = 0; u < GATHER_ILP; ++u) {
      const int64_t i = base + u * blockDim.x + threadIdx.x;
      if (i < a.frame4) __stcs(dst + i, v[u]);
    }
  }
}

// =====================================================================================
// Replay append (store_transition / initialize_expert_buffer, vn/DRL.py:449-477): n packed transition records
// [obs frame | next_obs frame | pobs | next_pobs | act | rew | done | engage] (floats; the record may live in pinned
// host memory: zero-copy) scattered into the ring store in ONE launch: obs -> slot[i], next_obs -> (slot[i]+1) % size
// (cpprb next_of="obs").  grid = (slices, n, 2).
// =====================================================================================
struct AppendArgs {
  float4* store; int64_t size, frame4;
  float* fld[6]; int width[6];              // pobs, next_pobs, act, rew, done, engage
  const float* rec; int64_t rec_floats;     // record pitch
  const int64_t* slot;                      // [n] (device or pinned host)
};
__global__ void __launch_bounds__(256) replay_append_kernel(const __grid_constant__ AppendArgs a) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.y;
  const int64_t s = a.slot[i];
  const float* rec = a.rec + (int64_t)i * a.rec_floats;
  if (blockIdx.x == 0 && blockIdx.z == 0) {
    int t = threadIdx.x;
    const float* f = rec + 8 * a.frame4;
    for (int k = 0; k < 6; ++k) {
      const int w = a.width[k];
      if (t < w) { if (a.fld[k]) a.fld[k][s * w + t] = f[t]; break; }
      t -= w; f += w;
    }
  }
  const int64_t row = blockIdx.z == 0 ? s : (s + 1) % a.size;
  // the next record of the same call overwrites that slot with its own obs (cpprb: the later add wins)
  if (blockIdx.z == 1 && i + 1 < (int)gridDim.y && a.slot[i + 1] == row) return;
  const float4* src = reinterpret_cast<const float4*>(rec) + (blockIdx.z == 0 ? 0 : a.frame4);
  float4* dst = a.store + row * a.frame4;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.frame4; j += (int64_t)gridDim.x * blockDim.x)
    dst[j] = src[j];
}

}  // namespace dgvit
