// heads.cuh — fused actor / critic MLP heads (vn/got_sac_network.py:113-121, 230-234).
//
//   x[B,K0] -> h1 = relu(W1 x + b1) [H1] -> h2 = relu(W2 h1 + b2) [H2] -> out = W3 h2 + b3
//
// The heads are ~50 kFLOP per sample (0.03 % of a network pass) and purely latency-bound when run
// as separate tiny GEMMs, so each direction is one launch: weights staged once per CTA in shared
// memory with float4 loads (the staging round trips, not the arithmetic, set the duration), 8 samples
// per pass, fp32 throughout.
// Twin critic heads (fc1/fc2/fc3 and fc11/fc21/fc31) run as blockIdx.y = 0/1 of one launch; the
// actor's two output layers (mean_linear, log_std_linear) are the "a" and "b" outputs of one head.
#pragma once
#include "common.cuh"

namespace dgvit {
namespace heads {

constexpr int S = 8;      // samples per pass
constexpr int H1 = 128;   // both reference heads use 128 hidden units in the first layer
constexpr int MAX_NO = 8;

struct HeadW {
  const float *W1, *b1, *W2, *b2, *W3a, *b3a, *W3b, *b3b;   // W3b/b3b optional second output layer
  float *h1, *h2, *outa, *outb;                               // saved activations + outputs
};
struct FwdArgs {
  HeadW w[2];
  int nheads, B, K1, K2, H2, NOa, NOb;
  const float *x1, *x2;     // input = [x1 | x2] (x2 optional: the critic's action)
  float* xcat;              // optional: materialised [B, K1+K2] input (needed by the dW kernel when K2 > 0)
  // optional fused cls pooling + RMSNorm (vn/GoalFormer.py:167-170,120-122): x1 = xraw / max(|xraw|, 1e-12) * sqrt(K1) * g,
  // computed here from the trunk's token-0 rows and written to z_out (= the buffer x1 points to) for the backward
  const float* xraw; const float* rms_g; float* z_out;
};

__host__ __device__ __forceinline__ int odd_pitch(int k) { return k | 1; }
__host__ __device__ __forceinline__ int al4(int n) { return (n + 3) & ~3; }

constexpr int FWD_THREADS = 512;   // 4 K-quarters x 128 output units
constexpr int BWD_THREADS = 512;

// global [rows][cols] (dense, 16-byte aligned, rows*cols % 4 == 0 not required) -> smem [rows][pitch], float4 loads
__device__ __forceinline__ void stage_matrix(float* dst, int pitch, const float* __restrict__ src, int rows, int cols,
                                             int tid, int nthreads) {
  const int n = rows * cols, n4 = n >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  for (int i = tid; i < n4; i += nthreads) {
    const float4 v = __ldg(s4 + i);
    const int e = i * 4;
    int r = e / cols, c = e - r * cols;
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dst[r * pitch + c] = vv[q];
      if (++c == cols) { c = 0; ++r; }
    }
  }
  for (int e = n4 * 4 + tid; e < n; e += nthreads) dst[(e / cols) * pitch + e % cols] = __ldg(src + e);
}

// One pass = S samples.  Layers 1 and 2: thread (j = tid & 127, q = tid >> 7) accumulates output unit j over the
// q-th quarter of the contraction for all S samples (weights conflict-free through the odd row pitch, inputs
// broadcast as float4); the four quarters are summed in a fixed order through shared memory.
__global__ void __launch_bounds__(FWD_THREADS) head_fwd_kernel(FwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const HeadW& w = a.w[blockIdx.y];
  const int K0 = a.K1 + a.K2, H2 = a.H2, NO = a.NOa + a.NOb;
  const int p1 = odd_pitch(K0), p2 = odd_pitch(H1), p3 = odd_pitch(H2);
  float* W1s = sm;                       // [H1][p1]
  float* W2s = W1s + al4(H1 * p1);       // [H2][p2]
  float* W3s = W2s + al4(H2 * p2);       // [NO][p3]
  float* b1s = W3s + al4(NO * p3);       // [H1]
  float* b2s = b1s + H1;                 // [H2]
  float* b3s = b2s + al4(H2);            // [NO]
  float* xs = b3s + MAX_NO;              // [K0][S]
  float* h1s = xs + al4(K0) * S;         // [H1][S]
  float* h2s = h1s + H1 * S;             // [H2][S]
  float* red = h2s + H1 * S;             // [4][S][H1]
  const int tid = threadIdx.x;
  stage_matrix(W1s, p1, w.W1, H1, K0, tid, FWD_THREADS);
  stage_matrix(W2s, p2, w.W2, H2, H1, tid, FWD_THREADS);
  stage_matrix(W3s, p3, w.W3a, a.NOa, H2, tid, FWD_THREADS);
  if (a.NOb) stage_matrix(W3s + a.NOa * p3, p3, w.W3b, a.NOb, H2, tid, FWD_THREADS);
  for (int i = tid; i < H1; i += FWD_THREADS) b1s[i] = w.b1[i];
  for (int i = tid; i < H2; i += FWD_THREADS) b2s[i] = w.b2[i];
  for (int i = tid; i < NO; i += FWD_THREADS) b3s[i] = i < a.NOa ? w.b3a[i] : w.b3b[i - a.NOa];
  // Everything above reads PARAMETERS only: under programmatic dependent launch it runs while the trunk's last kernel is
  // still finishing (the weight staging is most of this kernel's latency).  The launch right after a kernel that rewrites
  // parameters is never programmatic (common.cuh), so the weights are final here.  Activations only after the wait.
  pdl_wait();
  pdl_launch();
  const int j = tid & (H1 - 1), q = tid >> 7;
  for (int s0 = blockIdx.x * S; s0 < a.B; s0 += gridDim.x * S) {
    __syncthreads();
    if (a.xraw) {              // RMSNorm of the token-0 rows: warp s normalises sample s0 + s
      const int s = tid >> 5, lane = tid & 31, b = s0 + s;
      if (s < S) {
        float q = 0.f;
        if (b < a.B) for (int d = lane; d < a.K1; d += 32) { const float x = a.xraw[(int64_t)b * a.K1 + d]; q = fmaf(x, x, q); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float sc = sqrtf((float)a.K1) / fmaxf(sqrtf(q), 1e-12f);
        for (int d = lane; d < a.K1; d += 32) {
          const float v = b < a.B ? a.xraw[(int64_t)b * a.K1 + d] * sc * a.rms_g[d] : 0.f;
          xs[d * S + s] = v;
          if (b < a.B && blockIdx.y == 0) {
            a.z_out[(int64_t)b * a.K1 + d] = v;
            if (a.xcat) a.xcat[(int64_t)b * K0 + d] = v;
          }
        }
      }
    }
    for (int i = tid; i < K0 * S; i += FWD_THREADS) {
      const int s = i / K0, k = i % K0, b = s0 + s;
      if (a.xraw && k < a.K1) continue;
      float v = 0.f;
      if (b < a.B) v = k < a.K1 ? a.x1[(int64_t)b * a.K1 + k] : a.x2[(int64_t)b * a.K2 + (k - a.K1)];
      xs[k * S + s] = v;
      if (a.xcat && blockIdx.y == 0 && b < a.B) a.xcat[(int64_t)b * K0 + k] = v;
    }
    __syncthreads();
    auto layer = [&](const float* Ws, int pitch, const float* in, int K, int nout) {
      if (j < nout) {
        const int kq = (K + 3) >> 2, k0 = q * kq, k1 = min(K, k0 + kq);
        float acc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = 0.f;
#pragma unroll 4
        for (int k = k0; k < k1; ++k) {
          const float wv = Ws[j * pitch + k];
          const float4 xa = *reinterpret_cast<const float4*>(in + k * S), xb = *reinterpret_cast<const float4*>(in + k * S + 4);
          acc[0] = fmaf(wv, xa.x, acc[0]); acc[1] = fmaf(wv, xa.y, acc[1]); acc[2] = fmaf(wv, xa.z, acc[2]); acc[3] = fmaf(wv, xa.w, acc[3]);
          acc[4] = fmaf(wv, xb.x, acc[4]); acc[5] = fmaf(wv, xb.y, acc[5]); acc[6] = fmaf(wv, xb.z, acc[6]); acc[7] = fmaf(wv, xb.w, acc[7]);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) red[(q * S + s) * H1 + j] = acc[s];
      }
    };
    auto finish = [&](const float* bs, float* hs, float* hg, int nout) {   // relu(bias + quarters), fixed order
      for (int i = tid; i < nout * S; i += FWD_THREADS) {
        const int s = i / nout, u = i - s * nout;
        float v = bs[u] + red[(0 * S + s) * H1 + u];
        v += red[(1 * S + s) * H1 + u]; v += red[(2 * S + s) * H1 + u]; v += red[(3 * S + s) * H1 + u];
        v = fmaxf(v, 0.f);
        hs[u * S + s] = v;
        if (s0 + s < a.B) hg[(int64_t)(s0 + s) * nout + u] = v;
      }
    };
    layer(W1s, p1, xs, K0, H1);
    __syncthreads();
    finish(b1s, h1s, w.h1, H1);
    __syncthreads();
    layer(W2s, p2, h1s, H1, H2);
    __syncthreads();
    finish(b2s, h2s, w.h2, H2);
    __syncthreads();
    if (tid < S * NO * 4) {  // output layer(s): 4 lanes per (sample, output), quarter of H2 each
      const int pr = tid >> 2, part = tid & 3;
      const int s = pr / NO, o = pr % NO;
      const int kq = (H2 + 3) >> 2, k0 = part * kq, k1 = min(H2, k0 + kq);
      float acc = 0.f;
      for (int k = k0; k < k1; ++k) acc = fmaf(W3s[o * p3 + k], h2s[k * S + s], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += b3s[o];
      if (part == 0 && s0 + s < a.B) {
        if (o < a.NOa) w.outa[(int64_t)(s0 + s) * a.NOa + o] = acc;
        else w.outb[(int64_t)(s0 + s) * a.NOb + (o - a.NOa)] = acc;
      }
    }
  }
}

// ---- backward, input-gradient chain:  dout -> dh2 -> dh1 -> dx   (weights in natural layout)
struct BwdHead {
  const float *W1, *W2, *W3a, *W3b;
  const float *h1, *h2;
  const float *douta, *doutb;   // [B,NOa], [B,NOb]
  float *dh1, *dh2;             // [B,H1], [B,H2]  (relu-masked; operands of the dW kernel)
  float *dx;                    // [B,K0]
};
struct BwdArgs {
  BwdHead h[2];
  int nheads, B, K0, H2, NOa, NOb;
};

__global__ void __launch_bounds__(BWD_THREADS) head_bwd_dx_kernel(BwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const BwdHead& h = a.h[blockIdx.y];
  const int K0 = a.K0, H2 = a.H2, NO = a.NOa + a.NOb;
  float* W1s = sm;                       // [H1][K0]
  float* W2s = W1s + al4(H1 * K0);       // [H2][H1]
  float* W3s = W2s + H2 * H1;            // [NO][H2]
  float* dos = W3s + al4(NO * H2);       // [NO][S]
  float* dh2s = dos + MAX_NO * S;        // [H2][S]
  float* dh1s = dh2s + H2 * S;           // [H1][S]
  const int tid = threadIdx.x;
  stage_matrix(W1s, K0, h.W1, H1, K0, tid, BWD_THREADS);
  stage_matrix(W2s, H1, h.W2, H2, H1, tid, BWD_THREADS);
  stage_matrix(W3s, H2, h.W3a, a.NOa, H2, tid, BWD_THREADS);
  if (a.NOb) stage_matrix(W3s + a.NOa * H2, H2, h.W3b, a.NOb, H2, tid, BWD_THREADS);
  pdl_wait();            // (parameters staged before the wait, as in the forward)
  pdl_launch();
  for (int s0 = blockIdx.x * S; s0 < a.B; s0 += gridDim.x * S) {
    __syncthreads();
    if (tid < NO * S) {
      const int o = tid / S, s = tid % S, b = s0 + s;
      float v = 0.f;
      if (b < a.B) v = o < a.NOa ? h.douta[(int64_t)b * a.NOa + o] : h.doutb[(int64_t)b * a.NOb + (o - a.NOa)];
      dos[o * S + s] = v;
    }
    __syncthreads();
    for (int i = tid; i < H2 * S; i += BWD_THREADS) {       // dh2[s][j]: consecutive threads -> consecutive j
      const int j = i % H2, s = i / H2, b = s0 + s;
      float acc = 0.f;
      for (int o = 0; o < NO; ++o) acc = fmaf(dos[o * S + s], W3s[o * H2 + j], acc);
      const bool on = b < a.B && h.h2[(int64_t)b * H2 + j] > 0.f;
      acc = on ? acc : 0.f;
      dh2s[j * S + s] = acc;
      if (b < a.B) h.dh2[(int64_t)b * H2 + j] = acc;
    }
    __syncthreads();
    for (int i = tid; i < H1 * (S / 2); i += BWD_THREADS) { // dh1[s][k], two samples per thread
      const int k = i % H1, s = (i / H1) * 2, b = s0 + s;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
      for (int j = 0; j < H2; ++j) {
        const float wv = W2s[j * H1 + k];
        const float2 d = *reinterpret_cast<const float2*>(dh2s + j * S + s);
        a0 = fmaf(d.x, wv, a0); a1 = fmaf(d.y, wv, a1);
      }
      a0 = (b < a.B && h.h1[(int64_t)b * H1 + k] > 0.f) ? a0 : 0.f;
      a1 = (b + 1 < a.B && h.h1[(int64_t)(b + 1) * H1 + k] > 0.f) ? a1 : 0.f;
      dh1s[k * S + s] = a0; dh1s[k * S + s + 1] = a1;
      if (b < a.B) h.dh1[(int64_t)b * H1 + k] = a0;
      if (b + 1 < a.B) h.dh1[(int64_t)(b + 1) * H1 + k] = a1;
    }
    __syncthreads();
    for (int i = tid; i < K0 * (S / 2); i += BWD_THREADS) { // dx[s][d], two samples per thread
      const int d = i % K0, s = (i / K0) * 2, b = s0 + s;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
      for (int k = 0; k < H1; ++k) {
        const float wv = W1s[k * K0 + d];
        const float2 g = *reinterpret_cast<const float2*>(dh1s + k * S + s);
        a0 = fmaf(g.x, wv, a0); a1 = fmaf(g.y, wv, a1);
      }
      if (b < a.B) h.dx[(int64_t)b * K0 + d] = a0;
      if (b + 1 < a.B) h.dx[(int64_t)(b + 1) * K0 + d] = a1;
    }
  }
}

// ---- backward, parameter gradients of up to 8 linear layers in one launch:
//      dW[n][k] = sum_s dy[s][n] x[s][k] ; db[n] = sum_s dy[s][n]
// One CTA per 32x32 tile of dW.  Each of the 8 warps owns every 8th slab of 32 samples and accumulates the whole
// tile for it in registers (lane = k, 32 n-accumulators; dy rows broadcast from a per-warp smem slab); the eight
// per-warp tiles are then summed in a fixed order, so the result is deterministic.
struct DwJob {
  const float *dy, *x;   // [B,N] (row pitch ldy), [B,K]
  float *dW, *db;        // [N,K], [N]
  int N, K, ldy, tile0;  // tile0 = first tile index of this job
};
struct DwArgs {
  DwJob job[8];
  int njobs, B, total_tiles;
};
__global__ void __launch_bounds__(256) head_dw_kernel(DwArgs a) {
  pdl_wait();
  pdl_launch();
  __shared__ __align__(16) float slab[8][32][32];   // per-warp dy slab [s][n]; reused for the cross-warp sum
  __shared__ float bsum[8][32];
  int ji = 0;
  while (ji + 1 < a.njobs && (int)blockIdx.x >= a.job[ji + 1].tile0) ++ji;
  const DwJob& j = a.job[ji];
  const int t = blockIdx.x - j.tile0;
  const int kt = (j.K + 31) / 32;
  const int n0 = (t / kt) * 32, k0 = (t % kt) * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool n_ok = n0 + lane < j.N, k_ok = k0 + lane < j.K;
  float acc[32], accb = 0.f;
#pragma unroll
  for (int n = 0; n < 32; ++n) acc[n] = 0.f;
  for (int s0 = w * 32; s0 < a.B; s0 += 8 * 32) {
    float xr[32];
#pragma unroll
    for (int s = 0; s < 32; ++s) {
      const int b = s0 + s;
      const float d = (b < a.B && n_ok) ? j.dy[(int64_t)b * j.ldy + n0 + lane] : 0.f;
      xr[s] = (b < a.B && k_ok) ? j.x[(int64_t)b * j.K + k0 + lane] : 0.f;
      slab[w][s][lane] = d;
      accb += d;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 32; ++s) {
      const float xv = xr[s];
#pragma unroll
      for (int n4 = 0; n4 < 8; ++n4) {
        const float4 d = *reinterpret_cast<const float4*>(&slab[w][s][4 * n4]);
        acc[4 * n4] = fmaf(d.x, xv, acc[4 * n4]); acc[4 * n4 + 1] = fmaf(d.y, xv, acc[4 * n4 + 1]);
        acc[4 * n4 + 2] = fmaf(d.z, xv, acc[4 * n4 + 2]); acc[4 * n4 + 3] = fmaf(d.w, xv, acc[4 * n4 + 3]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int n = 0; n < 32; ++n) slab[w][n][lane] = acc[n];     // now [w][n][k]
  bsum[w][lane] = accb;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = w + 8 * i;
    float v = slab[0][n][lane];
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) v += slab[ww][n][lane];
    if (n0 + n < j.N && k_ok) j.dW[(int64_t)(n0 + n) * j.K + k0 + lane] = v;
  }
  if (j.db && k0 == 0 && w == 0 && n_ok) {
    float v = bsum[0][lane];
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) v += bsum[ww][lane];
    j.db[n0 + lane] = v;
  }
}

static size_t fwd_smem(int K0, int H2, int NO) {
  return sizeof(float) * ((size_t)al4(H1 * (K0 | 1)) + (size_t)al4(H2 * (H1 | 1)) + (size_t)al4(NO * (H2 | 1)) + H1 + al4(H2) +
                          MAX_NO + (size_t)al4(K0) * S + (size_t)H1 * S + (size_t)H1 * S + (size_t)4 * S * H1);
}
static size_t bwd_smem(int K0, int H2, int NO) {
  return sizeof(float) * ((size_t)al4(H1 * K0) + (size_t)H2 * H1 + (size_t)al4(NO * H2) + MAX_NO * S + (size_t)H2 * S +
                          (size_t)H1 * S);
}

static void launch_fwd(const FwdArgs& a, cudaStream_t st) {
  const int K0 = a.K1 + a.K2, NO = a.NOa + a.NOb;
  DG_REQUIRE(NO <= MAX_NO && a.H2 <= 128 && S * NO * 4 <= FWD_THREADS, "head_fwd: unsupported head shape");
  const size_t smem = fwd_smem(K0, a.H2, NO);
  DG_REQUIRE(smem <= 227 * 1024, "head_fwd: K0=%d needs %zu B smem", K0, smem);
  static size_t attr = 0;
  if (smem > attr) {
    DG_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = 227 * 1024;
  }
  dim3 grid((unsigned)std::min<int64_t>(cdiv(a.B, S), 148), (unsigned)a.nheads);
  if (skip_mask() & SKIP_HEADS) return;
  launch_k(head_fwd_kernel, grid, FWD_THREADS, smem, st, a);
  DG_LAUNCH_CHECK();
}
static void launch_bwd_dx(const BwdArgs& a, cudaStream_t st) {
  const int NO = a.NOa + a.NOb;
  const size_t smem = bwd_smem(a.K0, a.H2, NO);
  DG_REQUIRE(smem <= 227 * 1024 && NO <= MAX_NO, "head_bwd: K0=%d needs %zu B smem", a.K0, smem);
  static size_t attr = 0;
  if (smem > attr) {
    DG_CUDA(cudaFuncSetAttribute(head_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = 227 * 1024;
  }
  dim3 grid((unsigned)std::min<int64_t>(cdiv(a.B, S), 148), (unsigned)a.nheads);
  if (skip_mask() & SKIP_HEADS) return;
  launch_k(head_bwd_dx_kernel, grid, BWD_THREADS, smem, st, a);
  DG_LAUNCH_CHECK();
}
struct DwList {
  DwArgs a;
  DwList(int B) { a.njobs = 0; a.B = B; a.total_tiles = 0; }
  void add(const float* dy, int ldy, const float* x, float* dW, float* db, int N, int K) {
    DG_REQUIRE(a.njobs < 8, "too many dW jobs");
    DwJob& j = a.job[a.njobs++];
    j.dy = dy; j.x = x; j.dW = dW; j.db = db; j.N = N; j.K = K; j.ldy = ldy; j.tile0 = a.total_tiles;
    a.total_tiles += (int)(cdiv(N, 32) * cdiv(K, 32));
  }
  void launch(cudaStream_t st) {
    if (skip_mask() & SKIP_HEADS) return;
    launch_k(head_dw_kernel, a.total_tiles, 256, 0, st, a);
    DG_LAUNCH_CHECK();
  }
};

}  // namespace heads
}  // namespace dgvit
