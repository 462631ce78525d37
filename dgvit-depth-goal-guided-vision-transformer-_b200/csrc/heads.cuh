// heads.cuh — fused actor / critic MLP heads (vn/got_sac_network.py:113-121, 230-234).
//
//   x[B,K0] -> h1 = relu(W1 x + b1) [H1] -> h2 = relu(W2 h1 + b2) [H2] -> out = W3 h2 + b3
//
// The heads are ~50 kFLOP per sample (0.03 % of a network pass) and purely latency-bound when run
// as separate tiny GEMMs, so each direction is one launch: weights staged once per CTA in shared
// memory (odd row pitch -> conflict-free), 8 samples per pass, fp32 throughout.
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
};

__host__ __device__ __forceinline__ int odd_pitch(int k) { return k | 1; }
__host__ __device__ __forceinline__ int al4(int n) { return (n + 3) & ~3; }

__global__ void __launch_bounds__(128) head_fwd_kernel(FwdArgs a) {
  pdl_wait();
  pdl_launch();
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
  float* h1s = xs + K0 * S;              // [H1][S]
  float* h2s = h1s + H1 * S;             // [H2][S]
  const int tid = threadIdx.x;
  for (int i = tid; i < H1 * K0; i += blockDim.x) W1s[(i / K0) * p1 + i % K0] = w.W1[i];
  for (int i = tid; i < H2 * H1; i += blockDim.x) W2s[(i / H1) * p2 + i % H1] = w.W2[i];
  for (int i = tid; i < a.NOa * H2; i += blockDim.x) W3s[(i / H2) * p3 + i % H2] = w.W3a[i];
  for (int i = tid; i < a.NOb * H2; i += blockDim.x) W3s[(a.NOa + i / H2) * p3 + i % H2] = w.W3b[i];
  for (int i = tid; i < H1; i += blockDim.x) b1s[i] = w.b1[i];
  for (int i = tid; i < H2; i += blockDim.x) b2s[i] = w.b2[i];
  for (int i = tid; i < NO; i += blockDim.x) b3s[i] = i < a.NOa ? w.b3a[i] : w.b3b[i - a.NOa];
  for (int s0 = blockIdx.x * S; s0 < a.B; s0 += gridDim.x * S) {
    __syncthreads();
    for (int i = tid; i < K0 * S; i += blockDim.x) {
      const int s = i / K0, k = i % K0, b = s0 + s;
      float v = 0.f;
      if (b < a.B) v = k < a.K1 ? a.x1[(int64_t)b * a.K1 + k] : a.x2[(int64_t)b * a.K2 + (k - a.K1)];
      xs[k * S + s] = v;
      if (a.xcat && blockIdx.y == 0 && b < a.B) a.xcat[(int64_t)b * K0 + k] = v;
    }
    __syncthreads();
    {  // layer 1: thread j < H1
      const int j = tid;
      float acc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s] = b1s[j];
      for (int k = 0; k < K0; ++k) {
        const float wv = W1s[j * p1 + k];
        const float4 xa = *reinterpret_cast<const float4*>(xs + k * S), xb = *reinterpret_cast<const float4*>(xs + k * S + 4);
        acc[0] = fmaf(wv, xa.x, acc[0]); acc[1] = fmaf(wv, xa.y, acc[1]); acc[2] = fmaf(wv, xa.z, acc[2]); acc[3] = fmaf(wv, xa.w, acc[3]);
        acc[4] = fmaf(wv, xb.x, acc[4]); acc[5] = fmaf(wv, xb.y, acc[5]); acc[6] = fmaf(wv, xb.z, acc[6]); acc[7] = fmaf(wv, xb.w, acc[7]);
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        acc[s] = fmaxf(acc[s], 0.f);
        h1s[j * S + s] = acc[s];
        if (s0 + s < a.B) w.h1[(int64_t)(s0 + s) * H1 + j] = acc[s];
      }
    }
    __syncthreads();
    if (tid < H2) {  // layer 2
      const int j = tid;
      float acc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s] = b2s[j];
      for (int k = 0; k < H1; ++k) {
        const float wv = W2s[j * p2 + k];
        const float4 xa = *reinterpret_cast<const float4*>(h1s + k * S), xb = *reinterpret_cast<const float4*>(h1s + k * S + 4);
        acc[0] = fmaf(wv, xa.x, acc[0]); acc[1] = fmaf(wv, xa.y, acc[1]); acc[2] = fmaf(wv, xa.z, acc[2]); acc[3] = fmaf(wv, xa.w, acc[3]);
        acc[4] = fmaf(wv, xb.x, acc[4]); acc[5] = fmaf(wv, xb.y, acc[5]); acc[6] = fmaf(wv, xb.z, acc[6]); acc[7] = fmaf(wv, xb.w, acc[7]);
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        acc[s] = fmaxf(acc[s], 0.f);
        h2s[j * S + s] = acc[s];
        if (s0 + s < a.B) w.h2[(int64_t)(s0 + s) * H2 + j] = acc[s];
      }
    }
    __syncthreads();
    if (tid < S * NO) {  // output layer(s): thread = (sample, output)
      const int s = tid / NO, o = tid % NO;
      float acc = b3s[o];
      for (int k = 0; k < H2; ++k) acc = fmaf(W3s[o * p3 + k], h2s[k * S + s], acc);
      if (s0 + s < a.B) {
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

__global__ void __launch_bounds__(256) head_bwd_dx_kernel(BwdArgs a) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  const BwdHead& h = a.h[blockIdx.y];
  const int K0 = a.K0, H2 = a.H2, NO = a.NOa + a.NOb;
  float* W1s = sm;                       // [H1][K0]
  float* W2s = W1s + H1 * K0;            // [H2][H1]
  float* W3s = W2s + H2 * H1;            // [NO][H2]
  float* dos = W3s + NO * H2;            // [NO][S]
  float* dh2s = dos + MAX_NO * S;        // [H2][S]
  float* dh1s = dh2s + H2 * S;           // [H1][S]
  const int tid = threadIdx.x;
  for (int i = tid; i < H1 * K0; i += blockDim.x) W1s[i] = h.W1[i];
  for (int i = tid; i < H2 * H1; i += blockDim.x) W2s[i] = h.W2[i];
  for (int i = tid; i < a.NOa * H2; i += blockDim.x) W3s[i] = h.W3a[i];
  for (int i = tid; i < a.NOb * H2; i += blockDim.x) W3s[a.NOa * H2 + i] = h.W3b[i];
  for (int s0 = blockIdx.x * S; s0 < a.B; s0 += gridDim.x * S) {
    __syncthreads();
    if (tid < NO * S) {
      const int o = tid / S, s = tid % S, b = s0 + s;
      float v = 0.f;
      if (b < a.B) v = o < a.NOa ? h.douta[(int64_t)b * a.NOa + o] : h.doutb[(int64_t)b * a.NOb + (o - a.NOa)];
      dos[o * S + s] = v;
    }
    __syncthreads();
    for (int i = tid; i < H2 * S; i += blockDim.x) {       // dh2[s][j]
      const int j = i / S, s = i % S, b = s0 + s;
      float acc = 0.f;
      for (int o = 0; o < NO; ++o) acc = fmaf(dos[o * S + s], W3s[o * H2 + j], acc);
      const bool on = b < a.B && h.h2[(int64_t)b * H2 + j] > 0.f;
      acc = on ? acc : 0.f;
      dh2s[j * S + s] = acc;
      if (b < a.B) h.dh2[(int64_t)b * H2 + j] = acc;
    }
    __syncthreads();
    for (int i = tid; i < H1 * S; i += blockDim.x) {       // dh1[s][k]: consecutive threads -> consecutive k
      const int k = i % H1, s = i / H1, b = s0 + s;
      float acc = 0.f;
      for (int j = 0; j < H2; ++j) acc = fmaf(dh2s[j * S + s], W2s[j * H1 + k], acc);
      const bool on = b < a.B && h.h1[(int64_t)b * H1 + k] > 0.f;
      acc = on ? acc : 0.f;
      dh1s[k * S + s] = acc;
      if (b < a.B) h.dh1[(int64_t)b * H1 + k] = acc;
    }
    __syncthreads();
    for (int i = tid; i < K0 * S; i += blockDim.x) {       // dx[s][d]
      const int d = i % K0, s = i / K0, b = s0 + s;
      float acc = 0.f;
      for (int k = 0; k < H1; ++k) acc = fmaf(dh1s[k * S + s], W1s[k * K0 + d], acc);
      if (b < a.B) h.dx[(int64_t)b * K0 + d] = acc;
    }
  }
}

// ---- backward, parameter gradients of up to 8 linear layers in one launch:
//      dW[n][k] = sum_s dy[s][n] x[s][k] ; db[n] = sum_s dy[s][n]   (fixed order over s: deterministic)
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
  __shared__ float dys[32][33];
  __shared__ float xs[32][33];
  int ji = 0;
  while (ji + 1 < a.njobs && (int)blockIdx.x >= a.job[ji + 1].tile0) ++ji;
  const DwJob& j = a.job[ji];
  const int t = blockIdx.x - j.tile0;
  const int kt = (j.K + 31) / 32;
  const int n0 = (t / kt) * 32, k0 = (t % kt) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // thread: k = tx, n = ty + 8*i
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s0 = 0; s0 < a.B; s0 += 32) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = ty + 8 * i, b = s0 + s;
      dys[s][tx] = (b < a.B && n0 + tx < j.N) ? j.dy[(int64_t)b * j.ldy + n0 + tx] : 0.f;
      xs[s][tx] = (b < a.B && k0 + tx < j.K) ? j.x[(int64_t)b * j.K + k0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int s = 0; s < 32; ++s) {
      const float xv = xs[s][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = dys[s][ty + 8 * i];
        acc[i] = fmaf(d, xv, acc[i]);
        accb[i] += d;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + 8 * i, k = k0 + tx;
    if (n < j.N && k < j.K) j.dW[(int64_t)n * j.K + k] = acc[i];
    if (j.db && k0 == 0 && tx == 0 && n < j.N) j.db[n] = accb[i];
  }
}

static size_t fwd_smem(int K0, int H2, int NO) {
  return sizeof(float) * ((size_t)al4(H1 * (K0 | 1)) + (size_t)al4(H2 * (H1 | 1)) + (size_t)al4(NO * (H2 | 1)) + H1 + al4(H2) +
                          MAX_NO + (size_t)K0 * S + (size_t)H1 * S + (size_t)H2 * S);
}
static size_t bwd_smem(int K0, int H2, int NO) {
  return sizeof(float) * ((size_t)H1 * K0 + (size_t)H2 * H1 + (size_t)NO * H2 + MAX_NO * S + (size_t)H2 * S + (size_t)H1 * S);
}

static void launch_fwd(const FwdArgs& a, cudaStream_t st) {
  const int K0 = a.K1 + a.K2, NO = a.NOa + a.NOb;
  DG_REQUIRE(NO <= MAX_NO && a.H2 <= 128 && S * NO <= 128, "head_fwd: unsupported head shape");
  const size_t smem = fwd_smem(K0, a.H2, NO);
  DG_REQUIRE(smem <= 227 * 1024, "head_fwd: K0=%d needs %zu B smem", K0, smem);
  static size_t attr = 0;
  if (smem > attr) {
    DG_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = 227 * 1024;
  }
  dim3 grid((unsigned)std::min<int64_t>(cdiv(a.B, S), 148), (unsigned)a.nheads);
  launch_k(head_fwd_kernel, grid, 128, smem, st, a);
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
  launch_k(head_bwd_dx_kernel, grid, 256, smem, st, a);
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
    launch_k(head_dw_kernel, a.total_tiles, 256, 0, st, a);
    DG_LAUNCH_CHECK();
  }
};

}  // namespace heads
}  // namespace dgvit
