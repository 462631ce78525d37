// gemm_simt.cuh — strided CUDA-core GEMM with fused epilogues.
//
// This is the fp32 arithmetic path (parity mode: the reference computes every
// contraction in true fp32, SURVEY.md §7 "hard parts") and the path of the tiny head
// GEMMs.  C[m,n] = sum_k A(m,k) * B(k,n) with arbitrary element strides, so the same
// kernel serves y = x W^T (NT), dx = dy W (NN) and dW = dy^T x (TN, split-K).
#pragma once
#include "common.cuh"

namespace dgvit {

enum {
  EPI_NONE = 0,
  EPI_BIAS,        // C = acc + bias[n]
  EPI_BIAS_RELU,   // C = relu(acc + bias[n])
  EPI_BIAS_GELU2,  // C = acc + bias[n] ; C2 = gelu(C)
  EPI_BIAS_RESID,  // C(f32) = acc + (bias ? bias[n] : 0) + resid[m,n]   (resid may alias C)
  EPI_GELU_BWD,    // C = acc * gelu'(aux[m,n])
  EPI_RELU_BWD,    // C = acc * (auxf[m,n] > 0)
  EPI_GELU_BWD2    // C = acc * gelu'(aux[m,n]) ; C2 = gelu(aux[m,n])   (recomputes the activation for dW)
};

struct GemmArgs {
  int M = 0, N = 0, K = 0;
  const void* A = nullptr;
  int64_t a_sm = 0, a_sk = 0;
  const void* B = nullptr;
  int64_t b_sk = 0, b_sn = 0;
  void* C = nullptr;
  int64_t ldc = 0;
  int epi = EPI_NONE;
  int skip_pre = 0;           // EPI_BIAS_GELU2: do not store the pre-activation C (forward passes nobody differentiates)
  const float* bias = nullptr;
  const float* resid = nullptr;
  int64_t ldr = 0;
  void* C2 = nullptr;
  const void* aux = nullptr;  // typed like C for EPI_GELU_BWD, float for EPI_RELU_BWD
  int64_t ldaux = 0;
  int splitk = 1;
  float* partial = nullptr;  // [splitk, M, N]
  // optional LayerNorm of the output rows fused into the epilogue (tensor-core kernel only, N == 64, EPI_BIAS_RESID)
  const float* ln_gamma = nullptr; const float* ln_beta = nullptr;
  void* ln_out = nullptr; float* ln_mean = nullptr; float* ln_rstd = nullptr;
  struct ReduceList* defer = nullptr;   // split-K: queue the reduction instead of launching it (tensor-core kernel)
};

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_PAD = 4;

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  pdl_wait();
  pdl_launch();
  __shared__ __align__(16) float As[SG_BK][SG_BM + SG_PAD];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + SG_PAD];
  const TA* __restrict__ A = (const TA*)g.A;
  const TB* __restrict__ B = (const TB*)g.B;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  // split-K range
  const int kchunk = (int)cdiv(cdiv(g.K, g.splitk), SG_BK) * SG_BK;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(g.K, kbeg + kchunk);
  const bool a_kc = (g.a_sk == 1);
  const bool b_nc = (g.b_sn == 1);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int m, k;
      if (a_kc) { m = idx >> 4; k = idx & 15; } else { k = idx >> 6; m = idx & 63; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < kend) v = ldf(A + (int64_t)gm * g.a_sm + (int64_t)gk * g.a_sk);
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int n, k;
      if (b_nc) { k = idx >> 6; n = idx & 63; } else { n = idx >> 4; k = idx & 15; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < kend) v = ldf(B + (int64_t)gk * g.b_sk + (int64_t)gn * g.b_sn);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (g.splitk > 1) {
    float* P = g.partial + (int64_t)blockIdx.z * g.M * g.N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < g.N) P[(int64_t)m * g.N + n] = acc[i][j];
      }
    }
    return;
  }
  TC* C = (TC*)g.C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      switch (g.epi) {
        case EPI_NONE: break;
        case EPI_BIAS: v += g.bias[n]; break;
        case EPI_BIAS_RELU: v = fmaxf(v + g.bias[n], 0.f); break;
        case EPI_BIAS_GELU2:
          v += g.bias[n];
          stf((TC*)g.C2 + (int64_t)m * g.ldc + n, gelu_f(v));
          break;
        case EPI_BIAS_RESID: v += (g.bias ? g.bias[n] : 0.f) + g.resid[(int64_t)m * g.ldr + n]; break;
        case EPI_GELU_BWD: v *= gelu_grad_f(ldf((const TC*)g.aux + (int64_t)m * g.ldaux + n)); break;
        case EPI_RELU_BWD: v = ((const float*)g.aux)[(int64_t)m * g.ldaux + n] > 0.f ? v : 0.f; break;
        case EPI_GELU_BWD2: {
          const float x = ldf((const TC*)g.aux + (int64_t)m * g.ldaux + n);
          v *= gelu_grad_f(x);
          stf((TC*)g.C2 + (int64_t)m * g.ldc + n, gelu_f(x));
        } break;
      }
      if (!(g.epi == EPI_BIAS_GELU2 && g.skip_pre)) stf(C + (int64_t)m * g.ldc + n, v);
    }
  }
}

// out[i] = sum_s partial[s, i]   (fixed order: deterministic)
// out[i] = sum_k partial[k][i], k ascending (fixed order).  Four elements per thread, the split loop unrolled so
// that eight independent 16-byte loads are in flight (the kernel is pure load latency otherwise).
__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                       int S, int64_t n) {
  pdl_wait();
  pdl_launch();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n && (n & 3) == 0 && ((((uintptr_t)partial) | ((uintptr_t)out)) & 15) == 0) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (; k + 8 <= S; k += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)(k + u) * n + i));
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; k < S; ++k) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)k * n + i));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i) = s;
  } else {
    for (int64_t e = i; e < min(n, i + 4); ++e) {
      float s = 0.f;
      for (int k = 0; k < S; ++k) s += partial[(int64_t)k * n + e];
      out[e] = s;
    }
  }
}
static inline unsigned reduce_grid(int64_t n) { return (unsigned)cdiv(cdiv(n, 4), 256); }

// Several partial-sum reductions in one launch: out_j[i] = sum_k part_j[k * stride_j + i].  The backward of a
// transformer block produces five to ten of them (split-K weight gradients, LayerNorm parameter / bias partials);
// each is a few microseconds of pure latency as its own kernel.  Block = 16 warps x 128 consecutive floats of one
// job: warp w sums the partial rows k = w, w+16, ..., the 16 warp sums are added in a fixed order.
// (the partial sums are produced by the launches right before these kernels and their buffers are reused from block to block:
//  they must not be read through the non-coherent path, whose contract -- read-only for the kernel's lifetime -- a
//  programmatically launched kernel breaks: its lifetime starts while the producer still runs.  __ldcg = L2, coherent.)
struct ReduceJob {
  const float* part; float* out;
  int S, n; int64_t stride; int blk0;
};
struct ReduceJobs {
  ReduceJob job[12];
  int njobs;
};
constexpr int MR_WARPS = 16;
constexpr int MR_WIDE_S = 32;            // jobs with at most this many partial rows run in "wide" mode
// Two shapes of job.  Tall (S > 32 rows of a few hundred floats: LayerNorm / bias partials, one row per producer block):
// block = 16 warps x 128 consecutive floats, warp w sums rows w, w+16, ..., the 16 warp sums are added in a fixed order.
// Wide (S <= 32 rows of up to a quarter million floats: split-K weight gradients): with 9 rows only 9 of those 16 warps had
// a single load to do; there each thread owns one float4 column of 2048 per block and sums all S rows itself, eight loads
// in flight (round 2: 13.9 -> ~5 us per launch in the ncu list).
__global__ void __launch_bounds__(MR_WARPS * 32) multi_reduce_kernel(ReduceJobs a) {
  pdl_wait();
  pdl_launch();
  __shared__ float4 sm[MR_WARPS][32];
  int ji = 0;
  while (ji + 1 < a.njobs && (int)blockIdx.x >= a.job[ji + 1].blk0) ++ji;
  const ReduceJob& j = a.job[ji];
  if (j.S <= MR_WIDE_S) {
    const int64_t e = ((int64_t)((int)blockIdx.x - j.blk0) * (MR_WARPS * 32) + threadIdx.x) * 4;
    if (e >= j.n) return;
    const float* p = j.part + e;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (; k + 8 <= j.S; k += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(p + (int64_t)(k + u) * j.stride));
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; k < j.S; ++k) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(p + (int64_t)k * j.stride));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(j.out + e) = s;
    return;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t e = ((int64_t)((int)blockIdx.x - j.blk0) * 32 + lane) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e < j.n) {
    const float* p = j.part + e;
    int k = w;
    for (; k + 3 * MR_WARPS < j.S; k += 4 * MR_WARPS) {
      const float4 v0 = __ldcg(reinterpret_cast<const float4*>(p + (int64_t)k * j.stride));
      const float4 v1 = __ldcg(reinterpret_cast<const float4*>(p + (int64_t)(k + MR_WARPS) * j.stride));
      const float4 v2 = __ldcg(reinterpret_cast<const float4*>(p + (int64_t)(k + 2 * MR_WARPS) * j.stride));
      const float4 v3 = __ldcg(reinterpret_cast<const float4*>(p + (int64_t)(k + 3 * MR_WARPS) * j.stride));
      s.x += (v0.x + v1.x) + (v2.x + v3.x); s.y += (v0.y + v1.y) + (v2.y + v3.y);
      s.z += (v0.z + v1.z) + (v2.z + v3.z); s.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; k < j.S; k += MR_WARPS) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(p + (int64_t)k * j.stride));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  sm[w][lane] = s;
  __syncthreads();
  if (w == 0 && e < j.n) {
    float4 t = sm[0][lane];
#pragma unroll
    for (int ww = 1; ww < MR_WARPS; ++ww) { const float4 v = sm[ww][lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    *reinterpret_cast<float4*>(j.out + e) = t;
  }
}
// host side: a bump allocator over the partial-sum workspace plus the job table
struct ReduceList {
  ReduceJobs a;
  float* base; size_t cap, used; int blocks;
  ReduceList(float* b, size_t cap_) : base(b), cap(cap_), used(0), blocks(0) { a.njobs = 0; }
  float* alloc(size_t n) {
    n = (n + 3) & ~size_t(3);
    DG_REQUIRE(used + n <= cap, "reduction workspace too small: need %zu floats, have %zu", used + n, cap);
    float* p = base + used;
    used += n;
    return p;
  }
  void add(const float* part, float* out, int S, int64_t n, int64_t stride) {
    DG_REQUIRE(a.njobs < 12 && n % 4 == 0 && stride % 4 == 0 && ((((uintptr_t)part) | ((uintptr_t)out)) & 15) == 0,
               "multi_reduce: job table full or unaligned job");
    ReduceJob& j = a.job[a.njobs++];
    j.part = part; j.out = out; j.S = S; j.n = (int)n; j.stride = stride; j.blk0 = blocks;
    blocks += (int)cdiv(n, S <= MR_WIDE_S ? MR_WARPS * 32 * 4 : 128);
  }
  void launch(cudaStream_t st) {
    if (a.njobs && !(skip_mask() & SKIP_REDUCE)) {
      launch_k(multi_reduce_kernel, blocks, MR_WARPS * 32, 0, st, a);
      DG_LAUNCH_CHECK();
    }
    a.njobs = 0; blocks = 0; used = 0;
  }
};

template <typename TA, typename TB, typename TC>
static void gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return;
  DG_REQUIRE(g.splitk >= 1 && (g.splitk == 1 || (g.partial && g.epi == EPI_NONE)),
             "gemm_simt: split-K needs a partial buffer and EPI_NONE");
  dim3 grid((unsigned)cdiv(g.N, SG_BN), (unsigned)cdiv(g.M, SG_BM), (unsigned)g.splitk);
  launch_k(gemm_simt_kernel<TA, TB, TC>, grid, 256, 0, st, g);
  DG_LAUNCH_CHECK();
  if (g.splitk > 1) {
    DG_REQUIRE(g.ldc == g.N, "gemm_simt: split-K output must be dense");
    const int64_t n = (int64_t)g.M * g.N;
    launch_k(reduce_partials_kernel, reduce_grid(n), 256, 0, st, g.partial, (float*)g.C, g.splitk, n);
    DG_LAUNCH_CHECK();
  }
}

// split-K factor for the weight-gradient GEMMs (contraction over tokens)
static inline int pick_splitk(int64_t K) {
  int64_t s = K / 1024;
  if (s < 1) s = 1;
  if (s > 32) s = 32;
  return (int)s;
}

}  // namespace dgvit
