// qnet.cuh — CNN twin-Q critic `QNetwork` (vn/got_sac_network.py:125-170), the reference's shipped default
// `critic_type` (vn/config.yaml:61): three 5x5 stride-2 convolutions + ReLU, global average pool, concat with the
// ReLU'd goal embedding and the action, twin 290 -> 128 -> 32 -> n_act heads.
//
// Activations are kept channels-last ([B, H, W, C] fp32) so every kernel reads and writes contiguous channel
// vectors.  conv1 (one input channel, 25 taps) is a direct CUDA-core kernel; conv2 / conv3 are GEMMs over an
// explicit patch matrix ("col", taps outer / channels inner, so its rows are contiguous copies of input pixels)
// against a tap-permuted copy of the weights, through the same gemm() dispatch as the transformer: tcgen05/TMA
// with bf16 operands (precision bf16) or the fp32 CUDA-core kernel (precision fp32, parity mode).  The backward is
// the mirror image: dW = dY^T col (split-K), dcol = dY W, a gather-form col2im, and a direct kernel for conv1's
// weight gradient.  The heads reuse heads.cuh.
#pragma once
#include "heads.cuh"

namespace dgvit {
namespace qnet {

constexpr int KS = 5, ST = 2, TAPS = KS * KS;
constexpr int C1 = 16, C2 = 64, C3 = 256, EMB = 32;

struct Geo {
  int B, H0, W0, H1, W1, H2, W2, H3, W3, na, nps, K0;
  int64_t R1, R2, R3;
  Geo(int B_, int h, int w, int na_, int nps_) {
    B = B_; H0 = h; W0 = w; na = na_; nps = nps_;
    H1 = (H0 - KS) / ST + 1; W1 = (W0 - KS) / ST + 1;
    H2 = (H1 - KS) / ST + 1; W2 = (W1 - KS) / ST + 1;
    H3 = (H2 - KS) / ST + 1; W3 = (W2 - KS) / ST + 1;
    R1 = (int64_t)B * H1 * W1; R2 = (int64_t)B * H2 * W2; R3 = (int64_t)B * H3 * W3;
    K0 = C3 + EMB + na;
  }
};

static void make_layout(int na, int nps, dgvit_qnet_layout& L) {
  int64_t off = 0;
  auto take = [&](int64_t n) {
    int64_t o = off;
    off += (n + DGVIT_ALIGN_FLOATS - 1) / DGVIT_ALIGN_FLOATS * DGVIT_ALIGN_FLOATS;
    return o;
  };
  const int cin[3] = {1, C1, C2}, cout[3] = {C1, C2, C3};
  for (int i = 0; i < 3; ++i) { L.conv_w[i] = take((int64_t)cout[i] * cin[i] * TAPS); L.conv_b[i] = take(cout[i]); }
  const int K0 = C3 + EMB + na;
  L.fc1_w = take(128 * K0); L.fc1_b = take(128);
  L.fc2_w = take(32 * 128); L.fc2_b = take(32);
  L.fc3_w = take(na * 32); L.fc3_b = take(na);
  L.embed_w = take(EMB * nps); L.embed_b = take(EMB);
  L.fc11_w = take(128 * K0); L.fc11_b = take(128);
  L.fc21_w = take(32 * 128); L.fc21_b = take(32);
  L.fc31_w = take(na * 32); L.fc31_b = take(na);
  L.total = off;
}

// ------------------------------------------------------------------ kernels
// conv1 + ReLU: x [B,H0,W0] -> A1 [B,H1,W1,16].  Thread = output pixel, 16 channel accumulators.
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <typename TO>
__global__ void __launch_bounds__(256) conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, TO* __restrict__ out, int64_t R1,
                                                        int H0, int W0, int H1, int W1) {
  pdl_wait();
  pdl_launch();
  __shared__ float ws[TAPS][C1];
  __shared__ float bs[C1];
  for (int i = threadIdx.x; i < TAPS * C1; i += blockDim.x) ws[i % TAPS][i / TAPS] = w[i];   // w is [c][tap]
  if (threadIdx.x < C1) bs[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < (unsigned)R1; p += gridDim.x * blockDim.x) {
    const unsigned pr = p / W1, ox = p - pr * W1, bi = pr / H1, oy = pr - bi * H1;      // (R1 < 2^31: checked by the host)
    const float* xp = x + ((size_t)bi * H0 + oy * ST) * W0 + ox * ST;
    float acc[C1];
#pragma unroll
    for (int c = 0; c < C1; ++c) acc[c] = bs[c];
#pragma unroll
    for (int ky = 0; ky < KS; ++ky)
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const float v = __ldg(xp + ky * W0 + kx);   // (the frames are not written inside the update)
#pragma unroll
        for (int c = 0; c < C1; ++c) acc[c] = fmaf(ws[ky * KS + kx][c], v, acc[c]);
      }
    TO* o = out + (size_t)p * C1;
#pragma unroll
    for (int c = 0; c < C1; c += 4)
      st4<TO>(o + c, make_float4(fmaxf(acc[c], 0.f), fmaxf(acc[c + 1], 0.f), fmaxf(acc[c + 2], 0.f), fmaxf(acc[c + 3], 0.f)));
  }
}

template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<bf16>(bf16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a); u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<bf16>(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}

// patch matrix: col[(b,oy,ox)][tap*C + c] = A[b, 2oy+ky, 2ox+kx, c], both in the operand dtype (the activations A1 / A2 are
// kept in it: the GEMM would round them to it anyway).  Work item = 8 consecutive elements of col (one 16-byte copy in the
// bf16 path): consecutive threads write consecutive pieces of a col row and read consecutive channel groups
// of one input pixel (and, for C = 16, of the next tap's pixel, which is the next pixel of the frame): both sides are
// coalesced.  Four items per thread are loaded before the first is stored.
template <typename T> struct Vec8;                       // 8 consecutive elements in registers
template <> struct Vec8<float> { float4 a, b; };
template <> struct Vec8<bf16> { uint4 a; };
__device__ __forceinline__ void ld_vec8(const float* p, Vec8<float>& v) {
  v.a = reinterpret_cast<const float4*>(p)[0]; v.b = reinterpret_cast<const float4*>(p)[1];
}
__device__ __forceinline__ void ld_vec8(const bf16* p, Vec8<bf16>& v) { v.a = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st_vec8(float* p, const Vec8<float>& v) {
  reinterpret_cast<float4*>(p)[0] = v.a; reinterpret_cast<float4*>(p)[1] = v.b;
}
__device__ __forceinline__ void st_vec8(bf16* p, const Vec8<bf16>& v) {
  asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.a.x), "r"(v.a.y), "r"(v.a.z), "r"(v.a.w) : "memory");
}
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ A, T* __restrict__ col, int64_t items, int Hi, int Wi,
                                                     int Ho, int Wo, int C) {
  pdl_wait();
  pdl_launch();
  const unsigned C8 = C / 8;
  const unsigned per_row = TAPS * C8;
  constexpr int U = 4;
  // (32-bit index arithmetic: the host checks items < 2^31; 64-bit divisions cost more than the copy itself)
  const unsigned n_items = (unsigned)items, stride = gridDim.x * blockDim.x;
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n_items; i0 += U * stride) {
    Vec8<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned i = i0 + u * stride;
      if (i < n_items) {
        const unsigned r = i / per_row, q = i - r * per_row;
        const unsigned tap = q / C8, c8 = q - tap * C8;
        const unsigned rr = r / Wo, ox = r - rr * Wo;
        const unsigned bimg = rr / Ho, oy = rr - bimg * Ho;
        const unsigned ky = tap / KS, kx = tap - ky * KS;
        ld_vec8(A + ((size_t)(bimg * Hi + oy * ST + ky) * Wi + ox * ST + kx) * C + c8 * 8, v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned i = i0 + u * stride;
      if (i < n_items) st_vec8(col + (size_t)i * 8, v[u]);
    }
  }
}

// dA[b,iy,ix,c] = relu'(A) * sum over the (<= 3x3) output pixels whose window covers (iy,ix) of dcol[(b,oy,ox)][tap*C+c]
// (work item = (input pixel, 4 channels); an 8-channel / 16-byte-load version measured slower: 162 vs 96 us)
template <typename T, typename TO>
__global__ void col2im_kernel(const T* __restrict__ dcol, const T* __restrict__ A, TO* __restrict__ dA, int64_t items,
                              int Hi, int Wi, int Ho, int Wo, int C) {
  pdl_wait();
  pdl_launch();
  const unsigned C4 = C / 4, n_items = (unsigned)items;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const unsigned pix = i / C4, c4 = i - pix * C4;
    const unsigned pr = pix / Wi, ix = pix - pr * Wi;
    const unsigned b = pr / Hi, iy = pr - b * Hi;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ky = iy & 1; ky < KS; ky += 2) {
      const int oy = ((int)iy - ky) / 2;
      if ((int)iy - ky < 0 || oy >= Ho) continue;
      for (int kx = ix & 1; kx < KS; kx += 2) {
        const int ox = ((int)ix - kx) / 2;
        if ((int)ix - kx < 0 || ox >= Wo) continue;
        const float4 v = ld4<T>(dcol + ((size_t)(b * Ho + oy) * Wo + ox) * ((size_t)TAPS * C) + (ky * KS + kx) * C + c4 * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    const float4 a = ld4<T>(A + (size_t)pix * C + c4 * 4);
    acc.x = a.x > 0.f ? acc.x : 0.f; acc.y = a.y > 0.f ? acc.y : 0.f;
    acc.z = a.z > 0.f ? acc.z : 0.f; acc.w = a.w > 0.f ? acc.w : 0.f;
    st4<TO>(dA + (size_t)pix * C + c4 * 4, acc);
  }
}

// W [N][C][tap] fp32 -> Wp [N][tap][C] in the operand dtype ; and the inverse for the gradient
template <typename T>
__global__ void wperm_kernel(const float* __restrict__ W, T* __restrict__ Wp, int64_t n, int C) {
  pdl_wait();
  pdl_launch();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into Wp
  if (i >= n) return;
  const int c = (int)(i % C), tap = (int)((i / C) % TAPS);
  const int64_t o = i / ((int64_t)C * TAPS);
  stf(Wp + i, W[(o * C + c) * TAPS + tap]);
}
__global__ void wperm_back_kernel(const float* __restrict__ dWp, float* __restrict__ dW, int64_t n, int C) {
  pdl_wait();
  pdl_launch();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into dW
  if (i >= n) return;
  const int tap = (int)(i % TAPS), c = (int)((i / TAPS) % C);
  const int64_t o = i / ((int64_t)C * TAPS);
  dW[i] = dWp[(o * TAPS + tap) * C + c];
}

// pooled[b][c] = mean over the H3*W3 pixels of A3[b,:,:,c]
__global__ void avgpool_fwd_kernel(const float* __restrict__ A3, float* __restrict__ pooled, int B, int HW) {
  pdl_wait();
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C3) return;
  const int b = i / C3, c = i % C3;
  const float* p = A3 + (int64_t)b * HW * C3 + c;
  float s0 = 0.f, s1 = 0.f;
  int k = 0;
  for (; k + 1 < HW; k += 2) { s0 += p[(int64_t)k * C3]; s1 += p[(int64_t)(k + 1) * C3]; }
  if (k < HW) s0 += p[(int64_t)k * C3];
  pooled[i] = (s0 + s1) / (float)HW;
}
// dY3[(b,pix)][c] = relu'(A3) * dpool[b][c] / HW   (operand dtype of the conv3 backward GEMMs)
template <typename T>
__global__ void avgpool_bwd_kernel(const float* __restrict__ dx, int ldx, const float* __restrict__ A3, T* __restrict__ dY3,
                                   int64_t items, int HW) {
  pdl_wait();
  pdl_launch();
  const float inv = 1.0f / (float)HW;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)items; i += gridDim.x * blockDim.x) {
    const unsigned c4 = i % (C3 / 4), r = i / (C3 / 4), b = r / HW;
    const float4 a = *reinterpret_cast<const float4*>(A3 + (size_t)r * C3 + c4 * 4);
    const float* d = dx + (size_t)b * ldx + c4 * 4;
    st4<T>(dY3 + (size_t)r * C3 + c4 * 4, make_float4(a.x > 0.f ? d[0] * inv : 0.f, a.y > 0.f ? d[1] * inv : 0.f,
                                              a.z > 0.f ? d[2] * inv : 0.f, a.w > 0.f ? d[3] * inv : 0.f));
  }
}

// xcat[b] = [pooled[b] (256) | relu(fc_embed(pstate[b])) (32) | action[b] (na)]
__global__ void concat_kernel(const float* __restrict__ pooled, const float* __restrict__ pstate, const float* __restrict__ We,
                              const float* __restrict__ be, const float* __restrict__ action, float* __restrict__ xcat, int B,
                              int nps, int na) {
  pdl_wait();
  pdl_launch();
  const int K0 = C3 + EMB + na;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K0) return;
  const int b = i / K0, j = i % K0;
  float v;
  if (j < C3) v = pooled[b * C3 + j];
  else if (j < C3 + EMB) {
    const int e = j - C3;
    v = be[e];
    for (int k = 0; k < nps; ++k) v = fmaf(pstate[b * nps + k], We[e * nps + k], v);
    v = fmaxf(v, 0.f);
  } else v = action[b * na + (j - C3 - EMB)];
  xcat[i] = v;
}
// dx = dxa + dxb ; demb = relu'(emb) * dx[256:288] ; d_action = dx[288:]
__global__ void split_kernel(const float* __restrict__ dxa, const float* __restrict__ dxb, const float* __restrict__ xcat,
                             float* __restrict__ dx, float* __restrict__ demb, float* __restrict__ d_action, int B, int na) {
  pdl_wait();
  pdl_launch();
  const int K0 = C3 + EMB + na;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K0) return;
  const int b = i / K0, j = i % K0;
  const float v = dxa[i] + dxb[i];
  dx[i] = v;
  if (j >= C3 && j < C3 + EMB) demb[b * EMB + (j - C3)] = xcat[i] > 0.f ? v : 0.f;
  if (j >= C3 + EMB && d_action) d_action[b * na + (j - C3 - EMB)] = v;
}

// conv1 weight / bias gradient: part[block][c*25 + tap] = sum over this block's output pixels of dY1[p][c] * x[p, tap],
// part[block][400 + c] = sum dY1[p][c].  128 pixels are staged per round; each of the 4 warps takes 32 of them with
// lane = tap (lane 25 multiplies by 1: the bias sums) and all 16 channels in registers, so a pixel costs one shared load of
// x and four broadcast 16-byte loads of dY1 for 16 FMAs per lane (the first version had one accumulator per thread and two
// shared loads per FMA: 370 us at B = 256).
constexpr int C1W_PIX = 128, C1W_THREADS = 128, C1W_N = C1 * TAPS + C1;
__global__ void __launch_bounds__(C1W_THREADS) conv1_bwd_w_kernel(const float* __restrict__ dY1, const float* __restrict__ x,
                                                                 float* __restrict__ part, int64_t R1, int64_t pix_per_block,
                                                                 int H0, int W0, int H1, int W1) {
  pdl_wait();
  pdl_launch();
  __shared__ __align__(16) float dys[C1W_PIX][C1];
  __shared__ float xs[C1W_PIX][TAPS];
  __shared__ float red[4][C1W_N];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block, p1 = min(R1, p0 + pix_per_block);
  float acc[C1];
#pragma unroll
  for (int c = 0; c < C1; ++c) acc[c] = 0.f;
  for (int64_t q0 = p0; q0 < p1; q0 += C1W_PIX) {
    __syncthreads();
    for (int i = tid; i < C1W_PIX * (C1 / 4); i += C1W_THREADS) {
      const int64_t p = q0 + i / (C1 / 4);
      reinterpret_cast<float4*>(&dys[0][0])[i] =
          p < p1 ? reinterpret_cast<const float4*>(dY1 + p * C1)[i % (C1 / 4)] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {                                                    // thread = staged pixel: its coordinates once, then the 25 taps
      static_assert(C1W_PIX == C1W_THREADS, "one staged pixel per thread");
      const int64_t p = q0 + tid;
      if (p < p1) {
        const unsigned pu = (unsigned)p, pr = pu / W1, ox = pu - pr * W1, bi = pr / H1, oy = pr - bi * H1;
        const float* src = x + ((size_t)bi * H0 + oy * ST) * W0 + ox * ST;
#pragma unroll
        for (int t = 0; t < TAPS; ++t) xs[tid][t] = __ldg(src + (t / KS) * W0 + t % KS);
      } else {
#pragma unroll
        for (int t = 0; t < TAPS; ++t) xs[tid][t] = 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int s = warp * 32; s < warp * 32 + 32; ++s) {
      const float xv = lane < TAPS ? xs[s][lane] : 1.0f;
      const float4* d = reinterpret_cast<const float4*>(&dys[s][0]);
#pragma unroll
      for (int c4 = 0; c4 < C1 / 4; ++c4) {
        const float4 v = d[c4];
        acc[4 * c4] = fmaf(v.x, xv, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(v.y, xv, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(v.z, xv, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(v.w, xv, acc[4 * c4 + 3]);
      }
    }
  }
  if (lane <= TAPS) {
#pragma unroll
    for (int c = 0; c < C1; ++c) red[warp][lane < TAPS ? c * TAPS + lane : C1 * TAPS + c] = acc[c];
  }
  __syncthreads();
  float* P = part + (int64_t)blockIdx.x * C1W_N;
  for (int j = tid; j < C1W_N; j += C1W_THREADS) P[j] = (red[0][j] + red[1][j]) + (red[2][j] + red[3][j]);
}

// out_a[j] (j < split) / out_b[j - split] = sum over nb partial rows of part[k][j], k ascending
__global__ void reduce_cols_kernel(const float* __restrict__ part, float* __restrict__ outa, float* __restrict__ outb, int nb,
                                   int n, int split) {
  pdl_wait();
  pdl_launch();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int k = 0;
  for (; k + 3 < nb; k += 4) {
    s0 += part[(int64_t)k * n + j]; s1 += part[(int64_t)(k + 1) * n + j];
    s2 += part[(int64_t)(k + 2) * n + j]; s3 += part[(int64_t)(k + 3) * n + j];
  }
  for (; k < nb; ++k) s0 += part[(int64_t)k * n + j];
  const float s = (s0 + s1) + (s2 + s3);
  if (j < split) outa[j] = s; else outb[j - split] = s;
}

// ------------------------------------------------------------------ workspace
template <typename T>
struct Ws {
  T *A1, *A2;                // activations of conv1 / conv2 in the operand dtype, channels-last
  float *A3, *pooled, *xcat, *h1a, *h2a, *h1b, *h2b;
  T *col2, *col3, *Wp2, *Wp3;
  // backward
  float *dxa, *dxb, *dx, *demb, *dh1a, *dh2a, *dh1b, *dh2b, *dY1, *dWp, *dWp2, *partial;
  T *dY3, *dY2, *dcol;
  size_t partial_floats;
};
template <typename T>
static void carve(Carver& cv, const Geo& g, Ws<T>& w) {
  w.A1 = cv.take<T>(g.R1 * C1);
  w.A2 = cv.take<T>(g.R2 * C2);
  w.A3 = cv.take<float>(g.R3 * C3);
  w.pooled = cv.take<float>((int64_t)g.B * C3);
  w.xcat = cv.take<float>((int64_t)g.B * g.K0);
  w.h1a = cv.take<float>((int64_t)g.B * 128); w.h2a = cv.take<float>((int64_t)g.B * 32);
  w.h1b = cv.take<float>((int64_t)g.B * 128); w.h2b = cv.take<float>((int64_t)g.B * 32);
  w.col2 = cv.take<T>(g.R2 * (int64_t)(TAPS * C1));
  w.col3 = cv.take<T>(g.R3 * (int64_t)(TAPS * C2));
  w.Wp2 = cv.take<T>((int64_t)C2 * TAPS * C1);
  w.Wp3 = cv.take<T>((int64_t)C3 * TAPS * C2);
  w.dxa = cv.take<float>((int64_t)g.B * g.K0); w.dxb = cv.take<float>((int64_t)g.B * g.K0);
  w.dx = cv.take<float>((int64_t)g.B * g.K0);
  w.demb = cv.take<float>((int64_t)g.B * EMB);
  w.dh1a = cv.take<float>((int64_t)g.B * 128); w.dh2a = cv.take<float>((int64_t)g.B * 32);
  w.dh1b = cv.take<float>((int64_t)g.B * 128); w.dh2b = cv.take<float>((int64_t)g.B * 32);
  w.dY1 = cv.take<float>(g.R1 * C1);
  w.dWp = cv.take<float>((int64_t)C3 * TAPS * C2);
  w.dWp2 = cv.take<float>((int64_t)C2 * TAPS * C1);
  w.dY3 = cv.take<T>(g.R3 * C3);
  w.dY2 = cv.take<T>(g.R2 * C2);
  w.dcol = cv.take<T>(std::max(g.R3 * (int64_t)(TAPS * C2), g.R2 * (int64_t)(TAPS * C1)));
  // split-K partials of both dW GEMMs, column-sum partials of the bias gradients, conv1's per-block partials: all reduced
  // by ONE deferred launch at the end of the backward
  w.partial_floats = (size_t)32 * C3 * TAPS * C2 + (size_t)32 * C2 * TAPS * C1 + ((size_t)1 << 21);
  w.partial = cv.take<float>(w.partial_floats);
}

static unsigned gs(int64_t n, int bs = 256) {
  int64_t g = cdiv(n, bs);
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(g, 148 * 32));
}

// ------------------------------------------------------------------ forward / backward
template <typename T>
static void forward(const float* P, const dgvit_qnet_layout& L, const Geo& g, const float* img, const float* pstate,
                    const float* action, float* q1, float* q2, Ws<T>& w, cudaStream_t st) {
  DG_REQUIRE(g.R1 * C1 < ((int64_t)1 << 31) && g.R2 * TAPS * C1 < ((int64_t)1 << 31) && g.R3 * TAPS * C2 < ((int64_t)1 << 31),
             "qnet: batch too large for the 32-bit index arithmetic of the patch-matrix kernels (B <= ~5000 at 128x160)");
  launch_k(conv1_fwd_kernel<T>, gs(g.R1), 256, 0, st, img, P + L.conv_w[0], P + L.conv_b[0], w.A1, g.R1, g.H0, g.W0, g.H1, g.W1);
  DG_LAUNCH_CHECK();
  // conv2
  {
    const int64_t n = (int64_t)C2 * TAPS * C1;
    launch_k(wperm_kernel<T>, (unsigned)cdiv(n, 256), 256, 0, st, P + L.conv_w[1], w.Wp2, n, C1);
    DG_LAUNCH_CHECK();
    const int64_t items = g.R2 * TAPS * (C1 / 8);
    launch_k(im2col_kernel<T>, gs(cdiv(items, 4)), 256, 0, st, (const T*)w.A1, w.col2, items, g.H1, g.W1, g.H2, g.W2, C1);
    DG_LAUNCH_CHECK();
    linear_fwd<T, T, T>(w.col2, w.Wp2, w.A2, g.R2, C2, TAPS * C1, EPI_BIAS_RELU, P + L.conv_b[1], st);
  }
  // conv3
  {
    const int64_t n = (int64_t)C3 * TAPS * C2;
    launch_k(wperm_kernel<T>, (unsigned)cdiv(n, 256), 256, 0, st, P + L.conv_w[2], w.Wp3, n, C2);
    DG_LAUNCH_CHECK();
    const int64_t items = g.R3 * TAPS * (C2 / 8);
    launch_k(im2col_kernel<T>, gs(cdiv(items, 4)), 256, 0, st, (const T*)w.A2, w.col3, items, g.H2, g.W2, g.H3, g.W3, C2);
    DG_LAUNCH_CHECK();
    linear_fwd<T, T, float>(w.col3, w.Wp3, w.A3, g.R3, C3, TAPS * C2, EPI_BIAS_RELU, P + L.conv_b[2], st);
  }
  launch_k(avgpool_fwd_kernel, (unsigned)cdiv((int64_t)g.B * C3, 256), 256, 0, st, (const float*)w.A3, w.pooled, g.B, g.H3 * g.W3);
  DG_LAUNCH_CHECK();
  launch_k(concat_kernel, (unsigned)cdiv((int64_t)g.B * g.K0, 256), 256, 0, st, (const float*)w.pooled, pstate, P + L.embed_w,
           P + L.embed_b, action, w.xcat, g.B, g.nps, g.na);
  DG_LAUNCH_CHECK();
  heads::FwdArgs h;
  memset(&h, 0, sizeof(h));
  h.nheads = 2; h.B = g.B; h.K1 = g.K0; h.K2 = 0; h.H2 = 32; h.NOa = g.na; h.NOb = 0;
  h.x1 = w.xcat; h.x2 = nullptr; h.xcat = nullptr;
  h.w[0] = heads::HeadW{P + L.fc1_w, P + L.fc1_b, P + L.fc2_w, P + L.fc2_b, P + L.fc3_w, P + L.fc3_b, nullptr, nullptr,
                        w.h1a, w.h2a, q1, nullptr};
  h.w[1] = heads::HeadW{P + L.fc11_w, P + L.fc11_b, P + L.fc21_w, P + L.fc21_b, P + L.fc31_w, P + L.fc31_b, nullptr, nullptr,
                        w.h1b, w.h2b, q2, nullptr};
  heads::launch_fwd(h, st);
}

// conv layer weight gradient through the patch matrix: dWp (tap-permuted; split-K partials queued on `rl`) and db
template <typename T>
static void conv_bwd_w(const T* dY, const T* col, float* dWp, float* db, int64_t R, int Cout, int Cin, ReduceList& rl,
                       cudaStream_t st) {
  linear_bwd_w<T, T>(dY, col, dWp, db, R, Cout, TAPS * Cin, nullptr, st, -1, -1, &rl);
}

template <typename T>
static void backward(const float* P, float* G, const dgvit_qnet_layout& L, const Geo& g, const float* img, const float* pstate,
                     const float* dq1, const float* dq2, float* d_action, bool param_grads, Ws<T>& w, cudaStream_t st) {
  {
    heads::BwdArgs h;
    memset(&h, 0, sizeof(h));
    h.nheads = 2; h.B = g.B; h.K0 = g.K0; h.H2 = 32; h.NOa = g.na; h.NOb = 0;
    h.h[0] = heads::BwdHead{P + L.fc1_w, P + L.fc2_w, P + L.fc3_w, nullptr, w.h1a, w.h2a, dq1, nullptr, w.dh1a, w.dh2a, w.dxa};
    h.h[1] = heads::BwdHead{P + L.fc11_w, P + L.fc21_w, P + L.fc31_w, nullptr, w.h1b, w.h2b, dq2, nullptr, w.dh1b, w.dh2b, w.dxb};
    heads::launch_bwd_dx(h, st);
  }
  launch_k(split_kernel, (unsigned)cdiv((int64_t)g.B * g.K0, 256), 256, 0, st, (const float*)w.dxa, (const float*)w.dxb,
           (const float*)w.xcat, w.dx, w.demb, d_action, g.B, g.na);
  DG_LAUNCH_CHECK();
  if (!param_grads) return;
  {
    heads::DwList dw(g.B);
    dw.add(dq1, g.na, w.h2a, G + L.fc3_w, G + L.fc3_b, g.na, 32);
    dw.add(w.dh2a, 32, w.h1a, G + L.fc2_w, G + L.fc2_b, 32, 128);
    dw.add(w.dh1a, 128, w.xcat, G + L.fc1_w, G + L.fc1_b, 128, g.K0);
    dw.add(dq2, g.na, w.h2b, G + L.fc31_w, G + L.fc31_b, g.na, 32);
    dw.add(w.dh2b, 32, w.h1b, G + L.fc21_w, G + L.fc21_b, 32, 128);
    dw.add(w.dh1b, 128, w.xcat, G + L.fc11_w, G + L.fc11_b, 128, g.K0);
    dw.add(w.demb, EMB, pstate, G + L.embed_w, G + L.embed_b, EMB, g.nps);
    dw.launch(st);
  }
  ReduceList rl(w.partial, w.partial_floats);
  // conv3
  {
    const int64_t items = g.R3 * (C3 / 4);
    launch_k(avgpool_bwd_kernel<T>, gs(items), 256, 0, st, (const float*)w.dx, g.K0, (const float*)w.A3, w.dY3, items, g.H3 * g.W3);
    DG_LAUNCH_CHECK();
    conv_bwd_w<T>(w.dY3, w.col3, w.dWp, G + L.conv_b[2], g.R3, C3, C2, rl, st);
    linear_bwd_x<T, T, T>(w.dY3, w.Wp3, w.dcol, g.R3, C3, TAPS * C2, EPI_NONE, nullptr, 0, st);
    const int64_t it2 = g.R2 * (C2 / 4);
    launch_k(col2im_kernel<T, T>, gs(it2), 256, 0, st, (const T*)w.dcol, (const T*)w.A2, w.dY2, it2, g.H2, g.W2, g.H3, g.W3, C2);
    DG_LAUNCH_CHECK();
  }
  // conv2
  {
    conv_bwd_w<T>(w.dY2, w.col2, w.dWp2, G + L.conv_b[1], g.R2, C2, C1, rl, st);
    linear_bwd_x<T, T, T>(w.dY2, w.Wp2, w.dcol, g.R2, C2, TAPS * C1, EPI_NONE, nullptr, 0, st);
    const int64_t it1 = g.R1 * (C1 / 4);
    launch_k(col2im_kernel<T, float>, gs(it1), 256, 0, st, (const T*)w.dcol, (const T*)w.A1, w.dY1, it1, g.H1, g.W1, g.H2, g.W2, C1);
    DG_LAUNCH_CHECK();
  }
  // conv1
  {
    const int nblocks = (int)std::min<int64_t>(148 * 8, cdiv(g.R1, C1W_PIX));
    const int64_t ppb = cdiv(cdiv(g.R1, nblocks), C1W_PIX) * C1W_PIX;
    const int nb = (int)cdiv(g.R1, ppb);
    float* part = rl.alloc((size_t)nb * C1W_N);
    launch_k(conv1_bwd_w_kernel, nb, C1W_THREADS, 0, st, (const float*)w.dY1, img, part, g.R1, ppb, g.H0, g.W0, g.H1, g.W1);
    DG_LAUNCH_CHECK();
    rl.add(part, G + L.conv_w[0], nb, C1 * TAPS, C1W_N);
    rl.add(part + C1 * TAPS, G + L.conv_b[0], nb, C1, C1W_N);
  }
  rl.launch(st);       // every split-K / column-sum / per-block partial of this backward, one launch, fixed summation order
  {
    const int64_t n3 = (int64_t)C3 * TAPS * C2, n2 = (int64_t)C2 * TAPS * C1;
    launch_k(wperm_back_kernel, (unsigned)cdiv(n3, 256), 256, 0, st, (const float*)w.dWp, G + L.conv_w[2], n3, C2);
    DG_LAUNCH_CHECK();
    launch_k(wperm_back_kernel, (unsigned)cdiv(n2, 256), 256, 0, st, (const float*)w.dWp2, G + L.conv_w[1], n2, C1);
    DG_LAUNCH_CHECK();
  }
}

}  // namespace qnet
}  // namespace dgvit
