// depth.cu — depth normalise + Gaussian-noise augmentation + blur + 4x bilinear resize.
// Restates vn/env_lab.py:420-434 (callback: cv2.normalize MINMAX -> uint8), :78-90
// (add_nose: +N(0,50), clip, GaussianBlur 5x5), :69-76 (blurring: 11x11 on the centre band),
// :295-299 (cv2.resize to (W/4,H/4), /255).  HBM-bound separable passes.
#include "common.cuh"

namespace dgvit {

constexpr int MM_BLOCKS = 64;

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// pass 0: per-frame min/max partials  part[f][blk][2]
__global__ void depth_minmax_kernel(const float* __restrict__ raw, float* __restrict__ part, int64_t hw) {
  pdl_wait();
  pdl_launch();
  const int f = blockIdx.y;
  const float* src = raw + (int64_t)f * hw;
  float mn = INFINITY, mx = -INFINITY;
  const int64_t hw4 = hw >> 2;          // (H and W are multiples of 4, frames 16-byte aligned)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
    mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  __shared__ float smn[32], smx[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    part[((int64_t)f * gridDim.x + blockIdx.x) * 2 + 0] = mn;
    part[((int64_t)f * gridDim.x + blockIdx.x) * 2 + 1] = mx;
  }
}

// N(0,1) draw of pixel gi: one Philox4x32 call serves the 4 pixels of an aligned group (two Box-Muller pairs)
__device__ __forceinline__ float normal_at(const uint64_t* rng, int64_t gi) {
  uint32_t r[4];
  philox4x32(rng[0], (uint64_t)(gi >> 2), 0x6465707468000000ull | (rng[1] & 0xffffffffu), r);
  const int q = (int)(gi & 3);
  const float rad = sqrtf(-2.0f * __logf(u01(r[q & 2])));
  float sn, cs;
  __sincosf(6.283185307179586f * u01(r[(q & 2) + 1]), &sn, &cs);
  return rad * ((q & 1) ? sn : cs);
}
__device__ __forceinline__ float noisy_px(const float* __restrict__ raw, const float* __restrict__ noise,
                                          const uint64_t* rng, int64_t gi, double scale, double shift) {
  // cv2.normalize(NORM_MINMAX, 0..255) then .astype(uint8) (truncation)
  const float nrm = (float)((double)raw[gi] * scale + shift);
  const float u8 = (float)(int)fminf(fmaxf(nrm, 0.f), 255.f);
  const float nz = noise ? noise[gi] : 50.0f * normal_at(rng, gi);
  return fminf(fmaxf(u8 + nz, 0.f), 255.f);
}

// pass 1: noisy image + horizontal 5-tap [1 4 6 4 1]/16, BORDER_REFLECT_101
__global__ void depth_noise_hblur_kernel(const float* __restrict__ raw, const float* __restrict__ noise,
                                         const uint64_t* rng, const float* __restrict__ mmpart,
                                         float* __restrict__ S1, int H, int W) {
  pdl_wait();
  pdl_launch();
  const int f = blockIdx.z, y = blockIdx.y;
  __shared__ double sc[2];
  if (threadIdx.x == 0) {
    float mn = INFINITY, mx = -INFINITY;
    for (int b = 0; b < MM_BLOCKS; ++b) {
      mn = fminf(mn, mmpart[((int64_t)f * MM_BLOCKS + b) * 2]);
      mx = fmaxf(mx, mmpart[((int64_t)f * MM_BLOCKS + b) * 2 + 1]);
    }
    const double rng_ = (double)mx - (double)mn;
    const double s = rng_ > 2.220446049250313e-16 ? 255.0 / rng_ : 0.0;
    sc[0] = s;
    sc[1] = 0.0 - (double)mn * s;
  }
  __syncthreads();
  const double scale = sc[0], shift = sc[1];
  const int64_t rowbase = ((int64_t)f * H + y) * W;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
    const float k[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 5; ++t)
      s = fmaf(k[t], noisy_px(raw, noise, rng, rowbase + reflect101(x + t - 2, W), scale, shift), s);
    S1[rowbase + x] = s;
  }
}

// pass 2: vertical 5-tap
__global__ void depth_vblur5_kernel(const float* __restrict__ S1, float* __restrict__ S2, int H, int W) {
  pdl_wait();
  pdl_launch();
  const int f = blockIdx.z, y = blockIdx.y;
  const float k[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 5; ++t) s = fmaf(k[t], S1[((int64_t)f * H + reflect101(y + t - 2, H)) * W + x], s);
    S2[((int64_t)f * H + y) * W + x] = s;
  }
}

struct K11 { float k[11]; };

// pass 3: horizontal 11-tap on the centre band rows (reflect in x)
__global__ void depth_band_hblur_kernel(const float* __restrict__ S2, float* __restrict__ T1, K11 kk, int H, int W,
                                        int y1, int bh) {
  pdl_wait();
  pdl_launch();
  const int f = blockIdx.z, r = blockIdx.y;
  const float* src = S2 + ((int64_t)f * H + y1 + r) * W;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 11; ++t) s = fmaf(kk.k[t], src[reflect101(x + t - 5, W)], s);
    T1[((int64_t)f * bh + r) * W + x] = s;
  }
}

// pass 4: band vertical 11-tap (reflect inside the band) + bilinear downsample by `fac` + /255
__global__ void depth_resize_kernel(const float* __restrict__ S2, const float* __restrict__ T1, K11 kk,
                                    float* __restrict__ out, int H, int W, int y1, int bh, int fac) {
  pdl_wait();
  pdl_launch();
  const int f = blockIdx.z, oy = blockIdx.y;
  const int oh = H / fac, ow = W / fac, o = fac / 2 - 1;
  for (int ox = blockIdx.x * blockDim.x + threadIdx.x; ox < ow; ox += gridDim.x * blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int y = oy * fac + o + dy;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int x = ox * fac + o + dx;
        float v;
        if (y >= y1 && y < y1 + bh) {
          v = 0.f;
#pragma unroll
          for (int t = 0; t < 11; ++t)
            v = fmaf(kk.k[t], T1[((int64_t)f * bh + reflect101(y - y1 + t - 5, bh)) * W + x], v);
        } else {
          v = S2[((int64_t)f * H + y) * W + x];
        }
        acc += 0.25f * v;
      }
    }
    out[((int64_t)f * oh + oy) * ow + ox] = acc / 255.0f;
  }
}

// Four N(0,1) draws of the aligned pixel group gi4 = gi >> 2: ONE Philox4x32 call and two Box-Muller pairs (the per-pixel
// normal_at() above evaluates the same numbers; it recomputes the call for each of the 4 pixels)
__device__ __forceinline__ float4 normal4_at(const uint64_t* rng, int64_t gi4) {
  uint32_t r[4];
  philox4x32(rng[0], (uint64_t)gi4, 0x6465707468000000ull | (rng[1] & 0xffffffffu), r);
  const float rad0 = sqrtf(-2.0f * __logf(u01(r[0]))), rad1 = sqrtf(-2.0f * __logf(u01(r[2])));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u01(r[1]), &s0, &c0);
  __sincosf(6.283185307179586f * u01(r[3]), &s1, &c1);
  return make_float4(rad0 * c0, rad0 * s0, rad1 * c1, rad1 * s1);
}
__device__ __forceinline__ float noisy1(float raw, float nz, double scale, double shift) {
  const float nrm = (float)((double)raw * scale + shift);          // cv2.normalize(NORM_MINMAX, 0..255) ...
  const float u8 = (float)(int)fminf(fmaxf(nrm, 0.f), 255.f);      // ... .astype(uint8) (truncation)
  return fminf(fmaxf(u8 + nz, 0.f), 255.f);
}
// the noisy image A at the aligned 4-pixel group starting at column x4 of row gy (x4 may lie outside the image: every
// pixel is then taken at its BORDER_REFLECT_101 mirror, noise included: the blur mirrors the NOISY image)
__device__ __forceinline__ float4 noisy_group(const float* __restrict__ raw, const float* __restrict__ noise, const uint64_t* rng,
                                              int64_t fbase, int gy, int x4, int W, double scale, double shift) {
  const int64_t rb = fbase + (int64_t)gy * W;
  if (x4 >= 0 && x4 + 3 < W) {
    const float4 rv = *reinterpret_cast<const float4*>(raw + rb + x4);
    float4 nz;
    if (noise) nz = *reinterpret_cast<const float4*>(noise + rb + x4);
    else { nz = normal4_at(rng, (rb + x4) >> 2); nz.x *= 50.f; nz.y *= 50.f; nz.z *= 50.f; nz.w *= 50.f; }
    return make_float4(noisy1(rv.x, nz.x, scale, shift), noisy1(rv.y, nz.y, scale, shift), noisy1(rv.z, nz.z, scale, shift),
                       noisy1(rv.w, nz.w, scale, shift));
  }
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t gi = rb + reflect101(x4 + k, W);
    o[k] = noisy1(raw[gi], noise ? noise[gi] : 50.0f * normal_at(rng, gi), scale, shift);
  }
  return make_float4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------
// Fused pass: noise + 5x5 blur + centre-band 11x11 blur + 4x bilinear resize + /255 in one kernel.
// One CTA = an 8 x 32 tile of OUTPUT pixels (32 x 128 input pixels + halo) staged in shared memory:
//   A  = clip(u8(normalised) + noise)      halo 7 (band tiles) / 2 (other tiles)
//   Bh = horizontal 5-tap of A, S2 = vertical 5-tap of Bh  (the GaussianBlur(5,5) image)
//   T1 = horizontal 11-tap of S2 at the sampled columns (band tiles), vertical 11-tap at the
//        sampled rows with BORDER_REFLECT_101 inside the band, then the 2x2 average = cv2.resize.
// Out-of-image halo positions are filled by mirrored coordinates, which is exact for symmetric
// kernels.  HBM traffic: raw (+noise) once; neighbouring tiles' halos hit L2.
constexpr int TOH = 8, TOW = 32;                 // output tile
constexpr int CORE_R = TOH * 4 - 2, CORE_C = TOW * 4 - 2;   // span of sampled rows / cols (30, 126)
constexpr int A_R = CORE_R + 14, A_C = CORE_C + 14;         // 44 x 140 (halo 7)
constexpr int S_R = CORE_R + 10, S_C = CORE_C + 10;         // 40 x 136 (halo 5)
constexpr int A_P = 148, S_P = S_C + 1;                     // pitches: A rows hold aligned float4 groups (144 columns at most)
static_assert(A_P % 4 == 0 && A_P >= A_C + 4, "A pitch");
constexpr int H_P = TOW + 1;                                // fast path: horizontal 6-tap results [rows][TOW]
constexpr int FUSED_SMEM = (A_R * A_P + A_R * S_P) * 4 + 32;   // S2 reuses A, T1 reuses Bh

__global__ void __launch_bounds__(256) depth_fused_kernel(const float* __restrict__ raw, const float* __restrict__ noise,
                                                          const uint64_t* rng, const float* __restrict__ mmpart,
                                                          K11 kk, float* __restrict__ out, int H, int W, int y1, int bh) {
  pdl_wait();
  pdl_launch();
  extern __shared__ __align__(16) float smf[];
  float* A = smf;                       // [A_R][A_P]
  float* Bh = A + A_R * A_P;            // [A_R][S_P]
  float* S2 = A;                        // [S_R][S_P]     (A is dead once Bh exists)
  float* T1 = Bh;                       // [S_R][2*TOW]   (Bh is dead once S2 exists)
  double* sc = reinterpret_cast<double*>(Bh + A_R * S_P);
  const int f = blockIdx.z, oy0 = blockIdx.y * TOH, ox0 = blockIdx.x * TOW;
  const int oh = H / 4, ow = W / 4;
  const int tid = threadIdx.x;
  if (tid < 32) {        // frame min / max from the 64 partials: one warp, two partials per lane
    static_assert(MM_BLOCKS == 64, "two partials per lane");
    const float2 p0 = *reinterpret_cast<const float2*>(mmpart + ((int64_t)f * MM_BLOCKS + tid) * 2);
    const float2 p1 = *reinterpret_cast<const float2*>(mmpart + ((int64_t)f * MM_BLOCKS + 32 + tid) * 2);
    float mn = fminf(p0.x, p1.x), mx = fmaxf(p0.y, p1.y);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (tid == 0) {
      const double rg = (double)mx - (double)mn;
      const double s = rg > 2.220446049250313e-16 ? 255.0 / rg : 0.0;
      sc[0] = s;
      sc[1] = 0.0 - (double)mn * s;
    }
  }
  const int ys0 = oy0 * 4 + 1, xs0 = ox0 * 4 + 1;           // first sampled row / col of the tile
  const int y2 = y1 + bh;
  const bool band = (ys0 + CORE_R - 1 >= y1) && (ys0 < y2);  // some sampled row lies in the centre band
  const int hs = band ? 5 : 0;                                // halo of S2, A needs hs + 2
  const int ar = CORE_R + 2 * (hs + 2), ac = CORE_C + 2 * (hs + 2);
  const int sr = CORE_R + 2 * hs, scn = CORE_C + 2 * hs;
  __syncthreads();
  const double scale = sc[0], shift = sc[1];
  const float k5[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  const int64_t fbase = (int64_t)f * H * W;
  // stage A: one thread = one aligned group of 4 pixels (16-byte loads of raw / noise, one Philox call per group).
  // Column c of the window sits at A[r][c + ax] where ax = (window start) - (aligned start) in [0, 4).
  const int wx0 = xs0 - hs - 2;                                  // first window column (frame coordinates, may be < 0)
  const int xa = (wx0 >= 0 ? wx0 : wx0 - 3) / 4 * 4;             // aligned down
  const int ax = wx0 - xa;
  const int ng = (ax + ac + 3) / 4;                              // groups per row
  {
    // all 16-byte loads of a thread's groups are issued before any of them is used (the stage is a chain of HBM
    // latencies otherwise: 4 resident CTAs per SM hold one load per thread in flight)
    constexpr int U = 2;
    const int total = ar * ng;
    for (int base = tid; base < total; base += 256 * U) {
      float4 rv[U], nv[U];
      int rr[U], gg[U], gys[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * 256;
        rr[u] = band ? i / 36 : i / 34;                      // ng = 36 (band window) or 34: constant divisors
        gg[u] = i - rr[u] * ng;
        gys[u] = reflect101(ys0 - hs - 2 + min(rr[u], ar - 1), H);
        const int x4 = xa + 4 * gg[u];
        rv[u] = nv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total && x4 >= 0 && x4 + 3 < W) {
          const int64_t gi = fbase + (int64_t)gys[u] * W + x4;
          rv[u] = *reinterpret_cast<const float4*>(raw + gi);
          if (noise) nv[u] = *reinterpret_cast<const float4*>(noise + gi);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * 256;
        if (i >= total) continue;
        const int x4 = xa + 4 * gg[u];
        float4 o;
        if (x4 >= 0 && x4 + 3 < W) {
          float4 nz = nv[u];
          if (!noise) {
            nz = normal4_at(rng, (fbase + (int64_t)gys[u] * W + x4) >> 2);
            nz.x *= 50.f; nz.y *= 50.f; nz.z *= 50.f; nz.w *= 50.f;
          }
          o = make_float4(noisy1(rv[u].x, nz.x, scale, shift), noisy1(rv[u].y, nz.y, scale, shift),
                          noisy1(rv[u].z, nz.z, scale, shift), noisy1(rv[u].w, nz.w, scale, shift));
        } else {
          o = noisy_group(raw, noise, rng, fbase, gys[u], x4, W, scale, shift);      // image border: mirrored pixel by pixel
        }
        *reinterpret_cast<float4*>(A + rr[u] * A_P + 4 * gg[u]) = o;
      }
    }
  }
  __syncthreads();
  if (!band) {
    // Fast path (no sampled row in the centre band: 80 % of the tiles).  cv2.resize by 4 averages the GaussianBlur(5,5)
    // image at rows / columns 4o+1, 4o+2: the composite is ONE separable 6-tap filter with stride 4,
    //   w6 = ([1 4 6 4 1 0] + [0 1 4 6 4 1]) / 32 = [1 5 10 10 5 1] / 32   over pixels 4o-1 .. 4o+4,
    // so only 1/4 of the blurred image is ever formed: horizontal pass at the TOW output columns of every window row
    // (three conflict-free 16-byte shared loads per result), vertical pass at the TOH output rows.
    const float w6[6] = {1.f / 32, 5.f / 32, 10.f / 32, 10.f / 32, 5.f / 32, 1.f / 32};
    float* Hh = Bh;                                              // [ar][H_P]
    for (int i = tid; i < ar * TOW; i += blockDim.x) {
      const int r = i / TOW, j = i % TOW;
      // window column of pixel 4(ox0+j)-1 is 4j (hs = 0: the window starts at 4 ox0 - 1), i.e. A column 4j + ax
      const float* row = A + r * A_P + 4 * j + ax;               // ax = 3 here (wx0 = 4 ox0 - 1): row + 1 is 16-byte aligned
      const float a0 = row[0];
      const float4 m = *reinterpret_cast<const float4*>(row + 1);
      const float a5 = row[5];
      Hh[r * H_P + j] = w6[0] * a0 + w6[1] * m.x + w6[2] * m.y + w6[3] * m.z + w6[4] * m.w + w6[5] * a5;
    }
    __syncthreads();
    const int oy = oy0 + tid / TOW, ox = ox0 + tid % TOW;
    if (oy < oh && ox < ow) {
      const float* col = Hh + (tid / TOW) * 4 * H_P + tid % TOW;
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 6; ++t) acc = fmaf(w6[t], col[t * H_P], acc);
      out[((int64_t)f * oh + oy) * ow + ox] = acc / 255.0f;
    }
    return;
  }
  // band tiles: A columns are addressed through the alignment offset from here on
  A += ax;
  // (band tiles: hs = 5, so ar = A_R, sr = S_R, scn = S_C are compile-time constants: no run-time divisions below)
  for (int i = tid; i < A_R * S_C; i += 256) {                 // horizontal 5-tap
    const int r = i / S_C, c = i % S_C;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 5; ++t) s = fmaf(k5[t], A[r * A_P + c + t], s);
    Bh[r * S_P + c] = s;
  }
  __syncthreads();
  for (int i = tid; i < S_R * S_C; i += 256) {                 // vertical 5-tap -> GaussianBlur(5,5)
    const int r = i / S_C, c = i % S_C;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 5; ++t) s = fmaf(k5[t], Bh[(r + t) * S_P + c], s);
    S2[r * S_P + c] = s;
  }
  __syncthreads();
  if (band) {                                                  // horizontal 11-tap at the sampled columns
    for (int i = tid; i < S_R * 2 * TOW; i += 256) {
      const int r = i / (2 * TOW), j = i % (2 * TOW);
      const int c = (j >> 1) * 4 + (j & 1);                    // sampled col relative to xs0
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) s = fmaf(kk.k[t], S2[r * S_P + c + t], s);   // S2 col index = c + hs - 5 + t, hs = 5
      T1[r * 2 * TOW + j] = s;
    }
    __syncthreads();
  }
  {
    const int oy = oy0 + tid / TOW, ox = ox0 + tid % TOW;
    if (oy < oh && ox < ow) {
      float acc = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int y = oy * 4 + 1 + dy;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int j = (tid % TOW) * 2 + dx;
          float v;
          if (y >= y1 && y < y2) {
            v = 0.f;
#pragma unroll
            for (int t = 0; t < 11; ++t) {
              int yb = y - y1 + t - 5;                                       // reflect inside the band (one fold: bh > 5)
              yb = bh > 5 ? (yb < 0 ? -yb : (yb >= bh ? 2 * (bh - 1) - yb : yb)) : reflect101(yb, bh);
              const int yy = y1 + yb;
              v = fmaf(kk.k[t], T1[(yy - (ys0 - hs)) * 2 * TOW + j], v);
            }
          } else {
            v = S2[(y - (ys0 - hs)) * S_P + (j >> 1) * 4 + (j & 1) + hs];
          }
          acc += 0.25f * v;
        }
      }
      out[((int64_t)f * oh + oy) * ow + ox] = acc / 255.0f;
    }
  }
}

}  // namespace dgvit

using namespace dgvit;

extern "C" {

int dgvit_depth_scratch_bytes(int n, int H, int W, size_t* bytes) {
  return guarded([&] {
    DG_REQUIRE(bytes && n >= 1 && H >= 8 && W >= 8, "bad argument");
    Carver cv(nullptr, 0, true);
    cv.take<float>((size_t)n * MM_BLOCKS * 2);
    cv.take<float>((size_t)n * H * W);
    cv.take<float>((size_t)n * H * W);
    cv.take<float>((size_t)n * (H / 5) * W);
    *bytes = cv.off;
  });
}

int dgvit_depth_augment(const float* raw, const float* noise, const uint64_t* rng_state, int n, int H, int W,
                        float* out, void* scratch, size_t scratch_bytes, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(raw);
    DG_REQUIRE(raw && out && scratch && n >= 1, "null argument");
    DG_REQUIRE(noise || rng_state, "provide noise or rng_state");
    DG_REQUIRE(H % 4 == 0 && W % 4 == 0, "H and W must be multiples of 4");
    DG_REQUIRE((((uintptr_t)raw) & 15) == 0 && (((uintptr_t)noise) & 15) == 0, "raw / noise must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(scratch, scratch_bytes);
    float* mm = cv.take<float>((size_t)n * MM_BLOCKS * 2);
    float* S1 = cv.take<float>((size_t)n * H * W);
    float* S2 = cv.take<float>((size_t)n * H * W);
    const int bh = H / 5, y1 = H / 2 - bh / 2;   // get_center_band, env_lab.py:33-39
    float* T1 = cv.take<float>((size_t)n * bh * W);
    K11 kk;
    {  // cv2.getGaussianKernel(11, sigma<=0): sigma = 0.3*((11-1)*0.5-1)+0.8 = 2.0
      const double sigma = 0.3 * ((11 - 1) * 0.5 - 1) + 0.8;
      double k[11], sum = 0;
      for (int i = 0; i < 11; ++i) { const double x = i - 5.0; k[i] = exp(-(x * x) / (2 * sigma * sigma)); sum += k[i]; }
      for (int i = 0; i < 11; ++i) kk.k[i] = (float)(k[i] / sum);
    }
    launch_k(depth_minmax_kernel, dim3(MM_BLOCKS, n), 256, 0, st, raw, mm, (int64_t)H * W);
    DG_LAUNCH_CHECK();
    static DevOnce attr;
    if (attr.first()) {
      DG_CUDA(cudaFuncSetAttribute(depth_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM));
    }
    (void)S1; (void)S2; (void)T1;
    launch_k(depth_fused_kernel, dim3((unsigned)cdiv(W / 4, TOW), (unsigned)cdiv(H / 4, TOH), n), 256, FUSED_SMEM, st, raw,
             noise, rng_state, mm, kk, out, H, W, y1, bh);
    DG_LAUNCH_CHECK();
  });
}

}  // extern "C"
