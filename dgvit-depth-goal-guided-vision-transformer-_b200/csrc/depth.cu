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
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw4; i += 4 * stride) {      // four loads in flight
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = i + u * stride;
      v[u] = reinterpret_cast<const float4*>(src)[j < hw4 ? j : i];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      mn = fminf(fminf(mn, v[u].x), fminf(v[u].y, fminf(v[u].z, v[u].w)));
      mx = fmaxf(fmaxf(mx, v[u].x), fmaxf(v[u].y, fmaxf(v[u].z, v[u].w)));
    }
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

// uniform in (0,1] from the top 23 bits, by bit assembly (no integer -> float conversion instruction)
__device__ __forceinline__ float u01d(uint32_t x) { return 2.0f - __uint_as_float((x >> 9) | 0x3f800000u); }

// N(0,1) draw of pixel gi: one Philox4x32 call serves the 4 pixels of an aligned group (two Box-Muller pairs)
__device__ __forceinline__ float normal_at(const uint64_t* rng, int64_t gi) {
  uint32_t r[4];
  philox4x32(rng[0], (uint64_t)(gi >> 2), 0x6465707468000000ull | (rng[1] & 0xffffffffu), r);
  const int q = (int)(gi & 3);
  const float rad = sqrtf(-2.0f * __logf(u01d(r[q & 2])));
  float sn, cs;
  __sincosf(6.283185307179586f * u01d(r[(q & 2) + 1]), &sn, &cs);
  return rad * ((q & 1) ? sn : cs);
}
struct K11 { float k[11]; };

// Four N(0,1) draws of the aligned pixel group gi4 = gi >> 2: ONE Philox4x32 call and two Box-Muller pairs (the per-pixel
// normal_at() above evaluates the same numbers; it recomputes the call for each of the 4 pixels)
__device__ __forceinline__ float4 normal4_at(const uint64_t* rng, int64_t gi4) {
  uint32_t r[4];
  philox4x32(rng[0], (uint64_t)gi4, 0x6465707468000000ull | (rng[1] & 0xffffffffu), r);
  const float rad0 = sqrtf(-2.0f * __logf(u01d(r[0]))), rad1 = sqrtf(-2.0f * __logf(u01d(r[2])));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u01d(r[1]), &s0, &c0);
  __sincosf(6.283185307179586f * u01d(r[3]), &s1, &c1);
  return make_float4(rad0 * c0, rad0 * s0, rad1 * c1, rad1 * s1);
}
// per-frame normalisation: cv2.normalize(NORM_MINMAX, 0..255) computes dst = src * scale + shift in double
struct Norm { double scale, shift; float mn, sf; };
__device__ __noinline__ float u8_exact(float raw, double scale, double shift) {      // (a real call: must not be if-converted)
  const float nrm = (float)((double)raw * scale + shift);
  return (float)(int)fminf(fmaxf(nrm, 0.f), 255.f);
}
__device__ __forceinline__ float noisy1(float raw, float nz, const Norm& nm) {
  // u8 = trunc((float)(raw * scale + shift)) as the reference forms it (.astype(uint8)).  Fast path in float, without a
  // single conversion instruction: (raw - min) * scale is within 5e-5 of the double result and lies in [0, 255.0001] because
  // min / max come from the same frame; trunc by a round-toward-zero add of 2^23.  Whenever that value is within 1e-4 of an
  // integer (2e-4 of the pixels; also NaN / out-of-range input) the exact double path decides.
  const float f = (raw - nm.mn) * nm.sf;
  float t = __fadd_rz(f, 8388608.f) - 8388608.f;
  if (!(fabsf((f - t) - 0.5f) <= 0.4999f)) t = u8_exact(raw, nm.scale, nm.shift);
  return fminf(fmaxf(t + nz, 0.f), 255.f);
}
// four pixels at once: one (rare) branch per group instead of one per pixel
__device__ __noinline__ float4 u8_exact4(float4 rv, double scale, double shift) {
  float4 t;
  t.x = u8_exact(rv.x, scale, shift); t.y = u8_exact(rv.y, scale, shift);
  t.z = u8_exact(rv.z, scale, shift); t.w = u8_exact(rv.w, scale, shift);
  return t;
}
// packed fp32 pairs (sm_100 FADD2 / FMUL2: one issue slot for two lanes of arithmetic; these kernels are issue-bound)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2_rz(f32x2 a, f32x2 b) { f32x2 r; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

__device__ __forceinline__ float4 noisy4(const float4& rv, const float4& nz, const Norm& nm) {
  const f32x2 mn2 = pk2(nm.mn, nm.mn), sf2 = pk2(nm.sf, nm.sf), big = pk2(8388608.f, 8388608.f), half = pk2(0.5f, 0.5f);
  const f32x2 f01 = mul2(sub2(pk2(rv.x, rv.y), mn2), sf2), f23 = mul2(sub2(pk2(rv.z, rv.w), mn2), sf2);
  const f32x2 t01 = sub2(add2_rz(f01, big), big), t23 = sub2(add2_rz(f23, big), big);      // trunc (0 <= f < 2^23)
  float d0, d1, d2, d3;
  unpk2(sub2(sub2(f01, t01), half), d0, d1);
  unpk2(sub2(sub2(f23, t23), half), d2, d3);
  const float c = 0.4999f;
  const bool ok = (fabsf(d0) <= c) & (fabsf(d1) <= c) & (fabsf(d2) <= c) & (fabsf(d3) <= c);
  float4 o;
  if (ok) {
    unpk2(add2(t01, pk2(nz.x, nz.y)), o.x, o.y);
    unpk2(add2(t23, pk2(nz.z, nz.w)), o.z, o.w);
  } else {
    const float4 t = u8_exact4(rv, nm.scale, nm.shift);
    o = make_float4(t.x + nz.x, t.y + nz.y, t.z + nz.z, t.w + nz.w);
  }
  return make_float4(fminf(fmaxf(o.x, 0.f), 255.f), fminf(fmaxf(o.y, 0.f), 255.f), fminf(fmaxf(o.z, 0.f), 255.f),
                     fminf(fmaxf(o.w, 0.f), 255.f));
}
// the noisy image A at the aligned 4-pixel group starting at column x4 of row gy (x4 may lie outside the image: every
// pixel is then taken at its BORDER_REFLECT_101 mirror, noise included: the blur mirrors the NOISY image)
__device__ __forceinline__ float4 noisy_group(const float* __restrict__ raw, const float* __restrict__ noise, const uint64_t* rng,
                                              int64_t fbase, int gy, int x4, int W, const Norm& nm) {
  const int64_t rb = fbase + (int64_t)gy * W;
  if (x4 >= 0 && x4 + 3 < W) {
    const float4 rv = *reinterpret_cast<const float4*>(raw + rb + x4);
    float4 nz;
    if (noise) nz = *reinterpret_cast<const float4*>(noise + rb + x4);
    else { nz = normal4_at(rng, (rb + x4) >> 2); nz.x *= 50.f; nz.y *= 50.f; nz.z *= 50.f; nz.w *= 50.f; }
    return noisy4(rv, nz, nm);
  }
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t gi = rb + reflect101(x4 + k, W);
    o[k] = noisy1(raw[gi], noise ? noise[gi] : 50.0f * normal_at(rng, gi), nm);
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

__device__ __forceinline__ void depth_tile(const float* __restrict__ raw, const float* __restrict__ noise, const uint64_t* rng,
                                           const K11& kk, float* __restrict__ out, int H, int W, int y1, int bh, int f, int oy0,
                                           int ox0, int oy_end, const Norm& nm, float* smf) {
  float* A = smf;                       // [A_R][A_P]
  float* Bh = A + A_R * A_P;            // [A_R][S_P]
  float* S2 = A;                        // [S_R][S_P]     (A is dead once Bh exists)
  float* T1 = Bh;                       // [S_R][2*TOW]   (Bh is dead once S2 exists)
  const int ohf = H / 4, oh = min(ohf, oy_end), ow = W / 4;      // rows [.., oy_end) belong to this launch's tile rows
  const int tid = threadIdx.x;
  const int ys0 = oy0 * 4 + 1, xs0 = ox0 * 4 + 1;           // first sampled row / col of the tile
  const int y2 = y1 + bh;
  const bool band = (ys0 + CORE_R - 1 >= y1) && (ys0 < y2);  // some sampled row lies in the centre band
  const int hs = band ? 5 : 0;                                // halo of S2, A needs hs + 2
  const int ar = CORE_R + 2 * (hs + 2), ac = CORE_C + 2 * (hs + 2);
  const int sr = CORE_R + 2 * hs, scn = CORE_C + 2 * hs;
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
          o = noisy4(rv[u], nz, nm);
        } else {
          o = noisy_group(raw, noise, rng, fbase, gys[u], x4, W, nm);      // image border: mirrored pixel by pixel
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
      out[((int64_t)f * ohf + oy) * ow + ox] = acc / 255.0f;
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
      out[((int64_t)f * ohf + oy) * ow + ox] = acc / 255.0f;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Streaming role (every output row with no sample in the centre band: ~80 % of a frame).  Off the band the result is the
// separable stride-4 filter w6 = [1 5 10 10 5 1]/32 over pixels 4o-1 .. 4o+4 of the noisy image in both directions
// (GaussianBlur(5,5) sampled at 4o+1, 4o+2 and averaged = cv2.resize by 4), so nothing needs shared memory:
//   lane = one aligned 4-pixel group (one 16-byte load of raw, one of noise / one Philox call per row), its left and right
//   neighbour pixels arrive by shuffle from the adjacent lanes: lanes 1..30 of a warp own output columns, lanes 0 and 31
//   only carry the neighbours' pixels (column -1 mirrors to 1, column W to W-2: both inside the lane's own group);
//   the warp walks down a strip of STRIP output rows, 4 input rows per step; each lane's 16-byte pieces of the next two
//   steps are in flight as cp.async copies into the lane's own slots of a small shared-memory ring (no registers held
//   across the wait, no barrier: a lane reads back only what it copied); rows 4o+3, 4o+4 serve outputs o and o+1.
// Every pixel of the strip is read and turned into a noisy value once (+ 2 halo rows per strip and 2 of 32 lanes).
constexpr int SEG = 30;
#ifndef STREAM_BPS
#define STREAM_BPS 3
#endif

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

#ifndef DRING
#define DRING 2
#endif
constexpr int RING = DRING;                                        // steps (4 rows each) in flight per warp
template <bool NOISE> constexpr int stream_smem() { return 8 * RING * 4 * (NOISE ? 2 : 1) * 32 * 16; }   // 8 warps per block

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool NOISE>
__device__ __forceinline__ void depth_stream(const float* __restrict__ raw, const float* __restrict__ noise, const uint64_t* rng,
                                             float* __restrict__ out, int H, int W, int f, int oy0, int oy1, int seg,
                                             const Norm& nm, float4* ring) {
  const int lane = threadIdx.x & 31;
  const int ow = W >> 2, ohf = H >> 2;
  const int g = seg * SEG - 1 + lane;                   // this lane's pixel group = its output column
  const bool writer = lane >= 1 && lane <= SEG && g < ow;
  const int gc = min(max(g, 0), ow - 1);                // (lanes outside the row load a valid group; their values are not used)
  const int64_t fbase = (int64_t)f * H * W;
  const float w0 = 1.f / 32, w1 = 5.f / 32, w2 = 10.f / 32;
  constexpr int ARR = NOISE ? 2 : 1;
  auto row_of = [&](int y) { return y < 0 ? -y : (y >= H ? 2 * (H - 1) - y : y); };      // BORDER_REFLECT_101 (one fold)
  const float* rbase = raw + fbase + 4 * gc;
  const float* nbase = NOISE ? noise + fbase + 4 * gc : nullptr;
  auto slot = [&](int step, int t, int arr) { return ring + (((step % RING) * 4 + t) * ARR + arr) * 32 + lane; };
  // every lane copies its own 16 bytes and reads back only those: the copies need no barrier, only wait_group
  auto issue = [&](int step, int nsteps) {
    if (step < nsteps) {
      const int y = 4 * (oy0 + step) - 1;
      if (y >= 0 && y + 3 < H) {                         // (warp-uniform) rows y .. y+3 are consecutive: one offset, three adds
        const float* rp = rbase + (int64_t)y * W;
        const float* np = nbase + (int64_t)y * W;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          cp_async16(slot(step, t, 0), rp + t * W);
          if (NOISE) cp_async16(slot(step, t, 1), np + t * W);
        }
      } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int off = row_of(y + t) * W;           // (a frame has < 2^31 pixels; the last step uses two of its rows)
          cp_async16(slot(step, t, 0), rbase + off);
          if (NOISE) cp_async16(slot(step, t, 1), nbase + off);
        }
      }
    }
    cp_async_commit();
  };
  auto hrow = [&](int step, int t, int y) {
    const float4 rv = *slot(step, t, 0);
    float4 nz;
    if (NOISE) {
      nz = *slot(step, t, 1);
    } else {
      nz = normal4_at(rng, (fbase + (int64_t)row_of(y) * W + 4 * gc) >> 2);
      nz.x *= 50.f; nz.y *= 50.f; nz.z *= 50.f; nz.w *= 50.f;
    }
    const float4 a = noisy4(rv, nz, nm);
    float left = __shfl_up_sync(0xffffffffu, a.w, 1), right = __shfl_down_sync(0xffffffffu, a.x, 1);
    if (g == 0) left = a.y;
    if (g == ow - 1) right = a.z;
    return w0 * (left + right) + w1 * (a.x + a.w) + w2 * (a.y + a.z);
  };
  const int n_out = oy1 - oy0, nsteps = n_out + 1;
#pragma unroll
  for (int i = 0; i < RING - 1; ++i) issue(i, nsteps);
  float carry = 0.f;
  for (int i = 0; i < nsteps; ++i) {
    issue(i + RING - 1, nsteps);
    cp_async_wait<RING - 1>();                           // step i has landed
    const int y = 4 * (oy0 + i) - 1;
    const float h0 = hrow(i, 0, y), h1 = hrow(i, 1, y + 1);
    if (i >= 1 && writer) out[((int64_t)f * ohf + oy0 + i - 1) * ow + g] = (carry + w1 * h0 + w0 * h1) / 255.0f;
    if (i < n_out) {
      const float h2 = hrow(i, 2, y + 2), h3 = hrow(i, 3, y + 3);
      carry = w0 * h0 + w1 * h1 + w2 * (h2 + h3);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Band role (output rows with a sample in the centre band).  Same register streaming, one warp per (frame, 28-column
// segment) walking down ALL rows the band needs.  Horizontally everything the band does is one composite filter per
// output column: c16 = k5 * (k11 at column 4o+1  +  k11 at column 4o+2)/2 over pixels 4o-6 .. 4o+9 (own group, the two
// neighbours, half of the next two: 12 shuffles); image-border columns mirror exactly because every filter is symmetric
// (lanes outside the row hold the mirrored pixels).  Vertically: 5-tap sliding window in registers -> the blurred row goes
// into this lane's private shared-memory column; the 11-tap with BORDER_REFLECT_101 INSIDE the band reads that column
// back.  An output row with one sample outside the band takes that sample from the plain w6 path (two registers).
constexpr int BSEG = 28;
#ifndef BAND_BPS
#define BAND_BPS 3
#endif
struct BandGeo {
  int H, W, y1, bh, first, last, segs;      // output rows first..last touch the band
  int64_t items;                            // n * segs
};
struct StreamGeo {
  int H, W, ob0, ob1, strip, strips_lo, strips, segs;   // output rows [ob0, ob1) are not streamed
  int64_t items;
};

__device__ __forceinline__ void frame_scale(const float* __restrict__ mmpart, int f, int lane, Norm& nm) {
  static_assert(MM_BLOCKS == 64, "two partials per lane");
  const float2 p0 = *reinterpret_cast<const float2*>(mmpart + ((int64_t)f * MM_BLOCKS + lane) * 2);
  const float2 p1 = *reinterpret_cast<const float2*>(mmpart + ((int64_t)f * MM_BLOCKS + 32 + lane) * 2);
  float mn = fminf(p0.x, p1.x), mx = fmaxf(p0.y, p1.y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const double rg = (double)mx - (double)mn;
  nm.scale = rg > 2.220446049250313e-16 ? 255.0 / rg : 0.0;
  nm.shift = 0.0 - (double)mn * nm.scale;
  nm.mn = mn;
  nm.sf = (float)nm.scale;
}

template <bool NOISE>
__global__ void __launch_bounds__(256, BAND_BPS) depth_band_kernel(const float* __restrict__ raw, const float* __restrict__ noise,
                                                            const uint64_t* rng, const float* __restrict__ mmpart,
                                                            const __grid_constant__ K11 kk, float* __restrict__ out,
                                                            const __grid_constant__ BandGeo gm) {
  pdl_wait();
  pdl_launch();
  extern __shared__ __align__(16) float smf[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = blockIdx.x / gm.segs, seg = blockIdx.x % gm.segs;
  Norm nm;
  frame_scale(mmpart, f, lane, nm);
  const int H = gm.H, W = gm.W, y1 = gm.y1, bh = gm.bh, y2 = y1 + bh;
  const int ow = W >> 2, ohf = H >> 2;
  const int g = seg * BSEG - 2 + lane, x4 = 4 * g;
  const bool inimg = x4 >= 0 && x4 + 3 < W;
  const bool writer = lane >= 2 && lane < 2 + BSEG && g < ow;
  // mirrored lanes (W >= 16): group g < 0 is (group -g).x, (group -g-1).w,.z,.y; group g >= ow is (group 2ow-g-1).z,.y,.x,
  // (group 2ow-g-2).w; both source groups sit in this warp (lane = group - seg * BSEG + 2)
  const int vmode = g < 0 ? 1 : (g >= ow && g <= ow + 1 ? 2 : 0);
  const int srcA = (vmode == 1 ? -g : 2 * ow - g - 1) - seg * BSEG + 2, srcB = srcA - 1;
  const bool border = seg == 0 || (seg + 1) * BSEG + 2 > ow;          // (warp-uniform) some lane of this warp is mirrored
  const int x4c = min(max(x4, 0), W - 4);
  const int64_t fbase = (int64_t)f * H * W;
  const float k5[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  float c16[16];
  {
    float h12[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) h12[j] = 0.5f * ((j < 11 ? kk.k[j] : 0.f) + (j >= 1 ? kk.k[j - 1] : 0.f));
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < 5; ++t)
        if (m - t >= 0 && m - t < 12) a = fmaf(k5[t], h12[m - t], a);
      c16[m] = a;
    }
  }
  const float w0 = 1.f / 32, w1 = 5.f / 32, w2 = 10.f / 32;
  // blurred rows to form: the band itself plus the (at most one each side) sampled row just outside it
  const int ylo = 4 * gm.first + 1, yhi = 4 * gm.last + 2;
  const int sa = min(y1, ylo), sb = max(y2 - 1, yhi);
  const int ra = sa - 2, rb = sb + 2, nrows = rb - ra + 1;           // noisy rows (y1 - 3 .. y2 + 2 at most)
  float* ah16 = smf + lane;                                          // [nrows][32]  horizontal c16 of noisy row ra + i
  float* s2 = ah16 + (size_t)nrows * 32;                             // [bh][32]     blurred band row y1 + i
  float* ah6 = s2 + (size_t)bh * 32;                                 // [10][32]     horizontal w6 of rows y1-3..y1+1, y2-2..y2+2
  // ---- phase 1: warp w takes the noisy rows ra + w, ra + w + 8, ...; four rows' loads in flight
  for (int r0 = ra + warp; r0 <= rb; r0 += 32) {
    float4 rv[4], nv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      rv[u] = nv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int r = r0 + 8 * u;
      if (r <= rb && inimg) {
        const int64_t gi = fbase + (int64_t)((r < 0 || r >= H) ? reflect101(r, H) : r) * W + x4;
        rv[u] = ld_stream4(raw + gi);
        if (NOISE) nv[u] = ld_stream4(noise + gi);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + 8 * u;
      if (r > rb) break;                                             // (warp-uniform)
      const int rr = (r < 0 || r >= H) ? reflect101(r, H) : r;
      float4 nz = nv[u];
      if (!NOISE) {
        nz = normal4_at(rng, (fbase + (int64_t)rr * W + x4c) >> 2);
        nz.x *= 50.f; nz.y *= 50.f; nz.z *= 50.f; nz.w *= 50.f;
      }
      float4 a = noisy4(rv[u], nz, nm);                              // (lanes outside the row: replaced below)
      const unsigned FULL = 0xffffffffu;
      if (border) {
        // lanes left / right of the image hold the BORDER_REFLECT_101 mirror of the noisy row: pixel -p is pixel p,
        // pixel W-1+p is pixel W-1-p; both sit in the in-image lanes next to them
        const float Ax = __shfl_sync(FULL, a.x, srcA), Ay = __shfl_sync(FULL, a.y, srcA), Az = __shfl_sync(FULL, a.z, srcA);
        const float Bw = __shfl_sync(FULL, a.w, srcB), Bz = __shfl_sync(FULL, a.z, srcB), By = __shfl_sync(FULL, a.y, srcB);
        if (vmode == 1) a = make_float4(Ax, Bw, Bz, By);
        else if (vmode == 2) a = make_float4(Az, Ay, Ax, Bw);
      }
      const float m2z = __shfl_up_sync(FULL, a.z, 2), m2w = __shfl_up_sync(FULL, a.w, 2);
      const float m1x = __shfl_up_sync(FULL, a.x, 1), m1y = __shfl_up_sync(FULL, a.y, 1);
      const float m1z = __shfl_up_sync(FULL, a.z, 1), m1w = __shfl_up_sync(FULL, a.w, 1);
      const float p1x = __shfl_down_sync(FULL, a.x, 1), p1y = __shfl_down_sync(FULL, a.y, 1);
      const float p1z = __shfl_down_sync(FULL, a.z, 1), p1w = __shfl_down_sync(FULL, a.w, 1);
      const float p2x = __shfl_down_sync(FULL, a.x, 2), p2y = __shfl_down_sync(FULL, a.y, 2);
      float h16 = c16[0] * m2z, h16b = c16[1] * m2w;                 // two chains
      h16 = fmaf(c16[2], m1x, h16); h16b = fmaf(c16[3], m1y, h16b);
      h16 = fmaf(c16[4], m1z, h16); h16b = fmaf(c16[5], m1w, h16b);
      h16 = fmaf(c16[6], a.x, h16); h16b = fmaf(c16[7], a.y, h16b);
      h16 = fmaf(c16[8], a.z, h16); h16b = fmaf(c16[9], a.w, h16b);
      h16 = fmaf(c16[10], p1x, h16); h16b = fmaf(c16[11], p1y, h16b);
      h16 = fmaf(c16[12], p1z, h16); h16b = fmaf(c16[13], p1w, h16b);
      h16 = fmaf(c16[14], p2x, h16); h16b = fmaf(c16[15], p2y, h16b);
      ah16[(r - ra) * 32] = h16 + h16b;
      const float h6 = w0 * (m1w + p1x) + w1 * (a.x + a.w) + w2 * (a.y + a.z);
      if (r >= y1 - 3 && r <= y1 + 1) ah6[(r - (y1 - 3)) * 32] = h6;
      if (r >= y2 - 2 && r <= y2 + 2) ah6[(5 + r - (y2 - 2)) * 32] = h6;
    }
  }
  __syncthreads();
  // ---- phase 2: vertical 5-tap -> the GaussianBlur(5,5) rows of the band (already filtered horizontally for the samples)
  for (int q = y1 + warp; q < y2; q += 8) {
    float sacc = 0.f;
#pragma unroll
    for (int t = 0; t < 5; ++t) sacc = fmaf(k5[t], ah16[(q - 2 + t - ra) * 32], sacc);
    s2[(q - y1) * 32] = sacc;
  }
  float s6_lo = 0.f, s6_hi = 0.f;       // plain-path samples at rows y1 - 1 and y2 (read only when an output row straddles the edge)
  if (ylo < y1) {
#pragma unroll
    for (int t = 0; t < 5; ++t) s6_lo = fmaf(k5[t], ah6[t * 32], s6_lo);
  }
  if (yhi >= y2) {
#pragma unroll
    for (int t = 0; t < 5; ++t) s6_hi = fmaf(k5[t], ah6[(5 + t) * 32], s6_hi);
  }
  __syncthreads();
  // ---- phase 3: 11-tap down the band column (BORDER_REFLECT_101 inside the band), 2 x 2 average
  if (!writer) return;
  for (int oy = gm.first + warp; oy <= gm.last; oy += 8) {
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int y = 4 * oy + 1 + dy;
      float v;
      if (y >= y1 && y < y2) {
        v = 0.f;
        const int yb0 = y - y1 - 5;
        if (yb0 >= 0 && yb0 + 10 < bh) {
#pragma unroll
          for (int t = 0; t < 11; ++t) v = fmaf(kk.k[t], s2[(yb0 + t) * 32], v);
        } else {
#pragma unroll
          for (int t = 0; t < 11; ++t) v = fmaf(kk.k[t], s2[reflect101(yb0 + t, bh) * 32], v);     // reflect inside the band
        }
      } else {
        v = y < y1 ? s6_lo : s6_hi;
      }
      acc += 0.5f * v;
    }
    out[((int64_t)f * ohf + oy) * ow + g] = acc / 255.0f;
  }
}

// streaming role over the rows outside [ob0, ob1): one (frame, strip, 30-column segment) per warp
template <bool NOISE>
__global__ void __launch_bounds__(256, STREAM_BPS) depth_stream_kernel(const float* __restrict__ raw, const float* __restrict__ noise,
                                                           const uint64_t* rng, const float* __restrict__ mmpart,
                                                           float* __restrict__ out, const __grid_constant__ StreamGeo gm) {
  pdl_wait();
  pdl_launch();
  const int64_t item = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (item >= gm.items) return;
  const int f = (int)(item / ((int64_t)gm.strips * gm.segs));
  Norm nm;
  frame_scale(mmpart, f, threadIdx.x & 31, nm);
  const int seg = (int)(item % gm.segs);
  const int sidx = (int)((item / gm.segs) % gm.strips);
  int oy0, oy1;
  if (sidx < gm.strips_lo) { oy0 = sidx * gm.strip; oy1 = min(oy0 + gm.strip, gm.ob0); }
  else { oy0 = gm.ob1 + (sidx - gm.strips_lo) * gm.strip; oy1 = min(oy0 + gm.strip, gm.H / 4); }
  extern __shared__ __align__(16) float4 ring4[];
  depth_stream<NOISE>(raw, noise, rng, out, gm.H, gm.W, f, oy0, oy1, seg, nm,
                      ring4 + (threadIdx.x >> 5) * (RING * 4 * (NOISE ? 2 : 1) * 32));
}

// every row through the shared-memory tiles (dgvit_set_option("depth_strip", 0), and frames whose band column does not fit
// into shared memory): the first fused version, kept as the cross-check of the two streaming roles
__global__ void __launch_bounds__(256) depth_tile_kernel(const float* __restrict__ raw, const float* __restrict__ noise,
                                                         const uint64_t* rng, const float* __restrict__ mmpart,
                                                         const __grid_constant__ K11 kk, float* __restrict__ out, int H, int W,
                                                         int y1, int bh) {
  pdl_wait();
  pdl_launch();
  extern __shared__ __align__(16) float smf[];
  __shared__ Norm snm;
  const int f = blockIdx.z;
  if (threadIdx.x < 32) {
    Norm nm;
    frame_scale(mmpart, f, threadIdx.x, nm);
    if (threadIdx.x == 0) snm = nm;
  }
  __syncthreads();
  const Norm nm = snm;
  depth_tile(raw, noise, rng, kk, out, H, W, y1, bh, f, blockIdx.y * TOH, blockIdx.x * TOW, H / 4, nm, smf);
}

}  // namespace dgvit

using namespace dgvit;

namespace {
struct DepthSide { cudaStream_t s; cudaEvent_t fork, join; };
DepthSide& depth_side() {       // one second stream per device (fork / join by events: capturable)
  static DepthSide per_dev[64];
  static bool inited[64] = {};
  const int dev = current_device();
  if (!inited[dev]) {
    DG_CUDA(cudaStreamCreateWithFlags(&per_dev[dev].s, cudaStreamNonBlocking));
    DG_CUDA(cudaEventCreateWithFlags(&per_dev[dev].fork, cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&per_dev[dev].join, cudaEventDisableTiming));
    inited[dev] = true;
  }
  return per_dev[dev];
}
}  // namespace

extern "C" {

int g_depth_skip = 0;           // debug (profiles/depth_bench.py): 1 = no band kernel, 2 = no streaming kernel, 4 = no min/max pass
int g_depth_strip = -1;         // dgvit_set_option("depth_strip"): output rows per streaming strip; < 0 = chosen per call, 0 = every row through the tiles

int dgvit_depth_scratch_bytes(int n, int H, int W, size_t* bytes) {
  return guarded([&] {
    DG_REQUIRE(bytes && n >= 1 && H >= 8 && W >= 8, "bad argument");
    Carver cv(nullptr, 0, true);
    cv.take<float>((size_t)n * MM_BLOCKS * 2);
    *bytes = cv.off;
  });
}

int dgvit_depth_augment(const float* raw, const float* noise, const uint64_t* rng_state, int n, int H, int W,
                        float* out, void* scratch, size_t scratch_bytes, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(raw);
    DG_REQUIRE(raw && out && scratch && n >= 1, "null argument");
    DG_REQUIRE(noise || rng_state, "provide noise or rng_state");
    DG_REQUIRE(H >= 8 && W >= 8 && H % 4 == 0 && W % 4 == 0, "H and W must be multiples of 4 (>= 8)");
    DG_REQUIRE((((uintptr_t)raw) & 15) == 0 && (((uintptr_t)noise) & 15) == 0, "raw / noise must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(scratch, scratch_bytes);
    float* mm = cv.take<float>((size_t)n * MM_BLOCKS * 2);
    K11 kk;
    {  // cv2.getGaussianKernel(11, sigma<=0): sigma = 0.3*((11-1)*0.5-1)+0.8 = 2.0
      const double sigma = 0.3 * ((11 - 1) * 0.5 - 1) + 0.8;
      double k[11], sum = 0;
      for (int i = 0; i < 11; ++i) { const double x = i - 5.0; k[i] = exp(-(x * x) / (2 * sigma * sigma)); sum += k[i]; }
      for (int i = 0; i < 11; ++i) kk.k[i] = (float)(k[i] / sum);
    }
    if (!(g_depth_skip & 4)) {
      launch_k(depth_minmax_kernel, dim3(MM_BLOCKS, n), 256, 0, st, raw, mm, (int64_t)H * W);
      DG_LAUNCH_CHECK();
    }
    const int bh = H / 5, y1 = H / 2 - bh / 2, y2 = y1 + bh;     // get_center_band, env_lab.py:33-39
    const int oh = H / 4, ow = W / 4;
    // output row oy samples rows 4oy+1, 4oy+2: it touches the band when one of them lies in [y1, y2)
    int first = y1 - 2 <= 0 ? 0 : (y1 - 2 + 3) / 4, last = std::min(oh - 1, (y2 - 2) / 4);
    if (bh < 1 || last < first) { first = 0; last = -1; }
    const size_t band_smem = (size_t)(2 * bh + 16) * 32 * sizeof(float);     // ah16 [<= bh + 6] + s2 [bh] + ah6 [10] rows of 32
    static DevOnce attr;
    if (attr.first()) {
      DG_CUDA(cudaFuncSetAttribute(depth_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM));
      DG_CUDA(cudaFuncSetAttribute(depth_band_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      DG_CUDA(cudaFuncSetAttribute(depth_band_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      DG_CUDA(cudaFuncSetAttribute(depth_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, stream_smem<true>()));
      DG_CUDA(cudaFuncSetAttribute(depth_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stream_smem<false>()));
    }
    if (g_depth_strip == 0 || band_smem > 200 * 1024 || W < 16) {
      launch_k(depth_tile_kernel, dim3((unsigned)cdiv(ow, TOW), (unsigned)cdiv(oh, TOH), n), 256, FUSED_SMEM, st, raw, noise,
               rng_state, (const float*)mm, kk, out, H, W, y1, bh);
      DG_LAUNCH_CHECK();
      return;
    }
    // the band rows (a few long-running warps) on a second stream beside the streamed rows
    DepthSide& sd = depth_side();
    const bool have_band = last >= first;
    if (have_band && !(g_depth_skip & 1)) {
      BandGeo bg;
      bg.H = H; bg.W = W; bg.y1 = y1; bg.bh = bh; bg.first = first; bg.last = last;
      bg.segs = (int)cdiv(ow, BSEG);
      bg.items = (int64_t)n * bg.segs;
      DG_CUDA(cudaEventRecord(sd.fork, st));
      DG_CUDA(cudaStreamWaitEvent(sd.s, sd.fork, 0));
      const dim3 grid((unsigned)bg.items);
      if (noise) launch_k(depth_band_kernel<true>, grid, 256, band_smem, sd.s, raw, noise, rng_state, (const float*)mm, kk, out, bg);
      else launch_k(depth_band_kernel<false>, grid, 256, band_smem, sd.s, raw, noise, rng_state, (const float*)mm, kk, out, bg);
      DG_LAUNCH_CHECK();
      DG_CUDA(cudaEventRecord(sd.join, sd.s));
    }
    StreamGeo gm;
    gm.H = H; gm.W = W;
    gm.ob0 = have_band ? first : 0;
    gm.ob1 = have_band ? last + 1 : 0;
    gm.segs = (int)cdiv(ow, SEG);
    auto n_strips = [&](int strip) { return cdiv(gm.ob0, strip) + cdiv(oh - gm.ob1, strip); };
    int strip = g_depth_strip;
    if (strip < 0) {
      // strip height against wave quantisation: a warp's time grows with its rows (4 per output row + 2 halo + start-up), the
      // launch takes whole waves of 148 SMs x STREAM_BPS blocks x 8 warps
      const int64_t capacity = 148 * STREAM_BPS * 8;
      const int rows_max = std::max(gm.ob0, oh - gm.ob1);
      double best = 1e300;
      for (int cand = 3; cand <= 32; ++cand) {
        const int64_t items = (int64_t)n * gm.segs * n_strips(cand);
        const double cost = (double)cdiv(items, capacity) * (4.0 * std::min(cand, std::max(rows_max, 1)) + 6.0);
        if (cost < best) { best = cost; strip = cand; }
      }
    }
    gm.strip = std::max(1, std::min(strip, 64));
    gm.strips_lo = (int)cdiv(gm.ob0, gm.strip);
    gm.strips = (int)n_strips(gm.strip);
    gm.items = (int64_t)n * gm.strips * gm.segs;
    if (gm.items > 0 && !(g_depth_skip & 2)) {
      const int64_t blocks = cdiv(gm.items, 8);
      DG_REQUIRE(blocks < (int64_t)1 << 31, "too many frames for one launch");
      if (noise) launch_k(depth_stream_kernel<true>, dim3((unsigned)blocks), 256, stream_smem<true>(), st, raw, noise, rng_state, (const float*)mm, out, gm);
      else launch_k(depth_stream_kernel<false>, dim3((unsigned)blocks), 256, stream_smem<false>(), st, raw, noise, rng_state, (const float*)mm, out, gm);
      DG_LAUNCH_CHECK();
    }
    if (have_band && !(g_depth_skip & 1)) DG_CUDA(cudaStreamWaitEvent(st, sd.join, 0));
  });
}

}  // extern "C"
