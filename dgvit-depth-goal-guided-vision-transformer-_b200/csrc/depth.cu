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
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
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

__device__ __forceinline__ float noisy_px(const float* __restrict__ raw, const float* __restrict__ noise,
                                          const uint64_t* rng, int64_t gi, double scale, double shift) {
  // cv2.normalize(NORM_MINMAX, 0..255) then .astype(uint8) (truncation)
  const float nrm = (float)((double)raw[gi] * scale + shift);
  const float u8 = (float)(int)fminf(fmaxf(nrm, 0.f), 255.f);
  float nz;
  if (noise) {
    nz = noise[gi];
  } else {
    uint32_t r[4];
    philox4x32(rng[0], (uint64_t)gi, 0x6465707468000000ull | (rng[1] & 0xffffffffu), r);
    nz = 50.0f * sqrtf(-2.0f * logf(u01(r[0]))) * cospif(2.0f * u01(r[1]));
  }
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
    DG_REQUIRE(raw && out && scratch && n >= 1, "null argument");
    DG_REQUIRE(noise || rng_state, "provide noise or rng_state");
    DG_REQUIRE(H % 4 == 0 && W % 4 == 0, "H and W must be multiples of 4");
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
    const int xb = (int)cdiv(W, 256);
    launch_k(depth_noise_hblur_kernel, dim3(xb, H, n), 256, 0, st, raw, noise, rng_state, mm, S1, H, W);
    DG_LAUNCH_CHECK();
    launch_k(depth_vblur5_kernel, dim3(xb, H, n), 256, 0, st, S1, S2, H, W);
    DG_LAUNCH_CHECK();
    launch_k(depth_band_hblur_kernel, dim3(xb, bh, n), 256, 0, st, S2, T1, kk, H, W, y1, bh);
    DG_LAUNCH_CHECK();
    const int fac = 4;
    launch_k(depth_resize_kernel, dim3((unsigned)cdiv(W / fac, 128), H / fac, n), 128, 0, st, S2, T1, kk, out, H, W, y1, bh, fac);
    DG_LAUNCH_CHECK();
  });
}

}  // extern "C"
