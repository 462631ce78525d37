// common.cuh — shared host/device helpers of libdgvit (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dgvit.h"

namespace dgvit {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- error plumbing
inline std::string& last_error() {
  static thread_local std::string e;
  return e;
}
struct Fail {
  int code;
};
inline void fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  throw Fail{code};
}
#define DG_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      cudaGetLastError(); /* do not leave it pending for a later, unrelated call */            \
      ::dgvit::fail(DGVIT_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,                 \
                    cudaGetErrorString(e__));                                                  \
    }                                                                                          \
  } while (0)
// every kernel launch of the library passes through here: error check + launch accounting
// (bench.py reports the count as "gpu_launches")
inline long long& launch_counter() {
  static long long n = 0;
  return n;
}
#define DG_LAUNCH_CHECK()          \
  do {                             \
    ++::dgvit::launch_counter();   \
    DG_CUDA(cudaGetLastError());   \
  } while (0)

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may be
// scheduled while the previous kernel of the stream is still draining, run their input-independent
// prologue, and block in pdl_wait() until the previous grid has completed and flushed.  pdl_launch()
// (issued right after the wait) lets the NEXT kernel do the same.  The ~370 kernels of one update are
// mostly single-wave and latency-bound, so hiding launch + prologue latency matters.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// measurement knob (set_option "skip"): bit mask of kernel families whose launches are dropped, to read their marginal
// cost inside the graph-replayed step off an A/B run (results are garbage with any bit set)
enum { SKIP_MLP_FWD = 1, SKIP_MLP_BWD_X = 2, SKIP_MLP_BWD_W = 4, SKIP_ATTN_FWD = 8, SKIP_ATTN_BWD = 16, SKIP_GEMM_TC = 32,
       SKIP_LN_BWD = 64, SKIP_REDUCE = 128, SKIP_HEADS = 256, SKIP_EMBED = 512 };
inline int& skip_mask() {
  static int m = 0;
  return m;
}
inline bool& pdl_enabled() {
  static bool e = true;
  return e;
}
// Programmatic dependent launch and the non-coherent load path.  A kernel launched with the programmatic attribute starts
// its life while the kernel before it in the stream is still running (it blocks in griddepcontrol.wait before touching data).
// `ld.global.nc` / __ldg promise "read-only for the lifetime of the kernel", which that early start formally breaks for
// anything the previous kernel writes.  No wrong value was ever traced to it (the compiler emits LDG.CONSTANT for most
// `const __restrict__` loads anyway and the update is bit-reproducible), but as a precaution: (1) the reductions read the
// partial sums of the launch right before them with __ldcg; (2) the launch right after a kernel that REWRITES parameters
// (Adam) -- and the first launch of every C call on each stream, whose predecessor belongs to the caller -- is an ordinary,
// fully serialised launch.
struct NoPdlOnce {
  cudaStream_t strict[8];      // streams whose next launch must be an ordinary one
  int n_strict = 0;
  unsigned epoch = 0;          // bumped at every C entry point
  cudaStream_t seen[16];
  unsigned seen_epoch[16];
  int n_seen = 0;
  void mark_strict(cudaStream_t st) {
    for (int i = 0; i < n_strict; ++i)
      if (strict[i] == st) return;
    if (n_strict < 8) strict[n_strict++] = st;
    else strict[0] = st;
  }
  bool take(cudaStream_t st) {
    bool hit = false;
    for (int i = 0; i < n_strict; ++i)
      if (strict[i] == st) { strict[i] = strict[--n_strict]; hit = true; break; }
    int k = -1;
    for (int i = 0; i < n_seen; ++i)
      if (seen[i] == st) { k = i; break; }
    if (k < 0) {
      k = n_seen < 16 ? n_seen++ : 0;
      seen[k] = st; seen_epoch[k] = epoch - 1;
    }
    if (seen_epoch[k] != epoch) { seen_epoch[k] = epoch; hit = true; }      // first launch of this C call on this stream
    return hit;
  }
};
inline NoPdlOnce& no_pdl_once() {
  static thread_local NoPdlOnce r;
  return r;
}
static inline bool pdl_for_launch(cudaStream_t st) {
  const bool strict = no_pdl_once().take(st);          // (always consumed)
  return pdl_enabled() && !strict;
}
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_for_launch(st) ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) {
    cudaGetLastError();      // a refused launch must not linger as the "last error" of a later, unrelated call
    fail(DGVIT_ERR_CUDA, "kernel launch refused (%s): grid (%u,%u,%u) block %u dynamic smem %zu", cudaGetErrorString(e), grid.x,
         grid.y, grid.z, block.x, smem);
  }
}

// same, as thread-block clusters of `cluster` CTAs along x
template <typename... KArgs, typename... Args>
static inline void launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_for_launch(st) ? 2 : 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fail(DGVIT_ERR_CUDA, "cluster kernel launch refused (%s): grid (%u,%u,%u) block %u dynamic smem %zu cluster %d",
         cudaGetErrorString(e), grid.x, grid.y, grid.z, block.x, smem, cluster);
  }
}

// optional per-kernel timing of the launches tagged as the dominant kernel (bench.py roofline):
// events are recorded on the launching stream around each tagged launch.
struct Prof {
  bool on = false;
  int tag = 0;               // which kernel family is being timed
  std::vector<cudaEvent_t> ev;   // pairs (start, stop)
  size_t used = 0;
  double flops = 0, bytes = 0;
  long long launches = 0;
};
inline Prof& prof() {
  static Prof p;
  return p;
}
enum { PROF_NONE = 0, PROF_GEMM_MLP = 1, PROF_GEMM_ALL = 2, PROF_ATTENTION = 3, PROF_GATHER = 4, PROF_ADAM = 5,
       PROF_MLP_FUSED = 6, PROF_LN_BWD = 7, PROF_EMBED = 8, PROF_PATCH = 9 };
struct ProfScope {
  bool active = false;
  cudaStream_t st;
  ProfScope(int tag, double flops, double bytes, cudaStream_t s) : st(s) {
    Prof& p = prof();
    if (!p.on || p.tag != tag || p.used + 2 > p.ev.size()) return;
    active = true;
    p.flops += flops; p.bytes += bytes; p.launches++;
    cudaEventRecord(p.ev[p.used], st);
  }
  ~ProfScope() {
    if (!active) return;
    Prof& p = prof();
    cudaEventRecord(p.ev[p.used + 1], st);
    p.used += 2;
  }
};
#define DG_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) ::dgvit::fail(DGVIT_ERR_ARG, __VA_ARGS__);    \
  } while (0)

// exceptions never cross the C ABI
template <typename F>
static inline int guarded(F&& f) {
  try {
    ++no_pdl_once().epoch;           // the first launch of a call on each stream is fully serialised behind whatever the caller queued
    f();
    return DGVIT_OK;
  } catch (const Fail& e) {
    return e.code;
  } catch (const std::exception& e) {
    last_error() = e.what();
    return DGVIT_ERR_ARG;
  } catch (...) {
    last_error() = "unknown C++ exception";
    return DGVIT_ERR_ARG;
  }
}

// ---------------------------------------------------------------- device plumbing
// Function attributes, streams and events belong to ONE device: one-time initialisation is keyed by the current device.
struct DevOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d & 63;
}
// Every entry point runs on the device that owns its buffers, whatever device is current in the calling thread
// (SAC(device="cuda:1") without torch.cuda.set_device): the guard switches to the device of `p` and back.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* p) {
    if (!p) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return;
    cudaGetDevice(&prev);
    if (at.device != prev) { cudaSetDevice(at.device); switched = true; }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------- workspace carving
struct Carver {
  char* base;
  size_t off, cap;
  bool dry;  // only measure
  Carver(void* p, size_t cap_, bool dry_ = false) : base((char*)p), off(0), cap(cap_), dry(dry_) {}
  template <typename T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    size_t o = off;
    off += bytes;
    if (dry) return nullptr;
    if (off > cap) fail(DGVIT_ERR_WORKSPACE, "workspace too small: need >= %zu, have %zu", off, cap);
    return (T*)(base + o);
  }
};

__host__ __device__ static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU and its derivative (nn.GELU default, vn/GoalFormer.py:44)
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Branch-free GELU for the bf16 tensor-core epilogues (Abramowitz-Stegun 7.1.26 erf, |err| <= 1.5e-7,
// two MUFU ops: rcp + ex2).  exp(-x^2/2) is shared between erf and the Gaussian pdf of the derivative.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_fast(float x, float& y, float& dy) {
  const float t = rcp_approx(fmaf(0.23164189f, fabsf(x), 1.0f));        // p/sqrt(2) = 0.3275911/1.41421356
  const float e = ex2_approx(-0.72134752f * x * x);                     // exp(-x^2/2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float erfa = fmaf(-poly * t, e, 1.0f);                          // erf(|x|/sqrt2)
  const float cdf = fmaf(0.5f, copysignf(erfa, x), 0.5f);
  y = x * cdf;
  dy = fmaf(x * 0.39894228f, e, cdf);
}

// 3-term Abramowitz-Stegun erf (7.1.25, |err| <= 2.5e-5: far below bf16 resolution of the outputs)
__device__ __forceinline__ float gelu3(float x) {
  const float t = rcp_approx(fmaf(0.33267264f, fabsf(x), 1.0f));
  const float e = ex2_approx(-0.72134752f * x * x);
  float poly = fmaf(t, 0.7478556f, -0.0958798f);
  poly = fmaf(t, poly, 0.3480242f);
  const float erfa = fmaf(-poly * t, e, 1.0f);
  return x * fmaf(0.5f, copysignf(erfa, x), 0.5f);
}
__device__ __forceinline__ void gelu3_grad(float x, float& y, float& dy) {
  const float t = rcp_approx(fmaf(0.33267264f, fabsf(x), 1.0f));
  const float e = ex2_approx(-0.72134752f * x * x);
  float poly = fmaf(t, 0.7478556f, -0.0958798f);
  poly = fmaf(t, poly, 0.3480242f);
  const float erfa = fmaf(-poly * t, e, 1.0f);
  const float cdf = fmaf(0.5f, copysignf(erfa, x), 0.5f);
  y = x * cdf;
  dy = fmaf(x * 0.39894228f, e, cdf);
}

// ---------------------------------------------------------------- counter-based RNG
// Philox4x32-10; key = seed, counter = (index, stream_id, update counter).
struct Philox {
  uint32_t c[4];
  uint32_t k[2];
};
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__host__ __device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi,
                                                    uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
__host__ __device__ __forceinline__ float u01(uint32_t x) {  // (0,1]
  return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// Dropout keep decision shared by forward and backward.
struct DropDev {
  int mode;
  float p, scale;
  const uint8_t* mask;
  const uint64_t* rng;
  uint32_t stream_id;
  int64_t elem_offset;  // global element index of local element 0 (data-parallel slicing)
};
__device__ __forceinline__ float drop_factor(const DropDev& d, int64_t i) {
  if (d.mode == DGVIT_DROP_NONE) return 1.0f;
  if (d.mode == DGVIT_DROP_MASK) return d.mask[i] ? d.scale : 0.0f;
  uint32_t o[4];
  const uint64_t gi = (uint64_t)(i + d.elem_offset);
  philox4x32(d.rng[0], gi >> 2, ((uint64_t)d.stream_id << 32) | (d.rng[1] & 0xffffffffu), o);
  return u01(o[gi & 3]) > d.p ? d.scale : 0.0f;
}
// dropout factors of 4 consecutive elements starting at the 4-aligned element index i (one Philox call in RNG mode)
__device__ __forceinline__ float4 drop_factor4(const DropDev& d, int64_t i) {
  if (d.mode == DGVIT_DROP_NONE) return make_float4(1.f, 1.f, 1.f, 1.f);
  if (d.mode == DGVIT_DROP_MASK) {
    const uchar4 m = *reinterpret_cast<const uchar4*>(d.mask + i);
    return make_float4(m.x ? d.scale : 0.f, m.y ? d.scale : 0.f, m.z ? d.scale : 0.f, m.w ? d.scale : 0.f);
  }
  uint32_t o[4];
  const uint64_t gi = (uint64_t)(i + d.elem_offset);        // elem_offset is a multiple of N * D, hence of 4
  philox4x32(d.rng[0], gi >> 2, ((uint64_t)d.stream_id << 32) | (d.rng[1] & 0xffffffffu), o);
  return make_float4(u01(o[0]) > d.p ? d.scale : 0.f, u01(o[1]) > d.p ? d.scale : 0.f, u01(o[2]) > d.p ? d.scale : 0.f,
                     u01(o[3]) > d.p ? d.scale : 0.f);
}

}  // namespace dgvit
