// gemm_tc.cuh — bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled
// shared memory -> tcgen05.mma (fp32 accumulators in TMEM) -> tcgen05.ld epilogue.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2..9 = epilogue (two warps per TMEM lane quadrant, each owning half of the columns).
// TMEM accumulators are double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// One kernel serves the three contractions of the hot path through the operand "major":
//   y  = x W^T      A K-major,  B K-major      (forward linears)
//   dx = dy W       A K-major,  B MN-major     (input gradients)
//   dW = dy^T x     A MN-major, B MN-major     (weight gradients, split-K over tokens)
#pragma once
#include <cuda.h>

#include <map>
#include <tuple>

#include "common.cuh"
#include "gemm_simt.cuh"

namespace dgvit {
namespace tc {

constexpr int BM = 128;   // UMMA M (cta_group::1)
constexpr int BK = 64;    // 64 bf16 = one 128-byte swizzle span
constexpr int UK = 16;    // UMMA K for 16-bit inputs
// epilogue warps: 4 TMEM lane quadrants x (BN/32 column groups, at most 4)
template <int BN> struct EpiCfg {
  static constexpr int COL_GROUPS = (BN >= 128) ? 4 : 2;
  static constexpr int WARPS = 4 * COL_GROUPS;
  static constexpr int THREADS = 64 + WARPS * 32;
  static constexpr int COLS_PER_WARP = BN / COL_GROUPS;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One elected lane of a converged warp.  Unlike `lane == 0`, the compiler knows exactly one thread is active
// behind this predicate, so tcgen05.mma / commit / TMA (which take warp-uniform operands) are emitted straight-line
// instead of inside a per-active-thread ELECT + BRA.U.ANY loop (~90 cycles per instruction on the issuing thread).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P_elect;\n\t"
      "elect.sync _|P_elect, 0xffffffff;\n\t"
      "@P_elect mov.s32 %0, 1;\n\t"
      "}"
      : "+r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// shared -> global tile store (bulk async group); smem must stay untouched until wait_group.read
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32: thread t gets row (lane base + t), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------ descriptors
// shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------ kernel
struct TcArgs {
  int debug;
  int M, N, K;          // problem (K = contraction)
  int splitk;           // >1: partial[split][M][N] (or transposed) fp32
  int trans_out;        // store C^T (element (m,n) at C[n*ldc + m])
  int epi;
  int tma_store;        // bf16 output without epilogue math leaves through TMA tile stores (tmC)
  int skip_pre;         // EPI_BIAS_GELU2: the pre-activation output is not wanted
  void* C; void* C2; int64_t ldc;
  const float* bias; const float* resid; int64_t ldr;
  const void* aux; int64_t ldaux;
  float* partial;
  const float* ln_gamma; const float* ln_beta; bf16* ln_out; float* ln_mean; float* ln_rstd;   // fused LayerNorm (BN == 64)
};

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 3 : (BN >= 128 ? 5 : 6);
  static_assert(STAGES * STAGE_BYTES + 1024 + (BN / 64) * 16384 + 1024 <= 232448, "smem budget");
  static constexpr int TILES_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_BYTES = 1024;       // keeps the staging area 1024-byte aligned (swizzled TMA-store tiles)
  // per-epilogue-warp staging tiles for coalesced bf16 stores
  // ... or, for the TMA-store epilogue, one 128B-swizzled [128 rows][64 cols] bf16 tile per 64 output columns
  static constexpr int WARP_STAGING = EpiCfg<BN>::WARPS * 32 * (64 + 16);
  static constexpr int BLOCK_STAGING = (BN / 64) * 16384;
  static constexpr int STAGING_BYTES = WARP_STAGING > BLOCK_STAGING ? WARP_STAGING : BLOCK_STAGING;
  static constexpr int TOTAL = TILES_BYTES + BAR_BYTES + STAGING_BYTES + 1024;  // +1024 for manual alignment
};

template <typename TC>
__device__ __forceinline__ void store_row32(TC* dst, const float (&v)[32], int nvalid);
template <>
__device__ __forceinline__ void store_row32<float>(float* dst, const float (&v)[32], int nvalid) {
  if (nvalid == 32 && ((uintptr_t)dst & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i < nvalid) dst[i] = v[i];
  }
}
template <>
__device__ __forceinline__ void store_row32<bf16>(bf16* dst, const float (&v)[32], int nvalid) {
  if (nvalid == 32 && ((uintptr_t)dst & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * i + 0], v[8 * i + 1]);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
      __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
      u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
      reinterpret_cast<uint4*>(dst)[i] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i < nvalid) dst[i] = __float2bfloat16_rn(v[i]);
  }
}

__device__ __forceinline__ uint4 attn_pack8(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}

template <typename TC>
__device__ __forceinline__ void load_row32(const TC* src, float (&v)[32], int nvalid);
template <>
__device__ __forceinline__ void load_row32<float>(const float* src, float (&v)[32], int nvalid) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = i < nvalid ? src[i] : 0.f;
}
template <>
__device__ __forceinline__ void load_row32<bf16>(const bf16* src, float (&v)[32], int nvalid) {
  if (nvalid == 32 && ((uintptr_t)src & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = reinterpret_cast<const uint4*>(src)[i];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[8 * i + 2 * j] = __uint_as_float(w[j] << 16);
        v[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i < nvalid ? __bfloat162float(src[i]) : 0.f;
  }
}

template <int BN, bool A_MN, bool B_MN, typename TC>
__global__ void __launch_bounds__(EpiCfg<BN>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const TcArgs g) {
  using SL = SmemLayout<BN>;
  constexpr int STAGES = SL::STAGES;
  constexpr int ACC_STAGES = 2;
  constexpr uint32_t TMEM_COLS = (ACC_STAGES * BN <= 32) ? 32 : (ACC_STAGES * BN <= 64) ? 64 : (ACC_STAGES * BN <= 128) ? 128
                                 : (ACC_STAGES * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B needs 1024-B alignment
  uint64_t* full_bar = (uint64_t*)(smem + SL::TILES_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
  const int kb_total = (g.K + BK - 1) / BK;
  const int kb_per_split = (kb_total + g.splitk - 1) / g.splitk;
  const int total_tiles = m_tiles * n_tiles * g.splitk;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.tma_store) tma_prefetch_desc(&tmC);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EpiCfg<BN>::WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / (m_tiles * n_tiles);
        const int rem = tile % (m_tiles * n_tiles);
        const int mt = rem / n_tiles, nt = rem % n_tiles;
        const int kb0 = split * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * SL::STAGE_BYTES;
          uint8_t* sb = sa + SL::A_BYTES;
          mbar_expect_tx(&full_bar[stage], SL::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, mt * BM);                 // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)                                           // boxes {64 m, 64 k}
              tma_load_2d(sa + i * (64 * BK * 2), &tmA, &full_bar[stage], mt * BM + 64 * i, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, nt * BN);                 // box {64 k, BN n}
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sb + i * (64 * BK * 2), &tmB, &full_bar[stage], nt * BN + 64 * i, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / (m_tiles * n_tiles);
        const int kb0 = split * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SL::STAGE_BYTES);
          const uint32_t sb = sa + SL::A_BYTES;
          // K-major : rows of 128 B, 8-row swizzle atoms 1024 B apart; k-step = 32 B inside the span
          // MN-major: k-rows of 128 B (64 MN elements), 64-wide MN blocks 8192 B apart; k-step = 16 rows
          const uint64_t adesc = A_MN ? make_smem_desc(sa, 64 * BK * 2, 1024) : make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc(sb, 64 * BK * 2, 1024) : make_smem_desc(sb, 16, 1024);
          constexpr uint32_t a_step = (A_MN ? UK * 128 : UK * 2) >> 4;
          constexpr uint32_t b_step = (B_MN ? UK * 128 : UK * 2) >> 4;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_bf16(tacc, adesc + (uint64_t)(k * a_step), bdesc + (uint64_t)(k * b_step), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);      // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);          // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int grp = ew >> 2;                   // which column group of the tile
    constexpr int COLS_PER_WARP = EpiCfg<BN>::COLS_PER_WARP;
    constexpr int CHUNKS = COLS_PER_WARP / 32;
    // per-warp staging tile: 32 rows x 64 B (one 32-column bf16 chunk) + 16 B pad -> conflict-free,
    // used to turn row-per-lane register tiles into coalesced 64-byte row segments (and back)
    constexpr int ROW_PITCH = 64 + 16;
    uint8_t* stage_buf = smem + SL::TILES_BYTES + SL::BAR_BYTES + ew * (32 * ROW_PITCH);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / (m_tiles * n_tiles);
      const int rem = tile % (m_tiles * n_tiles);
      const int mt = rem / n_tiles, nt = rem % n_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row0 = mt * BM + quad * 32;
      const int row = row0 + lane;
      const bool row_ok = row < g.M;
      const bool rows_full = row0 + 32 <= g.M;
      const int colw = nt * BN + grp * COLS_PER_WARP;            // first column of this warp
      const uint32_t tbase = tmem_base + acc * BN + grp * COLS_PER_WARP + ((uint32_t)(quad * 32) << 16);
      const bool partial_out = (g.splitk > 1 || g.trans_out);
      // coalesced path: bf16 output, 16-byte aligned rows
      const bool can_stage = std::is_same<TC, bf16>::value && !partial_out && (g.ldc % 8 == 0) &&
                             (((uintptr_t)g.C & 15) == 0) && (g.epi != EPI_BIAS_GELU2 || ((uintptr_t)g.C2 & 15) == 0) &&
                             ((g.epi != EPI_GELU_BWD && g.epi != EPI_GELU_BWD2) ||
                              ((g.ldaux % 8 == 0) && (((uintptr_t)g.aux & 15) == 0))) &&
                             (g.epi != EPI_GELU_BWD2 || ((uintptr_t)g.C2 & 15) == 0);
      // lane <-> (row, 16-byte piece) mapping of the coalesced phase: 4 lanes per 64-byte row segment
      const int rr0 = lane >> 2, c16 = lane & 3;
      auto stage_store = [&](const float (&v)[32], bf16* Cb, int col) {
        uint8_t* dst = stage_buf + lane * ROW_PITCH;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(dst + i * 16) = attn_pack8(v + 8 * i);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 8 + rr0;
          const uint4 u = *reinterpret_cast<const uint4*>(stage_buf + rr * ROW_PITCH + c16 * 16);
          if ((rows_full || row0 + rr < g.M) && g.debug != 1)
            *reinterpret_cast<uint4*>(Cb + (int64_t)(row0 + rr) * g.ldc + col + c16 * 8) = u;
        }
        __syncwarp();
      };
      auto stage_load = [&](float (&x)[32], const bf16* Xb, int64_t ldx, int col) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 8 + rr0;
          uint4 u = make_uint4(0, 0, 0, 0);
          if (rows_full || row0 + rr < g.M)
            u = *reinterpret_cast<const uint4*>(Xb + (int64_t)(row0 + rr) * ldx + col + c16 * 8);
          *reinterpret_cast<uint4*>(stage_buf + rr * ROW_PITCH + c16 * 16) = u;
        }
        __syncwarp();
        const uint8_t* src = stage_buf + lane * ROW_PITCH;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 u = *reinterpret_cast<const uint4*>(src + i * 16);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            x[8 * i + 2 * j] = __uint_as_float(w[j] << 16);
            x[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
          }
        }
        __syncwarp();
      };
      if constexpr (std::is_same<TC, bf16>::value) {
        if (g.tma_store) {
          // ---- plain bf16 output: registers -> swizzled smem tile -> one TMA store per 64-column block.
          constexpr int WARPS_PER_BLK = 4 * (64 / COLS_PER_WARP);
          const int blk = (grp * COLS_PER_WARP) / 64;                       // 64-column block of the tile
          uint8_t* sblk = smem + SL::TILES_BYTES + SL::BAR_BYTES + blk * 16384;
          const bool leader = (quad == 0) && ((grp * COLS_PER_WARP) % 64 == 0);
          if (leader && elect_one_sync()) tma_store_wait_read();            // previous tile's store has drained the block
          named_bar_sync(1 + blk, WARPS_PER_BLK * 32);
          const int r = quad * 32 + lane;
#pragma unroll
          for (int ch = 0; ch < CHUNKS; ++ch) {
            float v[32];
            tmem_ld32(tbase + ch * 32, v);
            const int c16 = ((grp * COLS_PER_WARP + ch * 32) % 64) / 8;     // first 16-byte chunk inside the 128-byte row
            const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(sblk + rowoff + (((c16 + i) ^ (r & 7)) << 4)) = attn_pack8(v + 8 * i);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);                     // accumulator drained: next tile's MMAs may start
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          named_bar_sync(1 + blk, WARPS_PER_BLK * 32);
          if (leader && elect_one_sync()) {
            if (g.debug != 1) tma_store_2d(&tmC, sblk, nt * BN + blk * 64, mt * BM);
            tma_store_commit();
          }
          if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
          continue;
        }
      }
#pragma unroll 1
      for (int ch = 0; ch < CHUNKS; ++ch) {
        if (g.debug == 2) break;
        float v[32];
        tmem_ld32(tbase + ch * 32, v);
        const int col = colw + ch * 32;
        const int nvalid = min(32, g.N - col);
        if (nvalid <= 0) continue;                                  // warp-uniform
        if (partial_out) {
          if (!row_ok) continue;
          float* P = g.splitk > 1 ? g.partial + (int64_t)split * g.M * g.N : (float*)g.C;
          if (g.trans_out) {
            const int64_t ld = g.splitk > 1 ? g.M : g.ldc;
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < nvalid) P[(int64_t)(col + i) * ld + row] = v[i];
          } else {
            store_row32<float>(P + (int64_t)row * (g.splitk > 1 ? g.N : g.ldc) + col, v, nvalid);
          }
          continue;
        }
        const bool full = (nvalid == 32);                           // warp-uniform
        const bool staged = can_stage && full;
        // ---- bias (all epilogues that have one)
        if (g.bias && (g.epi == EPI_BIAS || g.epi == EPI_BIAS_RELU || g.epi == EPI_BIAS_GELU2 || g.epi == EPI_BIAS_RESID)) {
          if (full) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col) + i);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < nvalid) v[i] += __ldg(g.bias + col + i);
          }
        }
        TC* Cb = (TC*)g.C;
        switch (g.epi) {
          case EPI_BIAS_RELU:
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            break;
          case EPI_BIAS_GELU2:
            // first output: the pre-activation (saved for backward); then GELU in place
            if (g.skip_pre) {
            } else if (staged) {
              if constexpr (std::is_same<TC, bf16>::value) stage_store(v, (bf16*)g.C, col);
            } else if (row_ok) {
              store_row32<TC>((TC*)g.C + (int64_t)row * g.ldc + col, v, nvalid);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu3(v[i]);
            Cb = (TC*)g.C2;
            break;
          case EPI_BIAS_RESID:
            if (row_ok) {
              const float* R = g.resid + (int64_t)row * g.ldr + col;
              if (full && ((((uintptr_t)R) & 15) == 0)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 r4 = reinterpret_cast<const float4*>(R)[i];
                  v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (i < nvalid) v[i] += R[i];
              }
            }
            break;
          case EPI_GELU_BWD2:
          case EPI_GELU_BWD: {
            float xin[32];
            if (staged) {
              if constexpr (std::is_same<TC, bf16>::value) stage_load(xin, (const bf16*)g.aux, g.ldaux, col);
            } else {
              if (row_ok) load_row32<TC>((const TC*)g.aux + (int64_t)row * g.ldaux + col, xin, nvalid);
              else {
#pragma unroll
                for (int i = 0; i < 32; ++i) xin[i] = 0.f;
              }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) { float y, dy; gelu3_grad(xin[i], y, dy); v[i] *= dy; xin[i] = y; }
            if (g.epi == EPI_GELU_BWD2) {       // second output: the recomputed activation gelu(x)
              if (staged) {
                if constexpr (std::is_same<TC, bf16>::value) stage_store(xin, (bf16*)g.C2, col);
              } else if (row_ok) {
                store_row32<TC>((TC*)g.C2 + (int64_t)row * g.ldc + col, xin, nvalid);
              }
            }
          } break;
          case EPI_RELU_BWD:
            if (row_ok) {
              const float* X = (const float*)g.aux + (int64_t)row * g.ldaux + col;
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < nvalid) v[i] = X[i] > 0.f ? v[i] : 0.f;
            }
            break;
          default: break;
        }
        if (staged) {
          if constexpr (std::is_same<TC, bf16>::value) stage_store(v, (bf16*)Cb, col);
        } else if (row_ok) {
          store_row32<TC>(Cb + (int64_t)row * g.ldc + col, v, nvalid);
        }
        if constexpr (std::is_same<TC, float>::value && BN == 64) {
          if (g.ln_gamma) {
            // LayerNorm over the 64 output columns of a row (PreNorm of the following block, vn/GoalFormer.py:31-37):
            // this warp holds 32 of them, the other column group's warp of the same lane quadrant the rest;
            // two-pass statistics exchanged through shared memory
            float* part = reinterpret_cast<float*>(smem + SL::TILES_BYTES + SL::BAR_BYTES);   // [128][2]
            const int r = quad * 32 + lane;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) s += v[i];
            part[r * 2 + grp] = s;
            named_bar_sync(1 + quad, 64);
            const float mu = (part[r * 2] + part[r * 2 + 1]) * (1.0f / 64.0f);
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float c = v[i] - mu; q = fmaf(c, c, q); }
            named_bar_sync(1 + quad, 64);
            part[r * 2 + grp] = q;
            named_bar_sync(1 + quad, 64);
            const float rs = 1.0f / sqrtf((part[r * 2] + part[r * 2 + 1]) * (1.0f / 64.0f) + 1e-5f);
            named_bar_sync(1 + quad, 64);                 // table free for the next tile
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(g.ln_gamma + col) + i);
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.ln_beta + col) + i);
                v[4 * i] = (v[4 * i] - mu) * rs * g4.x + b4.x; v[4 * i + 1] = (v[4 * i + 1] - mu) * rs * g4.y + b4.y;
                v[4 * i + 2] = (v[4 * i + 2] - mu) * rs * g4.z + b4.z; v[4 * i + 3] = (v[4 * i + 3] - mu) * rs * g4.w + b4.w;
              }
              store_row32<bf16>(g.ln_out + (int64_t)row * 64 + col, v, 32);
              if (grp == 0 && g.ln_mean) { g.ln_mean[row] = mu; g.ln_rstd[row] = rs; }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (g.tma_store) tma_store_wait_read();     // (no-op for threads that issued nothing)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    DG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    DG_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 row-major matrix [outer, inner] with row pitch ld (elements); box {box_inner, box_outer}
static CUtensorMap make_map(const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(DGVIT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%ld outer=%ld ld=%ld", (int)r,
                              (long)inner, (long)outer, (long)ld);
  return m;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    DG_CUDA(cudaGetDevice(&dev));
    DG_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

template <int BN, bool A_MN, bool B_MN, typename TC>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const TcArgs& a, cudaStream_t st) {
  using SL = SmemLayout<BN>;
  static DevOnce attr;
  if (attr.first()) {
    DG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL));
  }
  const int tiles = (int)(cdiv(a.M, BM) * cdiv(a.N, BN) * a.splitk);
  const int grid = std::min(tiles, sm_count());
  if (skip_mask() & SKIP_GEMM_TC) return;
  launch_k(gemm_tc_kernel<BN, A_MN, B_MN, TC>, grid, EpiCfg<BN>::THREADS, SL::TOTAL, st, ta, tb, tc_, a);
  DG_LAUNCH_CHECK();
}

template <typename TC>
static void dispatch(bool a_mn, bool b_mn, int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_,
                     const TcArgs& a, cudaStream_t st) {
#define DG_TC_CASE(BN_)                                                        \
  if (bn == BN_) {                                                             \
    if (!a_mn && !b_mn) return launch<BN_, false, false, TC>(ta, tb, tc_, a, st);   \
    if (!a_mn && b_mn) return launch<BN_, false, true, TC>(ta, tb, tc_, a, st);     \
    if (a_mn && b_mn) return launch<BN_, true, true, TC>(ta, tb, tc_, a, st);       \
  }
  DG_TC_CASE(64)
  DG_TC_CASE(128)
  DG_TC_CASE(256)
#undef DG_TC_CASE
  fail(DGVIT_ERR_ARG, "gemm_tc: unsupported operand layout (a_mn=%d b_mn=%d bn=%d)", (int)a_mn, (int)b_mn, bn);
}

static bool g_tc_enabled = true;
static int g_debug = 0;   // measurement knob: 1 = skip epilogue global stores, 2 = skip the whole epilogue body

}  // namespace tc

// Returns false when the problem is not eligible (caller falls back to the CUDA-core kernel,
// which only happens for the tiny head GEMMs and fp32 operands).
template <typename TA, typename TB, typename TC>
static bool gemm_tc_try(const GemmArgs& g0, cudaStream_t st) {
  if constexpr (!(std::is_same<TA, bf16>::value && std::is_same<TB, bf16>::value)) {
    return false;
  } else {
    using namespace tc;
    if (!g_tc_enabled) return false;
    GemmArgs g = g0;
    int trans_out = 0;
    // weight-gradient shapes with a narrow output (e.g. dW2 [64, 2048]): compute C^T so that the
    // wide dimension rides on UMMA M = 128
    if (g.M < 128 && g.N >= 128 && g.epi == EPI_NONE && std::is_same<TC, float>::value) {
      std::swap(g.M, g.N);
      const void* a = g.A; int64_t a_sm = g.a_sm, a_sk = g.a_sk;
      g.A = g.B; g.a_sm = g.b_sn; g.a_sk = g.b_sk;
      g.B = a; g.b_sk = a_sk; g.b_sn = a_sm;
      trans_out = 1;
    }
    // tiny head-sized contractions stay on CUDA cores; a short M with a real K / N (batch-1 act, the pruned last
    // block) is one mostly-empty UMMA tile, still several times faster than the serial CUDA-core K loop
    if (g.K < 64 || g.N < 64 || g.M < 1) return false;
    const bool a_mn = (g.a_sm == 1 && g.a_sk != 1), b_mn = (g.b_sn == 1 && g.b_sk != 1);
    const bool a_k = (g.a_sk == 1), b_k = (g.b_sk == 1);
    if (!(a_mn || a_k) || !(b_mn || b_k)) return false;
    if (a_mn && !b_mn) return false;                       // combination not used by the hot path
    const int64_t lda = a_mn ? g.a_sk : g.a_sm, ldb = b_mn ? g.b_sk : g.b_sn;
    if (lda % 8 || ldb % 8 || ((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15)) return false;   // TMA: 16-B strides
    if (g.splitk > 1 && g.epi != EPI_NONE) return false;
    if (g.ln_gamma && !(std::is_same<TC, float>::value && g.N == 64 && g.epi == EPI_BIAS_RESID && g.splitk == 1 && !trans_out &&
                        g.ln_beta && g.ln_out && (((uintptr_t)g.ln_out) & 15) == 0))
      return false;
    if (trans_out && !std::is_same<TC, float>::value) return false;
    int bn = g.N >= 256 ? 256 : (g.N >= 128 ? 128 : 64);
    if (g.N % 8) return false;
    // tensor maps: K-major operand = [rows, K] (inner K); MN-major operand = [K, MN] (inner MN)
    CUtensorMap ta = a_mn ? make_map(g.A, g.M, g.K, lda, 64, BK) : make_map(g.A, g.K, g.M, lda, BK, BM);
    CUtensorMap tb = b_mn ? make_map(g.B, g.N, g.K, ldb, 64, BK) : make_map(g.B, g.K, g.N, ldb, BK, bn);
    TcArgs a;
    a.debug = g_debug;
    a.M = g.M; a.N = g.N; a.K = g.K;
    a.splitk = g.splitk; a.trans_out = trans_out; a.epi = g.epi; a.skip_pre = g.skip_pre;
    a.C = g.C; a.C2 = g.C2; a.ldc = g.ldc; a.bias = g.bias; a.resid = g.resid; a.ldr = g.ldr;
    a.aux = g.aux; a.ldaux = g.ldaux; a.partial = g.partial;
    a.ln_gamma = g.ln_gamma; a.ln_beta = g.ln_beta; a.ln_out = (bf16*)g.ln_out; a.ln_mean = g.ln_mean; a.ln_rstd = g.ln_rstd;
    // split-K must not leave an empty split (its accumulator would be undefined)
    const int kb_total = (int)cdiv(g.K, BK);
    if (a.splitk > kb_total) a.splitk = kb_total;
    {
      const int per = (int)cdiv(kb_total, a.splitk);
      a.splitk = (int)cdiv(kb_total, per);
    }
    if (g0.splitk > 1 && a.splitk == 1 && !trans_out) { /* direct store below */ }
    // plain bf16 outputs (QKV, dO) leave through TMA tile stores
    a.tma_store = (std::is_same<TC, bf16>::value && g.epi == EPI_NONE && a.splitk == 1 && !trans_out && g.N % 64 == 0 &&
                   g.ldc % 8 == 0 && (((uintptr_t)g.C) & 15) == 0) ? 1 : 0;
    CUtensorMap tcm = a.tma_store ? make_map(g.C, g.N, g.M, g.ldc, 64, BM) : ta;
    dispatch<TC>(a_mn, b_mn, bn, ta, tb, tcm, a, st);
    if (a.splitk > 1) {
      const int64_t n = (int64_t)g0.M * g0.N;
      DG_REQUIRE(g0.ldc == g0.N, "gemm_tc: split-K output must be dense");
      if (g0.defer && n % 4 == 0) {
        g0.defer->add(g.partial, (float*)g0.C, a.splitk, n, n);
      } else {
        launch_k(reduce_partials_kernel, reduce_grid(n), 256, 0, st, g.partial, (float*)g0.C, a.splitk, n);
        DG_LAUNCH_CHECK();
      }
    }
    return true;
  }
}

}  // namespace dgvit
