// attn_tc.cuh — fused multi-head attention for sm_100a (vn/GoalFormer.py:71-81), forward and
// backward, one (sample, head) per work item: TMA-staged Q/K/V tiles (128B swizzle), tcgen05.mma
// into TMEM, one thread per query row for the fp32 softmax (the 32x32b TMEM layout gives every
// thread a whole score row, so no cross-lane reduction is needed), P / dS handed back to the
// tensor core through shared memory.
//
//   forward : S = Q K^T ; P = softmax(S * dh^-0.5) ; O = P V
//   backward: S, dP = dO V^T ; dS = P o (dP - rowsum(dO o O)) * scale ;
//             dQ = dS K ; dK = dS^T Q ; dV = P^T dO
//
// The whole sequence (N <= 128 tokens, 65 in the shipped model) of one head fits a single
// 128-row UMMA tile, so there is no online-softmax loop.  One smem image serves two operand
// roles: a [rows][64] 128B-swizzled tile is K-major for (rows x 64) and MN-major for its
// transpose, which is how K feeds both S (K-major) and dQ (MN-major), dS feeds dQ and dK, etc.
#pragma once
#include "gemm_tc.cuh"

namespace dgvit {
namespace attn {

using namespace tc;

constexpr int DH = 64;          // dim_head (reference default, vn/GoalFormer.py:124)
constexpr int TILE = 16384;     // 128 rows x 128 B
constexpr int FWD_STAGES = 2;
constexpr int FWD_THREADS = 64 + 256;   // TMA warp, MMA warp, two softmax warpgroups (one per TMEM/smem stage)

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of 16-byte chunk `chunk` (8 bf16) of row `r` inside a [128][64]-bf16 128B-swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int r, int chunk) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}

struct AttnArgs {
  int B, N, H, KPAD;     // KPAD = N rounded up to 16 (keys / contraction rows actually multiplied)
  float scale;
  bf16* O;               // fwd: output [T, inner]
  const bf16* Oin;       // bwd: forward output
  const bf16* dO;        // bwd
  bf16* dQKV;            // bwd: [T, 3*inner]
};

// =====================================================================================
// forward
// =====================================================================================
// smem: a ring of LS load stages (Q | K | V, KPAD rows of 128 B each) decoupled from the two compute stages
// (P block0 | P block1 + one TMEM stage each), so the TMA producer runs LS items ahead of the softmax.
// The S MMA reads 128 query rows from a KPAD-row Q tile: rows >= KPAD alias the following tiles and only produce
// score rows that are never stored.
template <int NK> struct FwdSmem {
  static constexpr int LT = NK * 2048;                       // one Q / K / V tile
  static constexpr int LS = (26 / NK) < 4 ? (26 / NK) : 4;   // load stages that fit beside the P buffers
  static constexpr int LOAD_STAGE = 3 * LT;
  static constexpr int OFF_P = LS * LOAD_STAGE;
  static constexpr int OFF_BAR = OFF_P + FWD_STAGES * 2 * TILE;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
  static_assert(LS >= 2 && TOTAL <= 232448, "smem budget");
};

template <int NK>   // NK = KPAD / 16: key blocks held in registers by the softmax
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmKV, const AttnArgs a) {
  using SM = FwdSmem<NK>;
  constexpr int LS = SM::LS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + SM::OFF_BAR);
  uint64_t* qkv_full = bars;                    // [LS] TMA landed
  uint64_t* qkv_empty = qkv_full + LS;          // [LS] O-MMA retired: load stage reusable
  uint64_t* s_full = qkv_empty + LS;            // [2] S in TMEM
  uint64_t* p_ready = s_full + FWD_STAGES;      // [2] P written to smem (4 warps)
  uint64_t* o_full = p_ready + FWD_STAGES;      // [2] O in TMEM
  uint64_t* t_free = o_full + FWD_STAGES;       // [2] TMEM stage drained (4 warps)
  uint32_t* tmem_slot = (uint32_t*)(t_free + FWD_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inner = a.H * DH;
  const int items = a.B * a.H;
  constexpr uint32_t TCOLS_STAGE = 256;   // S: cols [0,128), O: cols [128,192)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < LS; ++i) { mbar_init(&qkv_full[i], 1); mbar_init(&qkv_empty[i], 1); }
    for (int i = 0; i < FWD_STAGES; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_full[i], 1); mbar_init(&t_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if (elect_one_sync()) {
      int ls = 0; uint32_t ph = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int b = it / a.H, h = it % a.H;
        mbar_wait(&qkv_empty[ls], ph ^ 1);
        uint8_t* base = smem + ls * SM::LOAD_STAGE;
        mbar_expect_tx(&qkv_full[ls], 3 * SM::LT);
        tma_load_2d(base, &tmKV, &qkv_full[ls], h * DH, b * a.N);
        tma_load_2d(base + SM::LT, &tmKV, &qkv_full[ls], inner + h * DH, b * a.N);
        tma_load_2d(base + 2 * SM::LT, &tmKV, &qkv_full[ls], 2 * inner + h * DH, b * a.N);
        if (++ls == LS) { ls = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc_s = make_idesc(128, NK * 16, false, false);
      constexpr uint32_t idesc_o = make_idesc(128, DH, false, true);
      int st = 0; uint32_t ph = 0;          // compute stage of the S issue
      int ls = 0; uint32_t lph = 0;         // load stage of the S issue
      int st2 = 0; uint32_t ph2 = 0;        // O side (one item behind)
      int ls2 = 0;
      int n_mine = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) ++n_mine;
      for (int i = 0; i <= n_mine; ++i) {
        if (i < n_mine) {
          mbar_wait(&t_free[st], ph ^ 1);
          mbar_wait(&qkv_full[ls], lph);
          tc_fence_after();
          const uint32_t sq = smem_u32(smem + ls * SM::LOAD_STAGE), sk = sq + SM::LT;
          const uint64_t qd = make_smem_desc(sq, 16, 1024), kd = make_smem_desc(sk, 16, 1024);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16(tmem_base + st * TCOLS_STAGE, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k > 0);
          umma_commit(&s_full[st]);
          if (++st == FWD_STAGES) { st = 0; ph ^= 1; }
          if (++ls == LS) { ls = 0; lph ^= 1; }
        }
        if (i >= 1) {
          mbar_wait(&p_ready[st2], ph2);
          tc_fence_after();
          const uint32_t sv = smem_u32(smem + ls2 * SM::LOAD_STAGE) + 2 * SM::LT;
          const uint32_t sp = smem_u32(smem + SM::OFF_P + st2 * 2 * TILE);
#pragma unroll
          for (int k = 0; k < NK; ++k) {
            // A = P (K-major, 64-key blocks one tile apart); B = V as MN-major (16 key rows per step)
            const uint64_t pd = make_smem_desc(sp + (k >> 2) * TILE + (k & 3) * 32, 16, 1024);
            const uint64_t vd = make_smem_desc(sv + k * 2048, 8192, 1024);
            umma_bf16(tmem_base + st2 * TCOLS_STAGE + 128, pd, vd, idesc_o, k > 0);
          }
          umma_commit(&o_full[st2]);
          umma_commit(&qkv_empty[ls2]);
          if (++st2 == FWD_STAGES) { st2 = 0; ph2 ^= 1; }
          if (++ls2 == LS) ls2 = 0;
        }
      }
    }
  } else {
    // ---- softmax + output: thread = query row.  Warpgroup wg owns pipeline stage wg, i.e. every
    //      second work item of this CTA, so two softmaxes are in flight while the tensor core runs.
    const int quad = warp & 3;
    const int wg = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const float sl2 = a.scale * 1.44269504088896f;
    const int st = wg; uint32_t ph = 0;
    for (int it = blockIdx.x + wg * gridDim.x; it < items; it += FWD_STAGES * gridDim.x) {
      const int b = it / a.H, h = it % a.H;
      uint8_t* P = smem + SM::OFF_P + st * 2 * TILE;
      mbar_wait(&s_full[st], ph);
      tc_fence_after();
      const uint32_t ts = tmem_base + st * TCOLS_STAGE + lane_off;
      // the whole score row in registers: one TMEM round trip, one exp per element
      uint32_t sr[NK][16];
#pragma unroll
      for (int k = 0; k < NK; ++k) tmem_ld16_nowait(ts + 16 * k, sr[k]);
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int i = 0; i < 16; ++i) if (16 * k + i < a.N) mx = fmaxf(mx, __uint_as_float(sr[k][i]));
      const float mb = mx * sl2;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c0 = 16 * k + 2 * i;
          const float e0 = (c0 < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[k][2 * i]), sl2, -mb)) : 0.f;
          const float e1 = (c0 + 1 < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[k][2 * i + 1]), sl2, -mb)) : 0.f;
          w[i] = pack2(e0, e1);
          // the tensor core sees the bf16-rounded value: sum what it sees
          sum += __uint_as_float(w[i] << 16) + __uint_as_float(w[i] & 0xffff0000u);
        }
        uint8_t* blk = P + (k >> 2) * TILE;
        const int ch = (k & 3) * 2;
        *reinterpret_cast<uint4*>(blk + sw128_off(r, ch)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(blk + sw128_off(r, ch + 1)) = make_uint4(w[4], w[5], w[6], w[7]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[st]);
      mbar_wait(&o_full[st], ph);
      tc_fence_after();
      const float inv = 1.0f / sum;
      {
        // tcgen05.ld is warp-aligned: every lane loads, only rows inside the sequence store
        bf16* dst = a.O + ((int64_t)b * a.N + r) * inner + h * DH;
        uint32_t orr[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16_nowait(ts + 128 + 16 * c, orr[c]);
        tmem_ld_wait();
        if (r < a.N) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = pack2(__uint_as_float(orr[c][2 * i]) * inv, __uint_as_float(orr[c][2 * i + 1]) * inv);
            reinterpret_cast<uint4*>(dst + 16 * c)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4*>(dst + 16 * c)[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_free[st]);
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================
// backward
// =====================================================================================
// smem: 2 load stages x (Q | K | V | dO) + P block0,1 + dS block0,1
constexpr int BWD_LOAD_STAGES = 2;
constexpr int BWD_THREADS = 64 + 128;

template <int NK>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmdO, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int LOAD_BYTES = 4 * TILE;
  uint8_t* Pt = smem + BWD_LOAD_STAGES * LOAD_BYTES;   // 2 tiles
  uint8_t* dSt = Pt + 2 * TILE;                        // 2 tiles
  uint64_t* bars = (uint64_t*)(dSt + 2 * TILE);
  uint64_t* ld_full = bars;                       // [2]
  uint64_t* ld_empty = ld_full + BWD_LOAD_STAGES; // [2]
  uint64_t* sdp_full = ld_empty + BWD_LOAD_STAGES;   // S and dP in TMEM
  uint64_t* pds_ready = sdp_full + 1;                // P, dS in smem (4 warps)
  uint64_t* out_full = pds_ready + 1;                // dQ, dK, dV in TMEM
  uint64_t* t_free = out_full + 1;                   // TMEM drained (4 warps)
  uint32_t* tmem_slot = (uint32_t*)(t_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inner = a.H * DH;
  const int items = a.B * a.H;
  // TMEM columns: S [0,128) dP [128,256) dQ [256,320) dK [320,384) dV [384,448)
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmdO);
    for (int i = 0; i < BWD_LOAD_STAGES; ++i) { mbar_init(&ld_full[i], 1); mbar_init(&ld_empty[i], 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_ready, 4); mbar_init(out_full, 1); mbar_init(t_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if (elect_one_sync()) {
      int st = 0; uint32_t ph = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int b = it / a.H, h = it % a.H;
        mbar_wait(&ld_empty[st], ph ^ 1);
        uint8_t* base = smem + st * LOAD_BYTES;
        mbar_expect_tx(&ld_full[st], 2 * TILE + 2 * a.KPAD * 128);
        tma_load_2d(base, &tmQ, &ld_full[st], h * DH, b * a.N);
        tma_load_2d(base + TILE, &tmKV, &ld_full[st], inner + h * DH, b * a.N);
        tma_load_2d(base + 2 * TILE, &tmKV, &ld_full[st], 2 * inner + h * DH, b * a.N);
        tma_load_2d(base + 3 * TILE, &tmdO, &ld_full[st], h * DH, b * a.N);
        if (++st == BWD_LOAD_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc_s = make_idesc(128, a.KPAD, false, false);    // S, dP
      const uint32_t idesc_q = make_idesc(128, DH, false, true);         // dQ = dS K      (A K-major, B MN-major)
      const uint32_t idesc_t = make_idesc(128, DH, true, true);          // dK, dV         (A MN-major, B MN-major)
      const int ksteps = a.KPAD / 16;
      int st = 0; uint32_t ph = 0; uint32_t iph = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        mbar_wait(t_free, iph ^ 1);
        mbar_wait(&ld_full[st], ph);
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + st * LOAD_BYTES), sk = sq + TILE, sv = sk + TILE, sdo = sv + TILE;
        const uint32_t sp = smem_u32(Pt), sds = smem_u32(dSt);
        {
          const uint64_t qd = make_smem_desc(sq, 16, 1024), kd = make_smem_desc(sk, 16, 1024);
          const uint64_t dod = make_smem_desc(sdo, 16, 1024), vd = make_smem_desc(sv, 16, 1024);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + 128, dod + (uint64_t)(k * 2), vd + (uint64_t)(k * 2), idesc_s, k > 0);
        }
        umma_commit(sdp_full);
        mbar_wait(pds_ready, iph);
        tc_fence_after();
        for (int k = 0; k < ksteps; ++k) {
          // dQ[128 x 64] += dS[:, 16k:16k+16] K[16k:16k+16, :]
          const uint64_t ad = make_smem_desc(sds + (k >> 2) * TILE + (k & 3) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc(sk + k * 2048, 8192, 1024);
          umma_bf16(tmem_base + 256, ad, bd, idesc_q, k > 0);
        }
        for (int k = 0; k < ksteps; ++k) {
          // dK[keys x 64] += dS^T[:, 16k rows of queries] Q[16k.., :]   (A = dS tile read MN-major)
          const uint64_t ad = make_smem_desc(sds + k * 2048, TILE, 1024);
          const uint64_t bd = make_smem_desc(sq + k * 2048, 8192, 1024);
          umma_bf16(tmem_base + 320, ad, bd, idesc_t, k > 0);
        }
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ad = make_smem_desc(sp + k * 2048, TILE, 1024);
          const uint64_t bd = make_smem_desc(sdo + k * 2048, 8192, 1024);
          umma_bf16(tmem_base + 384, ad, bd, idesc_t, k > 0);
        }
        umma_commit(out_full);
        umma_commit(&ld_empty[st]);
        iph ^= 1;
        if (++st == BWD_LOAD_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const float sl2 = a.scale * 1.44269504088896f;
    uint32_t iph = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int b = it / a.H, h = it % a.H;
      const bool valid = r < a.N;
      // delta = rowsum(dO o O) straight from global (bf16, 128 B per row each)
      float delta = 0.f;
      if (valid) {
        const uint4* po = reinterpret_cast<const uint4*>(a.Oin + ((int64_t)b * a.N + r) * inner + h * DH);
        const uint4* pd = reinterpret_cast<const uint4*>(a.dO + ((int64_t)b * a.N + r) * inner + h * DH);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 x = po[i], y = pd[i];
          const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            delta = fmaf(__uint_as_float(xs[j] << 16), __uint_as_float(ys[j] << 16), delta);
            delta = fmaf(__uint_as_float(xs[j] & 0xffff0000u), __uint_as_float(ys[j] & 0xffff0000u), delta);
          }
        }
      }
      mbar_wait(sdp_full, iph);
      tc_fence_after();
      const uint32_t ts = tmem_base + lane_off;
      // S row in registers (one exp per element); dP streamed 16 columns at a time
      uint32_t sr[NK][16];
#pragma unroll
      for (int k = 0; k < NK; ++k) tmem_ld16_nowait(ts + 16 * k, sr[k]);
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int i = 0; i < 16; ++i) if (16 * k + i < a.N) mx = fmaxf(mx, __uint_as_float(sr[k][i]));
      const float mb = mx * sl2;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e = (valid && 16 * k + i < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[k][i]), sl2, -mb)) : 0.f;
          sr[k][i] = __float_as_uint(e);
          sum += e;
        }
      const float inv = valid ? 1.0f / sum : 0.f;
      const float nds = -delta;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        uint32_t dpr[16];
        tmem_ld16_nowait(ts + 128 + 16 * k, dpr);
        tmem_ld_wait();
        uint32_t wp[8], wd[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float p0 = __uint_as_float(sr[k][2 * i]) * inv, p1 = __uint_as_float(sr[k][2 * i + 1]) * inv;
          const float d0 = p0 * (__uint_as_float(dpr[2 * i]) + nds) * a.scale, d1 = p1 * (__uint_as_float(dpr[2 * i + 1]) + nds) * a.scale;
          wp[i] = pack2(p0, p1);
          wd[i] = pack2(d0, d1);
        }
        const int ch = (k & 3) * 2;
        uint8_t* pb = Pt + (k >> 2) * TILE;
        uint8_t* db = dSt + (k >> 2) * TILE;
        *reinterpret_cast<uint4*>(pb + sw128_off(r, ch)) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
        *reinterpret_cast<uint4*>(pb + sw128_off(r, ch + 1)) = make_uint4(wp[4], wp[5], wp[6], wp[7]);
        *reinterpret_cast<uint4*>(db + sw128_off(r, ch)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        *reinterpret_cast<uint4*>(db + sw128_off(r, ch + 1)) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);
      mbar_wait(out_full, iph);
      tc_fence_after();
      bf16* dst = a.dQKV + ((int64_t)b * a.N + r) * (3 * inner) + h * DH;
#pragma unroll 1
      for (int w = 0; w < 3; ++w) {          // dQ, dK, dV rows (query r / key r)
        uint32_t orr[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16_nowait(ts + 256 + w * 64 + 16 * c, orr[c]);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = pack2(__uint_as_float(orr[c][2 * i]), __uint_as_float(orr[c][2 * i + 1]));
            reinterpret_cast<uint4*>(dst + w * inner + 16 * c)[0] = make_uint4(u[0], u[1], u[2], u[3]);
            reinterpret_cast<uint4*>(dst + w * inner + 16 * c)[1] = make_uint4(u[4], u[5], u[6], u[7]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_free);
      iph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================
// backward, pipelined (KPAD <= 80): two softmax warpgroups alternate work items.  S / dP accumulators and the P / dS
// shared-memory tiles are double-buffered (one set per warpgroup), the dQ / dK / dV accumulators are shared, the TMA
// ring runs three items ahead, and every tile holds KPAD rows only (a 128-row A operand over-reads into the next tile:
// those accumulator rows are never stored).  While one warpgroup drains and stores an item's gradients the other one is
// in its softmax, and the tensor core works on whichever is ready.
//   TMEM: stage g: S [160g, 160g+80)  dP [160g+80, 160g+160) ; dQ [320,384) dK [384,448) dV [448,512)
// =====================================================================================
constexpr int BWD2_THREADS = 64 + 256;
template <int NK> struct Bwd2Smem {
  static constexpr int BLK = NK * 2048;                 // one [KPAD rows][64 cols] bf16 tile
  static constexpr int LS = 3;                          // load stages: Q | K | V | dO
  static constexpr int OFF_PD = 0;                      // stage g: P block0, block1, dS block0, block1
  static constexpr int OFF_LD = 2 * 4 * BLK;
  static constexpr int OFF_BAR = OFF_LD + LS * 4 * BLK + 16384;   // slack behind the last tile: a 128-row A operand over-reads up to 16 KB - BLK
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
  static_assert(NK <= 5 && TOTAL <= 232448, "smem budget");
};

template <int NK>
__global__ void __launch_bounds__(BWD2_THREADS, 1)
attn_bwd2_tc_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmdO, const AttnArgs a) {
  using SM = Bwd2Smem<NK>;
  constexpr int BLK = SM::BLK, LS = SM::LS, KPAD = NK * 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + SM::OFF_BAR);
  uint64_t* ld_full = bars;                 // [LS]
  uint64_t* ld_empty = ld_full + LS;        // [LS]
  uint64_t* sdp_full = ld_empty + LS;       // [2] S and dP of the stage in TMEM
  uint64_t* sdp_free = sdp_full + 2;        // [2] (4 warps) both pulled into registers
  uint64_t* pds_ready = sdp_free + 2;       // [2] (4 warps) P, dS tiles written
  uint64_t* out_full = pds_ready + 2;       // [2] dQ, dK, dV of the stage's item in TMEM
  uint64_t* out_free = out_full + 2;        // [2] (4 warps) drained
  uint32_t* tmem_slot = (uint32_t*)(out_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inner = a.H * DH;
  const int items = a.B * a.H;
  int n_mine = 0;
  for (int it = blockIdx.x; it < items; it += gridDim.x) ++n_mine;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmdO);
    for (int i = 0; i < LS; ++i) { mbar_init(&ld_full[i], 1); mbar_init(&ld_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1); mbar_init(&sdp_free[i], 4); mbar_init(&pds_ready[i], 4);
      mbar_init(&out_full[i], 1); mbar_init(&out_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if (elect_one_sync()) {
      for (int i = 0; i < n_mine; ++i) {
        const int it = blockIdx.x + i * gridDim.x;
        const int b = it / a.H, h = it % a.H;
        const int ls = i % LS; const uint32_t ph = (i / LS) & 1;
        mbar_wait(&ld_empty[ls], ph ^ 1);
        uint8_t* base = smem + SM::OFF_LD + ls * 4 * BLK;
        mbar_expect_tx(&ld_full[ls], 4 * BLK);
        tma_load_2d(base, &tmKV, &ld_full[ls], h * DH, b * a.N);
        tma_load_2d(base + BLK, &tmKV, &ld_full[ls], inner + h * DH, b * a.N);
        tma_load_2d(base + 2 * BLK, &tmKV, &ld_full[ls], 2 * inner + h * DH, b * a.N);
        tma_load_2d(base + 3 * BLK, &tmdO, &ld_full[ls], h * DH, b * a.N);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc_s = make_idesc(128, KPAD, false, false);   // S, dP
      constexpr uint32_t idesc_q = make_idesc(128, DH, false, true);      // dQ = dS K      (A K-major, B MN-major)
      constexpr uint32_t idesc_t = make_idesc(128, DH, true, true);       // dK, dV         (A MN-major, B MN-major)
      auto tiles = [&](int i) { return smem_u32(smem + SM::OFF_LD + (i % LS) * 4 * BLK); };
      auto issue_sdp = [&](int i) {
        const int g = i & 1, j = i >> 1;
        if (j > 0) mbar_wait(&sdp_free[g], (uint32_t)((j - 1) & 1));
        mbar_wait(&ld_full[i % LS], (uint32_t)((i / LS) & 1));
        tc_fence_after();
        const uint32_t sq = tiles(i), sk = sq + BLK, sv = sk + BLK, sdo = sv + BLK;
        const uint64_t qd = make_smem_desc(sq, 16, 1024), kd = make_smem_desc(sk, 16, 1024);
        const uint64_t dod = make_smem_desc(sdo, 16, 1024), vd = make_smem_desc(sv, 16, 1024);
        const uint32_t ts = tmem_base + g * 160;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(ts, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(ts + 80, dod + (uint64_t)(k * 2), vd + (uint64_t)(k * 2), idesc_s, k > 0);
        umma_commit(&sdp_full[g]);
      };
      if (n_mine > 0) issue_sdp(0);
      for (int i = 0; i < n_mine; ++i) {
        if (i + 1 < n_mine) issue_sdp(i + 1);
        const int g = i & 1, j = i >> 1;
        mbar_wait(&pds_ready[g], (uint32_t)(j & 1));
        if (i > 0) mbar_wait(&out_free[g ^ 1], (uint32_t)(((i - 1) >> 1) & 1));   // the shared accumulators are drained
        tc_fence_after();
        const uint32_t sq = tiles(i), sk = sq + BLK, sdo = sk + 2 * BLK;
        const uint32_t sp = smem_u32(smem + SM::OFF_PD + g * 4 * BLK), sds = sp + 2 * BLK;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          // dQ[128 x 64] += dS[:, 16k:16k+16] K[16k:16k+16, :]
          const uint64_t ad = make_smem_desc(sds + (k >> 2) * BLK + (k & 3) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc(sk + k * 2048, 8192, 1024);
          umma_bf16(tmem_base + 320, ad, bd, idesc_q, k > 0);
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          // dK[keys x 64] += dS^T[:, 16k rows of queries] Q[16k.., :]   (A = dS tile read MN-major, 64-key blocks BLK apart)
          const uint64_t ad = make_smem_desc(sds + k * 2048, BLK, 1024);
          const uint64_t bd = make_smem_desc(sq + k * 2048, 8192, 1024);
          umma_bf16(tmem_base + 384, ad, bd, idesc_t, k > 0);
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          const uint64_t ad = make_smem_desc(sp + k * 2048, BLK, 1024);
          const uint64_t bd = make_smem_desc(sdo + k * 2048, 8192, 1024);
          umma_bf16(tmem_base + 448, ad, bd, idesc_t, k > 0);
        }
        umma_commit(&out_full[g]);
        umma_commit(&ld_empty[i % LS]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int g = (warp - 2) >> 2;                     // warpgroup = pipeline stage
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const float sl2 = a.scale * 1.44269504088896f;
    const bool valid = r < a.N;
    uint8_t* Pt = smem + SM::OFF_PD + g * 4 * BLK;
    uint8_t* dSt = Pt + 2 * BLK;
    for (int i = g, j = 0; i < n_mine; i += 2, ++j) {
      const int it = blockIdx.x + i * gridDim.x;
      const int b = it / a.H, h = it % a.H;
      mbar_wait(&sdp_full[g], (uint32_t)(j & 1));
      tc_fence_after();
      const uint32_t ts = tmem_base + g * 160 + lane_off;
      uint32_t sr[NK][16];
#pragma unroll
      for (int k = 0; k < NK; ++k) tmem_ld16_nowait(ts + 16 * k, sr[k]);
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int q = 0; q < 16; ++q) if (16 * k + q < a.N) mx = fmaxf(mx, __uint_as_float(sr[k][q]));
      const float mb = mx * sl2;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float e = (valid && 16 * k + q < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[k][q]), sl2, -mb)) : 0.f;
          sr[k][q] = __float_as_uint(e);
          sum += e;
        }
      const float inv = valid ? 1.0f / sum : 0.f;
      // delta = rowsum(dO o O) = sum_j P_j dP_j (O = P V): from the accumulators, no global read of O / dO
      float delta = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        uint32_t dpr[16];
        tmem_ld16_nowait(ts + 80 + 16 * k, dpr);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) delta = fmaf(__uint_as_float(sr[k][q]), __uint_as_float(dpr[q]), delta);
      }
      const float nds = -delta * inv;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        uint32_t dpr[16];
        tmem_ld16_nowait(ts + 80 + 16 * k, dpr);
        tmem_ld_wait();
        if (k == NK - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sdp_free[g]);     // S / dP of this stage may be overwritten (item i+2)
        }
        if (r < KPAD) {                                 // the tiles hold KPAD rows only
          uint32_t wp[8], wd[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float p0 = __uint_as_float(sr[k][2 * q]) * inv, p1 = __uint_as_float(sr[k][2 * q + 1]) * inv;
            const float d0 = p0 * (__uint_as_float(dpr[2 * q]) + nds) * a.scale;
            const float d1 = p1 * (__uint_as_float(dpr[2 * q + 1]) + nds) * a.scale;
            wp[q] = pack2(p0, p1);
            wd[q] = pack2(d0, d1);
          }
          const int ch = (k & 3) * 2;
          uint8_t* pb = Pt + (k >> 2) * BLK;
          uint8_t* db = dSt + (k >> 2) * BLK;
          *reinterpret_cast<uint4*>(pb + sw128_off(r, ch)) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
          *reinterpret_cast<uint4*>(pb + sw128_off(r, ch + 1)) = make_uint4(wp[4], wp[5], wp[6], wp[7]);
          *reinterpret_cast<uint4*>(db + sw128_off(r, ch)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          *reinterpret_cast<uint4*>(db + sw128_off(r, ch + 1)) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_ready[g]);
      mbar_wait(&out_full[g], (uint32_t)(j & 1));
      tc_fence_after();
      bf16* dst = a.dQKV + ((int64_t)b * a.N + r) * (3 * inner) + h * DH;
      const uint32_t to = tmem_base + 320 + lane_off;
#pragma unroll 1
      for (int w = 0; w < 3; ++w) {          // dQ, dK, dV rows (query r / key r)
        uint32_t orr[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16_nowait(to + w * 64 + 16 * c, orr[c]);
        tmem_ld_wait();
        if (w == 2) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&out_free[g]);     // the other warpgroup's item may now use dQ / dK / dV
        }
        if (valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t u[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) u[q] = pack2(__uint_as_float(orr[c][2 * q]), __uint_as_float(orr[c][2 * q + 1]));
            reinterpret_cast<uint4*>(dst + w * inner + 16 * c)[0] = make_uint4(u[0], u[1], u[2], u[3]);
            reinterpret_cast<uint4*>(dst + w * inner + 16 * c)[1] = make_uint4(u[4], u[5], u[6], u[7]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ host
static bool g_bwd2_enabled = true;      // set_option "attn_bwd2"
static bool eligible(int N, int dh, const void* p0, int64_t ld) {
  return tc::g_tc_enabled && dh == DH && N <= 128 && (ld % 8) == 0 && (((uintptr_t)p0) & 15) == 0;
}

static void fwd(const bf16* QKV, bf16* O, int B, int N, int H, cudaStream_t st) {
  const int inner = H * DH;
  const int64_t T = (int64_t)B * N;
  AttnArgs a;
  a.B = B; a.N = N; a.H = H; a.KPAD = (N + 15) / 16 * 16;
  a.scale = 1.0f / sqrtf((float)DH);
  a.O = O; a.Oin = nullptr; a.dO = nullptr; a.dQKV = nullptr;
  CUtensorMap tkv = make_map(QKV, 3 * inner, T, 3 * inner, 64, a.KPAD);
  const int grid = std::min(B * H, sm_count());
  if (skip_mask() & SKIP_ATTN_FWD) return;
  switch (a.KPAD / 16) {
#define DG_ATTN_F(NK_)                                                                                              \
    case NK_: {                                                                                                     \
      constexpr int smem = FwdSmem<NK_>::TOTAL;                                                                     \
      static DevOnce attr;                                                                                     \
      if (attr.first()) {                                                                                                  \
        DG_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<NK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  \
      }                                                                                                             \
      launch_k(attn_fwd_tc_kernel<NK_>, grid, FWD_THREADS, smem, st, tkv, a);                                       \
    } break;
    DG_ATTN_F(1) DG_ATTN_F(2) DG_ATTN_F(3) DG_ATTN_F(4) DG_ATTN_F(5) DG_ATTN_F(6) DG_ATTN_F(7) DG_ATTN_F(8)
#undef DG_ATTN_F
    default: fail(DGVIT_ERR_ARG, "attention: N=%d not supported by the tensor-core kernel", N);
  }
  DG_LAUNCH_CHECK();
}

static void bwd(const bf16* QKV, const bf16* O, const bf16* dO, bf16* dQKV, int B, int N, int H, cudaStream_t st) {
  const int inner = H * DH;
  const int64_t T = (int64_t)B * N;
  AttnArgs a;
  a.B = B; a.N = N; a.H = H; a.KPAD = (N + 15) / 16 * 16;
  a.scale = 1.0f / sqrtf((float)DH);
  a.O = nullptr; a.Oin = O; a.dO = dO; a.dQKV = dQKV;
  CUtensorMap tq = make_map(QKV, 3 * inner, T, 3 * inner, 64, 128);
  CUtensorMap tkv = make_map(QKV, 3 * inner, T, 3 * inner, 64, a.KPAD);
  CUtensorMap tdo = make_map(dO, inner, T, inner, 64, 128);
  const int smem = (BWD_LOAD_STAGES * 4 + 4) * TILE + 256 + 1024;
  const int grid = std::min(B * H, sm_count());
  if (skip_mask() & SKIP_ATTN_BWD) return;
  if (g_bwd2_enabled && a.KPAD <= 80) {         // pipelined kernel (two softmax warpgroups)
    CUtensorMap tdo2 = make_map(dO, inner, T, inner, 64, a.KPAD);
    switch (a.KPAD / 16) {
#define DG_ATTN_B2(NK_)                                                                                                    \
      case NK_: {                                                                                                          \
        constexpr int smem2 = Bwd2Smem<NK_>::TOTAL;                                                                        \
        static DevOnce attr2;                                                                                         \
        if (attr2.first()) {                                                                                                      \
          DG_CUDA(cudaFuncSetAttribute(attn_bwd2_tc_kernel<NK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));     \
        }                                                                                                                  \
        launch_k(attn_bwd2_tc_kernel<NK_>, grid, BWD2_THREADS, smem2, st, tkv, tdo2, a);                                   \
      } break;
      DG_ATTN_B2(1) DG_ATTN_B2(2) DG_ATTN_B2(3) DG_ATTN_B2(4) DG_ATTN_B2(5)
#undef DG_ATTN_B2
    }
    DG_LAUNCH_CHECK();
    return;
  }
  switch (a.KPAD / 16) {
#define DG_ATTN_B(NK_)                                                                                              \
    case NK_: {                                                                                                     \
      static DevOnce attr;                                                                                     \
      if (attr.first()) {                                                                                                  \
        DG_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<NK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  \
      }                                                                                                             \
      launch_k(attn_bwd_tc_kernel<NK_>, grid, BWD_THREADS, smem, st, tq, tkv, tdo, a);                              \
    } break;
    DG_ATTN_B(1) DG_ATTN_B(2) DG_ATTN_B(3) DG_ATTN_B(4) DG_ATTN_B(5) DG_ATTN_B(6) DG_ATTN_B(7) DG_ATTN_B(8)
#undef DG_ATTN_B
    default: fail(DGVIT_ERR_ARG, "attention: N=%d not supported by the tensor-core kernel", N);
  }
  DG_LAUNCH_CHECK();
}

}  // namespace attn
}  // namespace dgvit
