// attn_long_tc.cuh — fused multi-head attention for sequences longer than one 128-row UMMA tile (128 < N <= 384;
// the 2x-resolution DGViT variant has 257 tokens: BASELINE config 5, vn/GoalFormer.py:71-81,124-142), sm_100a.
// Same building blocks as attn_tc.cuh (TMA-staged 128B-swizzled tiles, tcgen05.mma into TMEM, one thread per query
// row for the fp32 softmax), with the key axis walked in 128-key chunks:
//
//   forward  (item = (sample, head); K, V resident in smem, query tiles streamed):
//       S[128 x KP] = Q_t K^T in TMEM (all keys: no online rescaling) ; row max over TMEM ; per 128-key chunk
//       P_c = exp(S_c - max) -> smem (double-buffered) -> O += P_c V_c ; O / rowsum -> HBM ; LSE -> HBM (for the backward)
//   backward, two launches (no atomics, fixed summation order):
//     dQ    (item = (sample, head)):            per query tile, per key chunk: S_c, dP_c -> P_c = exp(S_c - LSE),
//           dS_c = P_c (dP_c - delta) scale -> smem -> dQ += dS_c K_c ; delta = rowsum(dO o O) -> HBM
//     dK,dV (item = (sample, head, key tile)):  per query tile: S, dP -> P, dS tiles -> dV += P^T dO_t ; dK += dS^T Q_t
// The second launch recomputes S / dP (7 GEMM units instead of 5); attention is ~5 % of this model's FLOPs.
#pragma once
#include "attn_tc.cuh"

namespace dgvit {
namespace attnl {

using namespace tc;
using attn::DH;
using attn::TILE;
using attn::fence_async_smem;
using attn::pack2;
using attn::sw128_off;
using attn::tmem_ld16_nowait;
using attn::tmem_ld_wait;

constexpr int MAX_TILES = 3;                 // key / query tiles of 128 rows (N <= 384)
constexpr int SW = 8;                        // softmax warps: two threads per row of the tile, each owning every other
                                             // 64-column block of the scores (one warpgroup = one warp per scheduler)
constexpr int THREADS = 64 + SW * 32;        // TMA warp, MMA warp, softmax warps
__device__ __forceinline__ void sm_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SW * 32) : "memory"); }

struct LongArgs {
  int B, N, H, KP, NT;       // KP = N rounded up to 16 (key columns multiplied), NT = ceil(N / 128)
  float scale;
  bf16* O;                   // fwd: [T, inner]
  float* lse;                // [B, H, N] log2-domain log-sum-exp of the scaled scores (fwd writes, bwd reads)
  float* delta;              // [B, H, N] rowsum(dO o O) (dQ launch writes, dK/dV launch reads)
  const bf16* Oin;           // bwd
  const bf16* dO;            // bwd
  bf16* dQKV;                // bwd: [T, 3*inner]
};

// ---------------------------------------------------------------------------------------------------- forward
namespace f {
constexpr int OFF_K = 0, OFF_V = MAX_TILES * TILE, OFF_Q = 2 * MAX_TILES * TILE;   // Q: ring of 2 tiles
constexpr int OFF_P = OFF_Q + 2 * TILE;                                             // 2 chunk buffers x 2 tiles
constexpr int OFF_X = OFF_P + 4 * TILE;      // [2][128][2] floats: row max / row sum halves of the two threads of a row
constexpr int OFF_BAR = OFF_X + 2048;
constexpr int TOTAL = OFF_BAR + 256 + 1024;
static_assert(TOTAL <= 232448, "smem budget");
constexpr uint32_t T_O = 448;               // S: columns [0, KP) ; O: [448, 512)
}  // namespace f

__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_long_kernel(const __grid_constant__ CUtensorMap tmQKV, const LongArgs a) {
  using namespace f;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* kv_full = bars;            // 1
  uint64_t* kv_empty = kv_full + 1;    // 1
  uint64_t* q_full = kv_empty + 1;     // [2]
  uint64_t* q_empty = q_full + 2;      // [2]
  uint64_t* s_full = q_empty + 2;      // 1
  uint64_t* p_ready = s_full + 1;      // [2] (4 warps)
  uint64_t* p_free = p_ready + 2;      // [2]
  uint64_t* o_full = p_free + 2;       // 1
  uint64_t* s_free = o_full + 1;       // 1 (4 warps): S accumulator read for the last time (end of softmax pass 2)
  uint64_t* o_free = s_free + 1;       // 1 (4 warps): O accumulator drained
  uint32_t* tmem_slot = (uint32_t*)(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inner = a.H * DH;
  const int items = a.B * a.H;
  const int NC = a.NT;                                   // key chunks
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&p_ready[i], SW); mbar_init(&p_free[i], 1); }
    mbar_init(s_full, 1); mbar_init(o_full, 1); mbar_init(s_free, SW); mbar_init(o_free, SW);
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
      uint32_t kvn = 0, qn = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++kvn) {
        const int b = it / a.H, h = it % a.H;
        mbar_wait(kv_empty, (kvn & 1) ^ 1);
        mbar_expect_tx(kv_full, 2 * NC * TILE);
        for (int c = 0; c < NC; ++c) {
          tma_load_2d(smem + OFF_K + c * TILE, &tmQKV, kv_full, inner + h * DH, b * a.N + c * 128);
          tma_load_2d(smem + OFF_V + c * TILE, &tmQKV, kv_full, 2 * inner + h * DH, b * a.N + c * 128);
        }
        for (int t = 0; t < a.NT; ++t, ++qn) {
          const int s = qn & 1;
          mbar_wait(&q_empty[s], ((qn >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[s], TILE);
          tma_load_2d(smem + OFF_Q + s * TILE, &tmQKV, &q_full[s], h * DH, b * a.N + t * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t sk = smem_u32(smem + OFF_K), sv = smem_u32(smem + OFF_V), sp = smem_u32(smem + OFF_P);
      const uint32_t idesc_o = make_idesc(128, DH, false, true);
      uint32_t kvn = 0, qn = 0, tn = 0, pn[2] = {0, 0};
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++kvn) {
        mbar_wait(kv_full, kvn & 1);
        for (int t = 0; t < a.NT; ++t, ++qn, ++tn) {
          const int s = qn & 1;
          mbar_wait(&q_full[s], (qn >> 1) & 1);
          mbar_wait(s_free, (tn & 1) ^ 1);        // S of the next query tile is formed while the threads still store O
          tc_fence_after();
          const uint64_t qd = make_smem_desc(smem_u32(smem + OFF_Q + s * TILE), 16, 1024);
          for (int c = 0; c < NC; ++c) {
            const int nk = min(128, a.KP - 128 * c);
            const uint32_t idesc_s = make_idesc(128, nk, false, false);
            const uint64_t kd = make_smem_desc(sk + c * TILE, 16, 1024);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              umma_bf16(tmem_base + 128 * c, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k > 0);
          }
          umma_commit(s_full);
          umma_commit(&q_empty[s]);
          for (int c = 0; c < NC; ++c) {
            const int pb = c & 1;
            const int ks = min(128, a.KP - 128 * c) / 16;
            mbar_wait(&p_ready[pb], pn[pb] & 1);
            if (c == 0) mbar_wait(o_free, (tn & 1) ^ 1);    // the previous tile's O has been read
            tc_fence_after();
            for (int k = 0; k < ks; ++k) {
              const uint64_t pd = make_smem_desc(sp + pb * 2 * TILE + (k >> 2) * TILE + (k & 3) * 32, 16, 1024);
              const uint64_t vd = make_smem_desc(sv + c * TILE + k * 2048, 8192, 1024);
              umma_bf16(tmem_base + T_O, pd, vd, idesc_o, (c > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&p_free[pb]);
            ++pn[pb];
          }
          umma_commit(o_full);
        }
        umma_commit(kv_empty);
      }
    }
  } else {
    const int quad = warp & 3, hh = (warp - 2) >> 2;       // TMEM lane quadrant ; which 64-column blocks of a row
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const float sl2 = a.scale * 1.44269504088896f;
    float* xmax = reinterpret_cast<float*>(smem + OFF_X);
    float* xsum = xmax + 256;
    uint32_t tn = 0, pn[2] = {0, 0};
    const int nkb = a.KP / 16;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int b = it / a.H, h = it % a.H;
      for (int t = 0; t < a.NT; ++t, ++tn) {
        const int row = t * 128 + r;
        const bool valid = row < a.N;
        mbar_wait(s_full, tn & 1);
        tc_fence_after();
        const uint32_t ts = tmem_base + lane_off;
        // pass 1: row maximum over the N real key columns (this thread: the 64-column blocks hh, hh + 2, ...)
        float mx = -INFINITY;
        for (int k4 = hh * 4; k4 < nkb; k4 += 8) {
          uint32_t s0[16], s1[16], s2[16], s3[16];
          tmem_ld16_nowait(ts + 16 * k4, s0);
          if (k4 + 1 < nkb) tmem_ld16_nowait(ts + 16 * (k4 + 1), s1);
          if (k4 + 2 < nkb) tmem_ld16_nowait(ts + 16 * (k4 + 2), s2);
          if (k4 + 3 < nkb) tmem_ld16_nowait(ts + 16 * (k4 + 3), s3);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (16 * k4 + i < a.N) mx = fmaxf(mx, __uint_as_float(s0[i]));
            if (k4 + 1 < nkb && 16 * (k4 + 1) + i < a.N) mx = fmaxf(mx, __uint_as_float(s1[i]));
            if (k4 + 2 < nkb && 16 * (k4 + 2) + i < a.N) mx = fmaxf(mx, __uint_as_float(s2[i]));
            if (k4 + 3 < nkb && 16 * (k4 + 3) + i < a.N) mx = fmaxf(mx, __uint_as_float(s3[i]));
          }
        }
        xmax[r * 2 + hh] = mx;
        sm_bar();
        mx = fmaxf(xmax[r * 2], xmax[r * 2 + 1]);
        const float mb = mx * sl2;
        float sum = 0.f;
        // pass 2: P chunk by chunk (this thread: 64 of the chunk's 128 columns)
        for (int c = 0; c < NC; ++c) {
          const int pb = c & 1;
          const int nb = min(128, a.KP - 128 * c) / 16;
          uint8_t* P = smem + OFF_P + pb * 2 * TILE;
          mbar_wait(&p_free[pb], (pn[pb] & 1) ^ 1);        // the MMAs that read this buffer two chunks ago have retired
          for (int kb = hh * 4; kb < min(nb, hh * 4 + 4); ++kb) {
            uint32_t sr[16];
            tmem_ld16_nowait(ts + 128 * c + 16 * kb, sr);
            tmem_ld_wait();
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c0 = 128 * c + 16 * kb + 2 * i;
              const float e0 = (valid && c0 < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[2 * i]), sl2, -mb)) : 0.f;
              const float e1 = (valid && c0 + 1 < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[2 * i + 1]), sl2, -mb)) : 0.f;
              w[i] = pack2(e0, e1);
              sum += __uint_as_float(w[i] << 16) + __uint_as_float(w[i] & 0xffff0000u);   // what the tensor core sees
            }
            uint8_t* blk = P + (kb >> 2) * TILE;
            const int ch = (kb & 3) * 2;
            *reinterpret_cast<uint4*>(blk + sw128_off(r, ch)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(blk + sw128_off(r, ch + 1)) = make_uint4(w[4], w[5], w[6], w[7]);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[pb]);
          ++pn[pb];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);              // S has been read for the last time
        xsum[r * 2 + hh] = sum;
        sm_bar();
        sum = xsum[r * 2] + xsum[r * 2 + 1];
        mbar_wait(o_full, tn & 1);
        tc_fence_after();
        const float inv = valid ? 1.0f / sum : 0.f;
        uint32_t orr[2][16];
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld16_nowait(ts + T_O + hh * 32 + 16 * c, orr[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);              // O is in registers: the next tile's P V products may start
        if (valid) {
          bf16* dst = a.O + ((int64_t)b * a.N + row) * inner + h * DH + hh * 32;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = pack2(__uint_as_float(orr[c][2 * i]) * inv, __uint_as_float(orr[c][2 * i + 1]) * inv);
            reinterpret_cast<uint4*>(dst + 16 * c)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4*>(dst + 16 * c)[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
          if (hh == 0 && a.lse) a.lse[((int64_t)b * a.H + h) * a.N + row] = mb + log2f(sum);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------- backward: dQ
namespace q {
constexpr int OFF_K = 0, OFF_V = MAX_TILES * TILE, OFF_Q = 2 * MAX_TILES * TILE, OFF_DO = OFF_Q + TILE;
constexpr int OFF_DS = OFF_DO + TILE;                     // 2 chunk buffers x 2 tiles
constexpr int OFF_BAR = OFF_DS + 4 * TILE;
constexpr int TOTAL = OFF_BAR + 256 + 1024;
static_assert(TOTAL <= 232448, "smem budget");
constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256;
}  // namespace q

__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dq_long_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmdO, const LongArgs a) {
  using namespace q;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* kv_full = bars;            // 1
  uint64_t* kv_empty = kv_full + 1;    // 1
  uint64_t* qdo_full = kv_empty + 1;   // 1
  uint64_t* qdo_empty = qdo_full + 1;  // 1
  uint64_t* sdp_full = qdo_empty + 1;  // 1
  uint64_t* sdp_free = sdp_full + 1;   // 1 (4 warps)
  uint64_t* ds_ready = sdp_free + 1;   // [2] (4 warps)
  uint64_t* ds_free = ds_ready + 2;    // [2]
  uint64_t* dq_full = ds_free + 2;     // 1
  uint64_t* dq_free = dq_full + 1;     // 1 (4 warps)
  uint32_t* tmem_slot = (uint32_t*)(dq_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inner = a.H * DH;
  const int items = a.B * a.H;
  const int NC = a.NT;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1); mbar_init(qdo_full, 1); mbar_init(qdo_empty, 1);
    mbar_init(sdp_full, 1); mbar_init(sdp_free, SW); mbar_init(dq_full, 1); mbar_init(dq_free, SW);
    for (int i = 0; i < 2; ++i) { mbar_init(&ds_ready[i], SW); mbar_init(&ds_free[i], 1); }
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
      uint32_t kvn = 0, tn = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++kvn) {
        const int b = it / a.H, h = it % a.H;
        mbar_wait(kv_empty, (kvn & 1) ^ 1);
        mbar_expect_tx(kv_full, 2 * NC * TILE);
        for (int c = 0; c < NC; ++c) {
          tma_load_2d(smem + OFF_K + c * TILE, &tmQKV, kv_full, inner + h * DH, b * a.N + c * 128);
          tma_load_2d(smem + OFF_V + c * TILE, &tmQKV, kv_full, 2 * inner + h * DH, b * a.N + c * 128);
        }
        for (int t = 0; t < a.NT; ++t, ++tn) {
          mbar_wait(qdo_empty, (tn & 1) ^ 1);
          mbar_expect_tx(qdo_full, 2 * TILE);
          tma_load_2d(smem + OFF_Q, &tmQKV, qdo_full, h * DH, b * a.N + t * 128);
          tma_load_2d(smem + OFF_DO, &tmdO, qdo_full, h * DH, b * a.N + t * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t sk = smem_u32(smem + OFF_K), sv = smem_u32(smem + OFF_V), sds = smem_u32(smem + OFF_DS);
      const uint32_t idesc_q = make_idesc(128, DH, false, true);         // dQ = dS K (A K-major, B MN-major)
      const uint64_t qd = make_smem_desc(smem_u32(smem + OFF_Q), 16, 1024), dod = make_smem_desc(smem_u32(smem + OFF_DO), 16, 1024);
      uint32_t kvn = 0, tn = 0, cn = 0, dn[2] = {0, 0};
      auto issue_sdp = [&](int c) {                                       // S_c = Q K_c^T ; dP_c = dO V_c^T
        mbar_wait(sdp_free, (cn & 1) ^ 1);
        tc_fence_after();
        const int nk = min(128, a.KP - 128 * c);
        const uint32_t idesc_s = make_idesc(128, nk, false, false);
        const uint64_t kd = make_smem_desc(sk + c * TILE, 16, 1024), vd = make_smem_desc(sv + c * TILE, 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + T_S, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + T_DP, dod + (uint64_t)(k * 2), vd + (uint64_t)(k * 2), idesc_s, k > 0);
        umma_commit(sdp_full);
        ++cn;
      };
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++kvn) {
        mbar_wait(kv_full, kvn & 1);
        for (int t = 0; t < a.NT; ++t, ++tn) {
          mbar_wait(qdo_full, tn & 1);
          tc_fence_after();
          issue_sdp(0);
          for (int c = 0; c < NC; ++c) {
            if (c + 1 < NC) issue_sdp(c + 1);
            const int db = c & 1;
            const int ks = min(128, a.KP - 128 * c) / 16;
            mbar_wait(&ds_ready[db], dn[db] & 1);
            if (c == 0) mbar_wait(dq_free, (tn & 1) ^ 1);      // the previous tile's dQ has been read
            tc_fence_after();
            for (int k = 0; k < ks; ++k) {       // dQ[128 x 64] += dS_c[:, 16k..] K_c[16k.., :]
              const uint64_t ad = make_smem_desc(sds + db * 2 * TILE + (k >> 2) * TILE + (k & 3) * 32, 16, 1024);
              const uint64_t bd = make_smem_desc(sk + c * TILE + k * 2048, 8192, 1024);
              umma_bf16(tmem_base + T_DQ, ad, bd, idesc_q, (c > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&ds_free[db]);
            ++dn[db];
          }
          umma_commit(dq_full);
          umma_commit(qdo_empty);
        }
        umma_commit(kv_empty);
      }
    }
  } else {
    const int quad = warp & 3, hh = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const float sl2 = a.scale * 1.44269504088896f;
    uint32_t tn = 0, cn = 0, dn[2] = {0, 0};
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int b = it / a.H, h = it % a.H;
      for (int t = 0; t < a.NT; ++t, ++tn) {
        const int row = t * 128 + r;
        const bool valid = row < a.N;
        // delta = rowsum(dO o O) from global (bf16, 128 B per row each); log-sum-exp of the row from the forward
        float delta = 0.f, lse = 0.f;
        if (valid) {
          const uint4* po = reinterpret_cast<const uint4*>(a.Oin + ((int64_t)b * a.N + row) * inner + h * DH);
          const uint4* pd = reinterpret_cast<const uint4*>(a.dO + ((int64_t)b * a.N + row) * inner + h * DH);
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
          const int64_t si = ((int64_t)b * a.H + h) * a.N + row;
          lse = a.lse[si];
          if (hh == 0) a.delta[si] = delta;
        }
        const uint32_t ts = tmem_base + lane_off;
        for (int c = 0; c < NC; ++c, ++cn) {
          const int db = c & 1;
          const int nb = min(128, a.KP - 128 * c) / 16;
          const int kb0 = hh * 4, kb1 = min(nb, hh * 4 + 4);      // this thread's 64 columns of the chunk
          uint8_t* dS = smem + OFF_DS + db * 2 * TILE;
          mbar_wait(sdp_full, cn & 1);
          tc_fence_after();
          mbar_wait(&ds_free[db], (dn[db] & 1) ^ 1);
          if (kb0 >= kb1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_free);
          }
          for (int kb = kb0; kb < kb1; ++kb) {
            uint32_t sr[16], dpr[16];
            tmem_ld16_nowait(ts + T_S + 16 * kb, sr);
            tmem_ld16_nowait(ts + T_DP + 16 * kb, dpr);
            tmem_ld_wait();
            if (kb == kb1 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(sdp_free);           // S_c / dP_c are in registers: chunk c+1 may overwrite them
            }
            uint32_t wd[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c0 = 128 * c + 16 * kb + 2 * i;
              const float p0 = (valid && c0 < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[2 * i]), sl2, -lse)) : 0.f;
              const float p1 = (valid && c0 + 1 < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[2 * i + 1]), sl2, -lse)) : 0.f;
              wd[i] = pack2(p0 * (__uint_as_float(dpr[2 * i]) - delta) * a.scale, p1 * (__uint_as_float(dpr[2 * i + 1]) - delta) * a.scale);
            }
            uint8_t* blk = dS + (kb >> 2) * TILE;
            const int ch = (kb & 3) * 2;
            *reinterpret_cast<uint4*>(blk + sw128_off(r, ch)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            *reinterpret_cast<uint4*>(blk + sw128_off(r, ch + 1)) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ds_ready[db]);
          ++dn[db];
        }
        mbar_wait(dq_full, tn & 1);
        tc_fence_after();
        uint32_t orr[2][16];
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld16_nowait(ts + T_DQ + hh * 32 + 16 * c, orr[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_free);
        if (valid) {
          bf16* dst = a.dQKV + ((int64_t)b * a.N + row) * (3 * inner) + h * DH + hh * 32;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = pack2(__uint_as_float(orr[c][2 * i]), __uint_as_float(orr[c][2 * i + 1]));
            reinterpret_cast<uint4*>(dst + 16 * c)[0] = make_uint4(u[0], u[1], u[2], u[3]);
            reinterpret_cast<uint4*>(dst + 16 * c)[1] = make_uint4(u[4], u[5], u[6], u[7]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------- backward: dK, dV
namespace kv {
constexpr int OFF_K = 0, OFF_V = TILE, OFF_QDO = 2 * TILE;      // Q | dO ring of 2 stages
constexpr int OFF_P = OFF_QDO + 4 * TILE;                        // P: 2 tiles, dS: 2 tiles
constexpr int OFF_DS = OFF_P + 2 * TILE;
constexpr int OFF_BAR = OFF_DS + 2 * TILE;
constexpr int TOTAL = OFF_BAR + 256 + 1024;
static_assert(TOTAL <= 232448, "smem budget");
constexpr uint32_t T_S = 0, T_DP = 128, T_DK = 256, T_DV = 320;
}  // namespace kv

__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dkdv_long_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmdO, const LongArgs a) {
  using namespace kv;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* kv_full = bars;            // 1
  uint64_t* kv_empty = kv_full + 1;    // 1
  uint64_t* qdo_full = kv_empty + 1;   // [2]
  uint64_t* qdo_empty = qdo_full + 2;  // [2]
  uint64_t* sdp_full = qdo_empty + 2;  // 1
  uint64_t* sdp_free = sdp_full + 1;   // 1 (4 warps)
  uint64_t* pds_ready = sdp_free + 1;  // 1 (4 warps)
  uint64_t* pds_free = pds_ready + 1;  // 1
  uint64_t* out_full = pds_free + 1;   // 1
  uint64_t* out_free = out_full + 1;   // 1 (4 warps)
  uint32_t* tmem_slot = (uint32_t*)(out_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inner = a.H * DH;
  const int items = a.B * a.H * a.NT;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(sdp_full, 1); mbar_init(sdp_free, SW); mbar_init(pds_ready, SW); mbar_init(pds_free, 1);
    mbar_init(out_full, 1); mbar_init(out_free, SW);
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
      uint32_t kvn = 0, qn = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++kvn) {
        const int kt = it % a.NT, bh = it / a.NT;
        const int b = bh / a.H, h = bh % a.H;
        mbar_wait(kv_empty, (kvn & 1) ^ 1);
        mbar_expect_tx(kv_full, 2 * TILE);
        tma_load_2d(smem + OFF_K, &tmQKV, kv_full, inner + h * DH, b * a.N + kt * 128);
        tma_load_2d(smem + OFF_V, &tmQKV, kv_full, 2 * inner + h * DH, b * a.N + kt * 128);
        for (int t = 0; t < a.NT; ++t, ++qn) {
          const int s = qn & 1;
          mbar_wait(&qdo_empty[s], ((qn >> 1) & 1) ^ 1);
          mbar_expect_tx(&qdo_full[s], 2 * TILE);
          tma_load_2d(smem + OFF_QDO + s * 2 * TILE, &tmQKV, &qdo_full[s], h * DH, b * a.N + t * 128);
          tma_load_2d(smem + OFF_QDO + s * 2 * TILE + TILE, &tmdO, &qdo_full[s], h * DH, b * a.N + t * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t sk = smem_u32(smem + OFF_K), sv = smem_u32(smem + OFF_V);
      const uint32_t sp = smem_u32(smem + OFF_P), sds = smem_u32(smem + OFF_DS);
      const uint32_t idesc_t = make_idesc(128, DH, true, true);          // dK, dV (A MN-major, B MN-major)
      const uint64_t kd = make_smem_desc(sk, 16, 1024), vd = make_smem_desc(sv, 16, 1024);
      uint32_t kvn = 0, qn = 0, sn = 0, pn = 0, on = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++kvn, ++on) {
        const int kt = it % a.NT;
        const int nk = min(128, a.KP - 128 * kt);
        const uint32_t idesc_s = make_idesc(128, nk, false, false);
        mbar_wait(kv_full, kvn & 1);
        auto issue_sdp = [&](uint32_t qi) {                // S = Q_t K^T ; dP = dO_t V^T  (query tile in ring stage qi & 1)
          const int s = qi & 1;
          mbar_wait(&qdo_full[s], (qi >> 1) & 1);
          mbar_wait(sdp_free, (sn & 1) ^ 1);
          tc_fence_after();
          const uint32_t sq = smem_u32(smem + OFF_QDO + s * 2 * TILE);
          const uint64_t qd = make_smem_desc(sq, 16, 1024), dod = make_smem_desc(sq + TILE, 16, 1024);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + T_S, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + T_DP, dod + (uint64_t)(k * 2), vd + (uint64_t)(k * 2), idesc_s, k > 0);
          umma_commit(sdp_full);
          ++sn;
        };
        issue_sdp(qn);
        for (int t = 0; t < a.NT; ++t, ++qn, ++pn) {
          if (t + 1 < a.NT) issue_sdp(qn + 1);
          const int s = qn & 1;
          mbar_wait(pds_ready, pn & 1);
          if (t == 0) mbar_wait(out_free, (on & 1) ^ 1);       // the previous item's dK / dV have been read
          tc_fence_after();
          const uint32_t sq = smem_u32(smem + OFF_QDO + s * 2 * TILE), sdo = sq + TILE;
#pragma unroll
          for (int k = 0; k < 8; ++k) {          // 16 query rows per step
            const uint32_t acc = (t > 0 || k > 0) ? 1u : 0u;
            // dV[keys x 64] += P^T[:, 16k query rows] dO_t[16k.., :]   (A = P tile read MN-major, 64-key blocks TILE apart)
            umma_bf16(tmem_base + T_DV, make_smem_desc(sp + k * 2048, TILE, 1024), make_smem_desc(sdo + k * 2048, 8192, 1024), idesc_t, acc);
            umma_bf16(tmem_base + T_DK, make_smem_desc(sds + k * 2048, TILE, 1024), make_smem_desc(sq + k * 2048, 8192, 1024), idesc_t, acc);
          }
          umma_commit(pds_free);
          umma_commit(&qdo_empty[s]);
        }
        umma_commit(out_full);
        umma_commit(kv_empty);
      }
    }
  } else {
    const int quad = warp & 3, hh = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const float sl2 = a.scale * 1.44269504088896f;
    uint32_t sn = 0, pn = 0, on = 0;
    uint8_t* Pt = smem + OFF_P;
    uint8_t* dSt = smem + OFF_DS;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++on) {
      const int kt = it % a.NT, bh = it / a.NT;
      const int b = bh / a.H, h = bh % a.H;
      const int nb = min(128, a.KP - 128 * kt) / 16;
      const int kb0 = hh * 4, kb1 = min(nb, hh * 4 + 4);     // this thread's 64 key columns of the tile
      const uint32_t ts = tmem_base + lane_off;
      for (int t = 0; t < a.NT; ++t, ++sn, ++pn) {
        const int row = t * 128 + r;                       // query row of this thread
        const bool valid = row < a.N;
        float lse = 0.f, delta = 0.f;
        if (valid) {
          const int64_t si = ((int64_t)b * a.H + h) * a.N + row;
          lse = a.lse[si];
          delta = a.delta[si];
        }
        mbar_wait(sdp_full, sn & 1);
        tc_fence_after();
        mbar_wait(pds_free, (pn & 1) ^ 1);                 // the MMAs of the previous query tile have consumed the P / dS tiles
        if (kb0 >= kb1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sdp_free);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          uint32_t sr[16], dpr[16];
          tmem_ld16_nowait(ts + T_S + 16 * kb, sr);
          tmem_ld16_nowait(ts + T_DP + 16 * kb, dpr);
          tmem_ld_wait();
          if (kb == kb1 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_free);
          }
          uint32_t wp[8], wd[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c0 = 128 * kt + 16 * kb + 2 * i;
            const float p0 = (valid && c0 < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[2 * i]), sl2, -lse)) : 0.f;
            const float p1 = (valid && c0 + 1 < a.N) ? ex2_approx(fmaf(__uint_as_float(sr[2 * i + 1]), sl2, -lse)) : 0.f;
            wp[i] = pack2(p0, p1);
            wd[i] = pack2(p0 * (__uint_as_float(dpr[2 * i]) - delta) * a.scale, p1 * (__uint_as_float(dpr[2 * i + 1]) - delta) * a.scale);
          }
          const int ch = (kb & 3) * 2;
          uint8_t* pb = Pt + (kb >> 2) * TILE;
          uint8_t* db = dSt + (kb >> 2) * TILE;
          *reinterpret_cast<uint4*>(pb + sw128_off(r, ch)) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
          *reinterpret_cast<uint4*>(pb + sw128_off(r, ch + 1)) = make_uint4(wp[4], wp[5], wp[6], wp[7]);
          *reinterpret_cast<uint4*>(db + sw128_off(r, ch)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          *reinterpret_cast<uint4*>(db + sw128_off(r, ch + 1)) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_ready);
      }
      // dK (threads hh = 0) / dV (hh = 1) row of key (kt * 128 + r)
      mbar_wait(out_full, on & 1);
      tc_fence_after();
      const int key = kt * 128 + r;
      bf16* dst = a.dQKV + ((int64_t)b * a.N + key) * (3 * inner) + (1 + hh) * inner + h * DH;
      uint32_t orr[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16_nowait(ts + T_DK + hh * 64 + 16 * c, orr[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(out_free);
      if (key < a.N) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t u[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) u[i] = pack2(__uint_as_float(orr[c][2 * i]), __uint_as_float(orr[c][2 * i + 1]));
          reinterpret_cast<uint4*>(dst + 16 * c)[0] = make_uint4(u[0], u[1], u[2], u[3]);
          reinterpret_cast<uint4*>(dst + 16 * c)[1] = make_uint4(u[4], u[5], u[6], u[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------- host
static bool g_enabled = true;           // set_option "attn_long"
static bool eligible(int N, int dh, const void* p0, int64_t ld) {
  return tc::g_tc_enabled && g_enabled && dh == DH && N > 128 && N <= 128 * MAX_TILES && (ld % 8) == 0 && (((uintptr_t)p0) & 15) == 0;
}
// floats of statistics scratch per call: LSE and delta, [B, H, N] each
static size_t stats_floats(int B, int N, int H) { return (size_t)2 * B * H * N; }

static LongArgs make_args(int B, int N, int H) {
  LongArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.H = H; a.KP = (N + 15) / 16 * 16; a.NT = (N + 127) / 128;
  a.scale = 1.0f / sqrtf((float)DH);
  return a;
}

static void fwd(const bf16* QKV, bf16* O, float* lse, int B, int N, int H, cudaStream_t st) {
  const int inner = H * DH;
  const int64_t T = (int64_t)B * N;
  LongArgs a = make_args(B, N, H);
  a.O = O; a.lse = lse;
  CUtensorMap tm = make_map(QKV, 3 * inner, T, 3 * inner, 64, 128);
  static DevOnce attr;
  if (attr.first()) DG_CUDA(cudaFuncSetAttribute(attn_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, f::TOTAL));
  if (skip_mask() & SKIP_ATTN_FWD) return;
  launch_k(attn_fwd_long_kernel, std::min(B * H, sm_count()), THREADS, f::TOTAL, st, tm, a);
  DG_LAUNCH_CHECK();
}

static void bwd(const bf16* QKV, const bf16* O, const bf16* dO, bf16* dQKV, float* lse, float* delta, int B, int N, int H,
                cudaStream_t st) {
  const int inner = H * DH;
  const int64_t T = (int64_t)B * N;
  LongArgs a = make_args(B, N, H);
  a.Oin = O; a.dO = dO; a.dQKV = dQKV; a.lse = lse; a.delta = delta;
  CUtensorMap tm = make_map(QKV, 3 * inner, T, 3 * inner, 64, 128);
  CUtensorMap tdo = make_map(dO, inner, T, inner, 64, 128);
  static DevOnce attr;
  if (attr.first()) {
    DG_CUDA(cudaFuncSetAttribute(attn_bwd_dq_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, q::TOTAL));
    DG_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kv::TOTAL));
  }
  if (skip_mask() & SKIP_ATTN_BWD) return;
  launch_k(attn_bwd_dq_long_kernel, std::min(B * H, sm_count()), THREADS, q::TOTAL, st, tm, tdo, a);
  DG_LAUNCH_CHECK();
  launch_k(attn_bwd_dkdv_long_kernel, std::min(B * H * a.NT, sm_count()), THREADS, kv::TOTAL, st, tm, tdo, a);
  DG_LAUNCH_CHECK();
}

}  // namespace attnl
}  // namespace dgvit
