// dgvit.cu — host orchestration + C ABI of libdgvit.so (see include/dgvit.h).
//
// Data layout in HBM (per network): one flat fp32 parameter arena in the reference's
// registration order (+ grads, Adam m/v, bf16 shadow arenas of the same geometry);
// activations live in a caller-provided workspace carved deterministically from
// (cfg, B, precision), so a backward call finds what its forward saved.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "kernels.cuh"
#include "heads.cuh"
#ifdef DGVIT_WITH_TC
#include "gemm_tc.cuh"
#include "attn_tc.cuh"
#include "attn_long_tc.cuh"
#include "mlp_tc.cuh"
#include "patch_tc.cuh"
#endif

namespace dgvit {

// ------------------------------------------------------------------ layout
static void make_layout(const dgvit_cfg& c, dgvit_layout& L) {
  DG_REQUIRE(c.depth >= 1 && c.depth <= DGVIT_MAX_DEPTH, "depth %d out of range", c.depth);
  DG_REQUIRE(c.dim % 32 == 0 && c.dim >= 32 && c.dim <= 256, "dim %d must be a multiple of 32 in [32,256]", c.dim);
  DG_REQUIRE(c.img_h % c.patch_h == 0 && c.img_w % c.patch_w == 0, "image not divisible by patch");
  DG_REQUIRE(c.heads >= 1 && c.dim_head >= 1 && c.mlp_dim >= 1 && c.n_act >= 1 && c.n_pstate >= 1, "bad cfg");
  memset(&L, 0, sizeof(L));
  int64_t off = 0;
  auto take = [&](int64_t n) {
    int64_t o = off;
    off += (n + DGVIT_ALIGN_FLOATS - 1) / DGVIT_ALIGN_FLOATS * DGVIT_ALIGN_FLOATS;
    return o;
  };
  const int64_t D = c.dim, P = (c.img_h / c.patch_h) * (c.img_w / c.patch_w), pd = c.patch_h * c.patch_w;
  const int64_t inner = (int64_t)c.heads * c.dim_head, M = c.mlp_dim;
  L.pos = take((P + 1) * D);
  L.cls = take(D);
  L.rms_g = take(D);
  L.patch_w = take(D * pd);
  L.patch_b = take(D);
  for (int l = 0; l < c.depth; ++l) {
    dgvit_block_layout& b = L.block[l];
    b.ln1_w = take(D); b.ln1_b = take(D);
    b.qkv_w = take(3 * inner * D);
    b.out_w = take(D * inner); b.out_b = take(D);
    b.ln2_w = take(D); b.ln2_b = take(D);
    b.fc1_w = take(M * D); b.fc1_b = take(M);
    b.fc2_w = take(D * M); b.fc2_b = take(D);
  }
  L.mlp_head_ln_w = take(D); L.mlp_head_ln_b = take(D);
  L.mlp_head_w = take(2 * D); L.mlp_head_b = take(2);
  const int64_t mlp_head_end = off;
  L.n_skip = 0;
  L.skip_begin[L.n_skip] = L.cls; L.skip_end[L.n_skip++] = L.rms_g;
  L.skip_begin[L.n_skip] = L.mlp_head_ln_w; L.skip_end[L.n_skip++] = mlp_head_end;
  if (c.kind == DGVIT_ACTOR) {
    L.embed_w = take(D * c.n_pstate); L.embed_b = take(D);
    L.fc1_w = take(128 * D); L.fc1_b = take(128);
    L.fc2_w = take(128 * 128); L.fc2_b = take(128);
    L.mean_w = take(c.n_act * 128); L.mean_b = take(c.n_act);
    L.lstd_w = take(c.n_act * 128); L.lstd_b = take(c.n_act);
    L.alpha_grad_slot = take(1);
    L.skip_begin[L.n_skip] = L.alpha_grad_slot; L.skip_end[L.n_skip++] = off;
  } else {
    L.conv1_w = take(16 * 4 * 25); L.conv1_b = take(16);
    L.conv2_w = take(64 * 16 * 25); L.conv2_b = take(64);
    L.conv3_w = take(256 * 64 * 25); L.conv3_b = take(256);
    L.skip_begin[L.n_skip] = L.conv1_w; L.skip_end[L.n_skip++] = off;
    L.fc1_w = take(128 * (D + c.n_act)); L.fc1_b = take(128);
    L.fc2_w = take(32 * 128); L.fc2_b = take(32);
    L.fc3_w = take(c.n_act * 32); L.fc3_b = take(c.n_act);
    L.embed_w = take(D * c.n_pstate); L.embed_b = take(D);
    L.fc11_w = take(128 * (D + c.n_act)); L.fc11_b = take(128);
    L.fc21_w = take(32 * 128); L.fc21_b = take(32);
    L.fc31_w = take(c.n_act * 32); L.fc31_b = take(c.n_act);
  }
  L.total = off;
}

struct Dims {
  int B, N, P, D, H, dh, inner, M, pd, L, na, nps;
  int64_t T;
  Dims(const dgvit_cfg& c, int B_) {
    B = B_; P = (c.img_h / c.patch_h) * (c.img_w / c.patch_w); N = P + 1; D = c.dim; H = c.heads;
    dh = c.dim_head; inner = H * dh; M = c.mlp_dim; pd = c.patch_h * c.patch_w; L = c.depth;
    na = c.n_act; nps = c.n_pstate; T = (int64_t)B * N;
  }
};

// weights in the arithmetic dtype of the contractions
template <typename A> struct WSel;
template <> struct WSel<float> {
  static const float* w(const dgvit_net& n, int64_t off) { return n.params + off; }
};
template <> struct WSel<bf16> {
  static const bf16* w(const dgvit_net& n, int64_t off) { return (const bf16*)n.shadow + off; }
};
// f16 copy of a tensor: second half of the shadow arena (see store_shadows4)
static const void* w_f16(const dgvit_net& n, const dgvit_layout& L, int64_t off) { return n.shadow + L.total + off; }

// ------------------------------------------------------------------ GEMM dispatch
// single-query-row attention in the last block (set_option "attention_row0": bit 0 forward, bit 1 backward).  Measured
// at B=256 (ncu, cold): backward 22 us vs 38 us for the full tcgen05 kernel, forward 16 us vs 13.5 us; inside the graph-replayed
// step both help (same-box A/B: off 104.8 k, backward only 106.2 k, both 106.6 k samples/s).
static int g_row0_mode = 3;
static thread_local int g_cur_tag = PROF_NONE;
struct TagScope {
  int prev;
  explicit TagScope(int t) : prev(g_cur_tag) { g_cur_tag = t; }
  ~TagScope() { g_cur_tag = prev; }
};

template <typename TA, typename TB, typename TC>
static void gemm(const GemmArgs& g, cudaStream_t st) {
  const double fl = 2.0 * g.M * g.N * g.K;
  ProfScope ps(g_cur_tag, fl, 0.0, st);
  ProfScope ps_all(PROF_GEMM_ALL, fl, 0.0, st);
#ifdef DGVIT_WITH_TC
  if (gemm_tc_try<TA, TB, TC>(g, st)) return;
#endif
  gemm_simt<TA, TB, TC>(g, st);
}

// y[R,N] = x[R,K] W[N,K]^T (+epilogue)
template <typename TA, typename TB, typename TC>
static void linear_fwd(const TA* x, const TB* W, TC* y, int64_t R, int N, int K, int epi, const float* bias,
                       cudaStream_t st, const float* resid = nullptr, void* C2 = nullptr, int64_t ldc = -1,
                       int64_t ldx = -1, int64_t ldr = -1, int skip_pre = 0) {
  GemmArgs g;
  g.skip_pre = skip_pre;
  g.M = (int)R; g.N = N; g.K = K;
  g.A = x; g.a_sm = ldx < 0 ? K : ldx; g.a_sk = 1;
  g.B = W; g.b_sk = 1; g.b_sn = K;
  g.C = y; g.ldc = ldc < 0 ? N : ldc;
  g.epi = epi; g.bias = bias; g.resid = resid; g.ldr = ldr < 0 ? g.ldc : ldr; g.C2 = C2;
  gemm<TA, TB, TC>(g, st);
}
// dx[R,K] = dy[R,N] W[N,K]  (+epilogue)
template <typename TA, typename TB, typename TC>
static void linear_bwd_x(const TA* dy, const TB* W, TC* dx, int64_t R, int N, int K, int epi, const void* aux,
                         int64_t ldaux, cudaStream_t st, int64_t ldy = -1, const float* resid = nullptr,
                         int64_t ldc = -1) {
  GemmArgs g;
  g.M = (int)R; g.N = K; g.K = N;
  g.A = dy; g.a_sm = ldy < 0 ? N : ldy; g.a_sk = 1;
  g.B = W; g.b_sk = K; g.b_sn = 1;
  g.C = dx; g.ldc = ldc < 0 ? K : ldc;
  g.epi = epi; g.aux = aux; g.ldaux = ldaux; g.resid = resid; g.ldr = K;
  gemm<TA, TB, TC>(g, st);
}
// dW[N,K] = dy[R,N]^T x[R,K]   (split-K over R, deterministic) ; db[N] = colsum(dy)
template <typename TA, typename TB>
static void linear_bwd_w(const TA* dy, const TB* x, float* dW, float* db, int64_t R, int N, int K,
                         float* partial, cudaStream_t st, int64_t ldy = -1, int64_t ldx = -1, ReduceList* defer = nullptr) {
  if (ldy < 0) ldy = N;
  if (ldx < 0) ldx = K;
  if (defer) partial = defer->alloc((size_t)pick_splitk(R) * N * K);
  GemmArgs g;
  g.defer = defer;
  g.M = N; g.N = K; g.K = (int)R;
  g.A = dy; g.a_sm = 1; g.a_sk = ldy;
  g.B = x; g.b_sk = ldx; g.b_sn = 1;
  g.C = dW; g.ldc = K;
  g.splitk = pick_splitk(R); g.partial = partial;
  gemm<TA, TB, float>(g, st);
  if (db) {
    // long, narrow dy (the CNN critic's convolutions): whole-row warps, many row ranges
    const bool tall = R >= 32768 && N % 2 == 0 && N <= 512 && 256 % (N / 2) == 0 && ldy % 2 == 0 && ((((uintptr_t)dy) & 7) == 0);
    const int S = tall ? (int)std::min<int64_t>(R / 256, 148 * 4)
                       : (int)std::min<int64_t>(std::max<int64_t>(R / 64, 1), 128);   // row ranges per column block
    const int64_t rpb = cdiv(R, S);
    dim3 grid((unsigned)cdiv(N, 256), (unsigned)S);
    if (defer) partial = defer->alloc((size_t)S * N);      // the GEMM's partials may still be queued
    if (tall) launch_k(colsum_tall_kernel<TA>, (unsigned)S, 256, 0, st, dy, ldy, partial, R, N, rpb);
    else launch_k(colsum_partial_kernel<TA>, grid, 128, 0, st, dy, ldy, partial, R, N, rpb);
    DG_LAUNCH_CHECK();
    if (defer && N % 4 == 0) {
      defer->add(partial, db, S, N, N);
    } else {
      launch_k(reduce_partials_kernel, reduce_grid(N), 256, 0, st, partial, db, S, N);
      DG_LAUNCH_CHECK();
    }
  }
}

// db[N] = colsum(dy[R, N]) through the deferred reduction (fixed summation order)
template <typename TA>
static void bias_grad_colsum(const TA* dy, int64_t ldy, float* db, int64_t R, int N, ReduceList& rl, cudaStream_t st) {
  // narrow dy (N = 64: the patch-embedding bias): whole-row warps instead of 32 active threads per block
  const bool tall = R >= 4096 && N % 2 == 0 && N <= 512 && 256 % (N / 2) == 0 && ldy % 2 == 0 && ((((uintptr_t)dy) & 7) == 0);
  const int S = tall ? (int)std::min<int64_t>(R / 128, 148 * 2)
                     : (int)std::min<int64_t>(std::max<int64_t>(R / 64, 1), 128);   // row ranges per column block
  float* partial = rl.alloc((size_t)S * N);
  if (tall) launch_k(colsum_tall_kernel<TA>, (unsigned)S, 256, 0, st, dy, ldy, partial, R, N, cdiv(R, S));
  else launch_k(colsum_partial_kernel<TA>, dim3((unsigned)cdiv(N, 256), (unsigned)S), 128, 0, st, dy, ldy, partial, R, N, cdiv(R, S));
  DG_LAUNCH_CHECK();
  rl.add(partial, db, S, N, N);
}

static unsigned grid1d(int64_t n, int bs = 256) {
  int64_t g = cdiv(n, bs);
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace dgvit
#include "qnet.cuh"
namespace dgvit {

// ------------------------------------------------------------------ trunk context
template <typename A>
struct LayerBuf {
  float *Xa, *Xm, *mean1, *rstd1, *mean2, *rstd2;
  A *Xn1, *QKV, *O, *Xn2, *Hpre, *Hact;
  float* lse = nullptr;     // [B, H, N] softmax statistics of the long-sequence attention kernels (N > 128 only)
};
template <typename A>
struct TrunkCtx {
  A* Pm; float* Xp; float* tok; float* Xout; float* z;
  const A* Pm_ext = nullptr;   // patch matrix computed elsewhere for the same frames (the update patchifies s and s' once)
  LayerBuf<A> L[DGVIT_MAX_DEPTH];
  // backward scratch (only when saved)
  float *dX, *dXn, *dtok, *dz, *dg_rows, *partial;
  A *dH, *dO, *dQKV, *dXp;
  float* attn_delta = nullptr;   // [B, H, N] rowsum(dO o O) of the long-sequence attention backward (N > 128 only)
  bf16* dXh;  // bf16 copy of dX (operand of the tensor-core GEMMs); null in the fp32 path
  float* dXc; bf16* dXch;   // last block: compact [B, D] residual gradient of the token-0 rows
  bf16 *dXh2, *dXch2;       // second bf16 copies (gradient after LayerNorm-2): lets the weight-gradient stream keep reading
                            // the block's incoming gradient while the dX chain moves on
  size_t partial_floats, misc_floats;
  bool save;   // activations kept for a backward pass
  // residual-stream gradient in the operand dtype
  const A* dx_op() const {
    if constexpr (std::is_same<A, float>::value) return dX; else return dXh;
  }
  const A* dxc_op() const {
    if constexpr (std::is_same<A, float>::value) return dXc; else return dXch;
  }
};

template <typename A>
static void carve_trunk(Carver& cv, const Dims& d, bool save, TrunkCtx<A>& c) {
  c.save = save;
  c.Pm = cv.take<A>((int64_t)d.B * d.P * d.pd);
  c.Xp = cv.take<float>((int64_t)d.B * d.P * d.D);
  c.tok = cv.take<float>((int64_t)d.B * d.D);
  c.z = cv.take<float>((int64_t)d.B * d.D);
  const int nl = save ? d.L : 1;
  for (int l = 0; l < nl; ++l) {
    LayerBuf<A>& b = c.L[l];
    b.Xa = cv.take<float>(d.T * d.D);
    b.Xm = cv.take<float>(d.T * d.D);
    b.mean1 = cv.take<float>(d.T); b.rstd1 = cv.take<float>(d.T);
    b.mean2 = cv.take<float>(d.T); b.rstd2 = cv.take<float>(d.T);
    b.Xn1 = cv.take<A>(d.T * d.D);
    b.QKV = cv.take<A>(d.T * 3 * d.inner);
    b.O = cv.take<A>(d.T * d.inner);
    b.lse = (d.N > 128 && !std::is_same<A, float>::value) ? cv.take<float>(d.T * d.H) : nullptr;
    b.Xn2 = cv.take<A>(d.T * d.D);
    b.Hpre = cv.take<A>(d.T * d.M);
    b.Hact = cv.take<A>(d.T * d.M);
  }
  if (save) {
    c.Xout = cv.take<float>(d.T * d.D);
  } else {
    for (int l = 1; l < d.L; ++l) c.L[l] = c.L[0];
    c.Xout = c.L[0].Xa;  // ping-pong: Xa -> Xm -> Xa
  }
  c.dX = nullptr; c.dXn = nullptr; c.dtok = nullptr; c.dz = nullptr; c.dg_rows = nullptr; c.partial = nullptr;
  c.dH = nullptr; c.dO = nullptr; c.dQKV = nullptr; c.dXp = nullptr; c.dXh = nullptr; c.partial_floats = 0; c.misc_floats = 0;
  c.dXc = nullptr; c.dXch = nullptr; c.dXh2 = nullptr; c.dXch2 = nullptr; c.attn_delta = nullptr;
  if (save) {
    c.dX = cv.take<float>(d.T * d.D);
    c.dXn = cv.take<float>(d.T * d.D);
    c.dtok = cv.take<float>((int64_t)d.B * d.D);
    c.dz = cv.take<float>((int64_t)d.B * d.D);
    c.dg_rows = cv.take<float>((int64_t)d.B * d.D);
    c.dH = cv.take<A>(d.T * d.M);
    c.dO = cv.take<A>(d.T * d.inner);
    c.dQKV = cv.take<A>(d.T * 3 * d.inner);
    c.attn_delta = (d.N > 128 && !std::is_same<A, float>::value) ? cv.take<float>(d.T * d.H) : nullptr;
    c.dXp = cv.take<A>((int64_t)d.B * d.P * d.D);
    c.dXh = std::is_same<A, float>::value ? nullptr : cv.take<bf16>(d.T * d.D);
    c.dXc = cv.take<float>((int64_t)d.B * d.D);
    c.dXch = std::is_same<A, float>::value ? nullptr : cv.take<bf16>((int64_t)d.B * d.D);
    c.dXh2 = std::is_same<A, float>::value ? nullptr : cv.take<bf16>(d.T * d.D);
    c.dXch2 = std::is_same<A, float>::value ? nullptr : cv.take<bf16>((int64_t)d.B * d.D);
    int64_t mx = (int64_t)d.D * d.M;
    mx = std::max<int64_t>(mx, (int64_t)3 * d.inner * d.D);
    mx = std::max<int64_t>(mx, (int64_t)d.D * d.pd);
    mx = std::max<int64_t>(mx, (int64_t)128 * 128);
    mx = std::max<int64_t>(mx, (int64_t)3 * d.D * 148 * 4 / 32 + 3 * d.D);
    // one block's worth of queued reductions (split-K dW of qkv / out / both MLP matrices + LayerNorm partials)
    const int64_t per_block = (int64_t)32 * (4 * d.inner * d.D + 2 * d.D * d.M) + (int64_t)8 * d.D * 148 * 4 + 4 * d.M + 4096;
    // + the reductions outside the blocks (rms gain, pos embedding, patch dW split-K + bias)
    c.misc_floats = (size_t)(64 * d.D + (int64_t)32 * d.N * d.D + (int64_t)80 * d.D * d.pd + 320 * d.D + 1024);
    c.partial_floats = (size_t)std::max<int64_t>(mx * 32, per_block) + c.misc_floats;
    c.partial = cv.take<float>(c.partial_floats);
  }
}

static DropDev make_drop(const dgvit_drop& d, const Dims& dm, int64_t sample_offset) {
  DropDev r;
  r.mode = d.mode;
  r.p = d.p;
  r.scale = 1.0f / (float)(1.0 - (double)d.p);
  r.mask = d.keep_mask;
  r.rng = d.rng_state;
  r.stream_id = d.stream_id;
  r.elem_offset = sample_offset * dm.N * dm.D;
  if (d.mode == DGVIT_DROP_MASK) DG_REQUIRE(d.keep_mask != nullptr, "DROP_MASK without keep_mask");
  if (d.mode == DGVIT_DROP_RNG) DG_REQUIRE(d.rng_state != nullptr, "DROP_RNG without rng_state");
  return r;
}

template <typename A>
static void launch_ln_fwd(const float* X, const float* g, const float* b, A* Y, float* mean, float* rstd,
                          int64_t T, int D, cudaStream_t st) {
  const int wpb = 8;
  const unsigned grid = (unsigned)cdiv(T, wpb);
  switch (D / 32) {
#define LNF(V) case V: launch_k(layernorm_fwd_kernel<A, V>, grid, wpb * 32, 0, st, X, g, b, Y, mean, rstd, T); break;
    LNF(1) LNF(2) LNF(3) LNF(4) LNF(5) LNF(6) LNF(7) LNF(8)
#undef LNF
    default: fail(DGVIT_ERR_ARG, "unsupported dim %d", D);
  }
  DG_LAUNCH_CHECK();
}
// dxsum (optional): receives colsum over rows of the updated dX_io (a bias gradient, see layernorm_bwd_kernel)
static int g_lnb_wpb = 16, g_lnb_bps = 2;   // set_option "ln_bwd_warps", "ln_bwd_blocks_per_sm"
static void launch_ln_bwd(const float* dY, const float* X, const float* mean, const float* rstd,
                          const float* gamma, float* dX_io, bf16* dX_lp, float* dgamma, float* dbeta, float* dxsum,
                          float* partial, int64_t T, int D, cudaStream_t st, ReduceList* defer = nullptr,
                          const float* dX_row0 = nullptr, int Ntok = 1) {
  // 16 warps per block and at most two blocks per SM: 296 partial rows for the reduction instead of 592
  const int wpb = g_lnb_wpb;
  int nblocks = (int)std::min<int64_t>(cdiv(T, wpb), 148 * g_lnb_bps);
  if (skip_mask() & SKIP_LN_BWD) return;
  // algorithmic bytes per row: dY, X, dX in / out (fp32) + the bf16 copy of dX + mean / rstd
  ProfScope ps(PROF_LN_BWD, 0.0, (double)T * (D * (16.0 + (dX_lp ? 2.0 : 0.0)) + 8.0), st);
  if (defer && D % 4 == 0) partial = defer->alloc((size_t)nblocks * 3 * D); else defer = nullptr;
  const size_t smem = (size_t)wpb * 3 * D * sizeof(float);
  switch (D / 32) {
#define LNB(V) case V: launch_k(layernorm_bwd_kernel<V>, nblocks, wpb * 32, smem, st, dY, X, mean, rstd, gamma, dX_io, dX_lp, partial, T, dX_row0, Ntok); break;
    LNB(1) LNB(2) LNB(3) LNB(4) LNB(5) LNB(6) LNB(7) LNB(8)
#undef LNB
    default: fail(DGVIT_ERR_ARG, "unsupported dim %d", D);
  }
  DG_LAUNCH_CHECK();
  if (defer) {
    defer->add(partial, dgamma, nblocks, D, 3 * D);
    defer->add(partial + D, dbeta, nblocks, D, 3 * D);
    if (dxsum) defer->add(partial + 2 * D, dxsum, nblocks, D, 3 * D);
    return;
  }
  launch_k(ln_param_reduce_kernel, (unsigned)cdiv(3 * D, 32), 1024, 0, st, partial, dgamma, dbeta, dxsum, nblocks, D);
  DG_LAUNCH_CHECK();
}

template <typename A>
static void launch_attention_fwd(const A* QKV, A* O, const Dims& d, cudaStream_t st, float* lse = nullptr) {
  ProfScope ps(PROF_ATTENTION, 4.0 * d.B * d.H * (double)d.N * d.N * d.dh, 0.0, st);
#ifdef DGVIT_WITH_TC
  if constexpr (std::is_same<A, bf16>::value) {
    if (attn::eligible(d.N, d.dh, QKV, 3 * d.inner)) return attn::fwd(QKV, O, d.B, d.N, d.H, st);
    // more than one 128-row tile (257 tokens at 2x resolution): key-chunked tcgen05 kernels (attn_long_tc.cuh)
    if (lse && attnl::eligible(d.N, d.dh, QKV, 3 * d.inner)) return attnl::fwd(QKV, O, lse, d.B, d.N, d.H, st);
  }
#endif
  const int threads = 256, nw = threads / 32;
  const size_t smem = ((size_t)d.N * (d.dh + 1) + (size_t)d.N * d.dh + (size_t)nw * d.N + (size_t)nw * d.dh) * sizeof(float);
  DG_REQUIRE(smem <= 227 * 1024, "attention_fwd: N=%d dh=%d needs %zu B smem", d.N, d.dh, smem);
  static DevOnce attr_done;  // attribute is per-function; set once per dtype instantiation
  if (attr_done.first()) {
    DG_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  launch_k(attention_fwd_kernel<A>, d.B * d.H, threads, smem, st, QKV, O, d.N, d.H, d.dh, 1.0f / sqrtf((float)d.dh));
  DG_LAUNCH_CHECK();
}
template <typename A>
static void launch_attention_bwd(const A* QKV, const A* O, const A* dO, A* dQKV, const Dims& d, cudaStream_t st,
                                 float* lse = nullptr, float* delta = nullptr) {
  ProfScope ps(PROF_ATTENTION, 10.0 * d.B * d.H * (double)d.N * d.N * d.dh, 0.0, st);
#ifdef DGVIT_WITH_TC
  if constexpr (std::is_same<A, bf16>::value) {
    if (attn::eligible(d.N, d.dh, QKV, 3 * d.inner)) return attn::bwd(QKV, O, dO, dQKV, d.B, d.N, d.H, st);
    if (lse && delta && attnl::eligible(d.N, d.dh, QKV, 3 * d.inner) && ((((uintptr_t)dO) | ((uintptr_t)dQKV) | ((uintptr_t)O)) & 15) == 0)
      return attnl::bwd(QKV, O, dO, dQKV, lse, delta, d.B, d.N, d.H, st);
  }
#endif
  const int threads = 256, nw = threads / 32;
  const size_t small = ((size_t)2 * d.N * (d.dh + 1) + 2 * (size_t)d.N + 2 * (size_t)nw * d.N + 2 * (size_t)nw * d.dh) * sizeof(float);
  const size_t full = small + (size_t)2 * d.N * (d.dh + 1) * sizeof(float);
  const bool qdo_smem = full <= 227 * 1024;
  const size_t smem = qdo_smem ? full : small;
  DG_REQUIRE(smem <= 227 * 1024, "attention_bwd: N=%d dh=%d needs %zu B smem (unsupported in this build)", d.N, d.dh, smem);
  static DevOnce attr_done;
  if (attr_done.first()) {
    DG_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<A, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DG_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<A, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const float scale = 1.0f / sqrtf((float)d.dh);
  if (qdo_smem) launch_k(attention_bwd_kernel<A, true>, d.B * d.H, threads, smem, st, QKV, O, dO, dQKV, d.N, d.H, d.dh, scale);
  else launch_k(attention_bwd_kernel<A, false>, d.B * d.H, threads, smem, st, QKV, O, dO, dQKV, d.N, d.H, d.dh, scale);
  DG_LAUNCH_CHECK();
}

// last block: one query row per (sample, head) (kernels.cuh, attention_row0_*); false = shape not supported
template <typename A>
static bool launch_attention_row0(const A* QKV, A* O, const A* dO, A* dQKV, const Dims& d, cudaStream_t st) {
  if (d.dh != R0_DH || d.N > R0_MAXN || !(g_row0_mode & (dO ? 2 : 1))) return false;
  if ((((uintptr_t)QKV) & 15) || (dO && ((((uintptr_t)dO) | ((uintptr_t)dQKV)) & 15))) return false;
  const int items = d.B * d.H;
  // a handful of (sample, head) items (batch-1 act: 4) is one under-filled block walking 65 keys serially: 15 us against 6 us
  // for the tensor-core kernel on the whole (tiny) tile
  if (!dO && items < 32 && std::is_same<A, bf16>::value) return false;
  const float scale = 1.0f / sqrtf((float)d.dh);
  if (!dO) launch_k(attention_row0_fwd_kernel<A>, (unsigned)cdiv(items, 4), 128, 0, st, QKV, O, items, d.N, d.H, scale);
  else launch_k(attention_row0_bwd_kernel<A>, (unsigned)cdiv(items, 4), 128, 0, st, QKV, dO, dQKV, items, d.N, d.H, scale);
  DG_LAUNCH_CHECK();
  return true;
}

// the fused tcgen05 MLP (mlp_tc.cuh) replaces fc1 -> GELU -> fc2 (+residual) when eligible;
// forward and backward must take the same decision (the fused forward saves only the pre-activation)
template <typename A>
static bool mlp_fused(const Dims& d, int64_t R, const void* xn2, const void* w1, const void* w2, const float* resid,
                      const float* out) {
#ifdef DGVIT_WITH_TC
  if constexpr (std::is_same<A, bf16>::value) return mlp::eligible(d.D, d.M, R, xn2, w1, w2, resid, d.D, out, d.D);
#endif
  return false;
}

// ------------------------------------------------------------------ trunk forward
// Rearrange 'b (h p1) (w p2) -> b (h w) (p1 p2)' (vn/GoalFormer.py:138) + cast to the operand dtype: the A matrix of the patch GEMM
template <typename A>
static void launch_patchify(const float* img, A* Pm, int B, const Dims& d, const dgvit_cfg& cfg, cudaStream_t st) {
  const int64_t total4 = (int64_t)B * d.P * d.pd / 4;
  DG_REQUIRE(cfg.patch_w % 4 == 0 && (((uintptr_t)img) & 15) == 0, "patchify: patch_w %% 4 and 16-byte aligned frames required");
  ProfScope ps(PROF_PATCH, 0.0, (double)total4 * 4 * (4.0 + sizeof(A)), st);
  launch_k(patchify_kernel<A>, grid1d(total4), 256, 0, st, img, Pm, total4, cfg.img_h, cfg.img_w, cfg.patch_h, cfg.patch_w);
  DG_LAUNCH_CHECK();
}
// GoT.forward (vn/GoalFormer.py:156-171) given the goal token tok[B,D]; writes c.z.
template <typename A>
static void trunk_forward(const dgvit_net& net, const dgvit_layout& L, const Dims& d, const float* img,
                          const DropDev& drop, TrunkCtx<A>& c, cudaStream_t st, const GoalTok& gt, bool fuse_rms = false) {
  const dgvit_cfg& cfg = net.cfg;
  const float* P = net.params;
  // K1: patch embedding
  bool embedded = false;
#ifdef DGVIT_WITH_TC
  if constexpr (std::is_same<A, bf16>::value) {
    // im2col-free: frame -> swizzled smem A tiles -> tcgen05 -> bias + goal token + pos + dropout + LayerNorm-1, one launch
    if (!(skip_mask() & SKIP_EMBED) && patch::eligible(cfg, img, WSel<A>::w(net, L.patch_w), c.L[0].Xn1)) {
      const dgvit_block_layout& b0 = L.block[0];
      LayerBuf<A>& B0 = c.L[0];
      patch::PatchArgs pa;
      pa.img = img; pa.n_tok = (int64_t)d.B * d.P; pa.P = d.P; pa.N = d.N;
      pa.bias = P + L.patch_b; pa.pos = P + L.pos; pa.gt = gt; pa.tok = c.tok; pa.drop = drop;
      pa.X0 = B0.Xa; pa.Y = (bf16*)B0.Xn1; pa.gamma = P + b0.ln1_w; pa.beta = P + b0.ln1_b; pa.mean = B0.mean1; pa.rstd = B0.rstd1;
      ProfScope ps(PROF_EMBED, 2.0 * d.B * d.P * d.pd * d.D, (double)d.B * (cfg.img_h * cfg.img_w * 4.0 + d.N * (d.D * 6.0 + 8.0)), st);
      ProfScope ps2(PROF_GEMM_ALL, 2.0 * d.B * d.P * d.pd * d.D, 0.0, st);
      patch::fwd(cfg, pa, (const bf16*)WSel<A>::w(net, L.patch_w), st);
      embedded = true;
    }
  }
#endif
  if (!embedded && !(skip_mask() & SKIP_EMBED)) {
    if (!c.Pm_ext) launch_patchify<A>(img, c.Pm, d.B, d, cfg, st);
    linear_fwd<A, A, float>(c.Pm_ext ? c.Pm_ext : c.Pm, WSel<A>::w(net, L.patch_w), c.Xp, (int64_t)d.B * d.P, d.D, d.pd, EPI_BIAS,
                            P + L.patch_b, st);
    // goal token (fc_embed) + cat + pos + dropout + the first block's LayerNorm-1 in one warp-per-row kernel
    {
      const dgvit_block_layout& b0 = L.block[0];
      LayerBuf<A>& B0 = c.L[0];
      const int wpb = 8;
      const unsigned grid = (unsigned)cdiv(d.T, wpb);
      // algorithmic bytes per token row: patch-embedding row in, residual stream out (fp32), LN output, mean / rstd
      ProfScope ps(PROF_EMBED, 0.0, (double)d.T * (d.D * (8.0 + sizeof(A)) + 8.0), st);
      switch (d.D / 32) {
#define EMB(V) case V: launch_k(embed_ln_kernel<A, V>, grid, wpb * 32, 0, st, gt, c.tok, (const float*)c.Xp, P + L.pos, B0.Xa, drop, \
                                P + b0.ln1_w, P + b0.ln1_b, B0.Xn1, B0.mean1, B0.rstd1, d.T, d.N); break;
        EMB(1) EMB(2) EMB(3) EMB(4) EMB(5) EMB(6) EMB(7) EMB(8)
#undef EMB
        default: fail(DGVIT_ERR_ARG, "unsupported dim %d", d.D);
      }
      DG_LAUNCH_CHECK();
    }
  }
  bool ln1_done = !(skip_mask() & SKIP_EMBED);      // block 0's LayerNorm-1 ran inside embed_ln_kernel
  for (int l = 0; l < d.L; ++l) {
    const dgvit_block_layout& b = L.block[l];
    LayerBuf<A>& B_ = c.L[l];
    float* Xnext = (l + 1 < d.L) ? c.L[l + 1].Xa : c.Xout;
    // attention block: x = attn(LN(x)) + x   (LN already applied by the previous block's fused MLP kernel when ln1_done)
    if (!ln1_done) launch_ln_fwd<A>(B_.Xa, P + b.ln1_w, P + b.ln1_b, B_.Xn1, B_.mean1, B_.rstd1, d.T, d.D, st);
    ln1_done = false;
    linear_fwd<A, A, A>(B_.Xn1, WSel<A>::w(net, b.qkv_w), B_.QKV, d.T, 3 * d.inner, d.D, EPI_NONE, nullptr, st);
    // (last block: only the token-0 row of the attention output is ever read)
    if (!(l == d.L - 1 && launch_attention_row0<A>(B_.QKV, B_.O, nullptr, nullptr, d, st))) launch_attention_fwd<A>(B_.QKV, B_.O, d, st, B_.lse);
    // Only token 0 of the last block's output is consumed (x[:, 0], vn/GoalFormer.py:167): there the
    // out-projection, LayerNorm, MLP and residuals run on the B token-0 rows only (compact [B, D]
    // buffers).  Exact: the pruned rows never reach z, so outputs and every gradient are unchanged.
    const bool last = (l == d.L - 1);
    const int64_t R = last ? d.B : d.T;
    const int64_t ostride = last ? (int64_t)d.N * d.inner : d.inner;
    const int64_t xstride = last ? (int64_t)d.N * d.D : d.D;
    // out-projection + bias + residual + LayerNorm-2 (x = attn(LN x) + x ; ff input = LN(x), vn/GoalFormer.py:103-104):
    //   (a) bf16 path, inner = 256: the prologue of the fused MLP kernel below (no launch of its own);
    //   (b) else the tensor-core GEMM with LayerNorm-2 in its epilogue;  (c) else GEMM + LayerNorm kernels.
    bool ln2_done = false, front_done = false;
#ifdef DGVIT_WITH_TC
    mlp::FrontFuse fr;
    if constexpr (std::is_same<A, bf16>::value) {
      fr.o = (const bf16*)B_.O; fr.ldo = ostride; fr.Wo = (const bf16*)WSel<A>::w(net, b.out_w); fr.ob = P + b.out_b;
      fr.xa = B_.Xa; fr.ldxa = xstride; fr.xm = B_.Xm; fr.gamma = P + b.ln2_w; fr.beta = P + b.ln2_b;
      fr.xn2 = (bf16*)B_.Xn2; fr.mean = B_.mean2; fr.rstd = B_.rstd2;
      front_done = mlp::front_eligible(d.inner, fr) &&
                   mlp_fused<A>(d, R, B_.Xn2, WSel<A>::w(net, b.fc1_w), WSel<A>::w(net, b.fc2_w), B_.Xm, Xnext);
      if (!front_done) {
        GemmArgs g;
        g.M = (int)R; g.N = d.D; g.K = d.inner;
        g.A = B_.O; g.a_sm = ostride; g.a_sk = 1;
        g.B = WSel<A>::w(net, b.out_w); g.b_sk = 1; g.b_sn = d.inner;
        g.C = B_.Xm; g.ldc = d.D;
        g.epi = EPI_BIAS_RESID; g.bias = P + b.out_b; g.resid = B_.Xa; g.ldr = xstride;
        g.ln_gamma = P + b.ln2_w; g.ln_beta = P + b.ln2_b; g.ln_out = B_.Xn2; g.ln_mean = B_.mean2; g.ln_rstd = B_.rstd2;
        ProfScope ps_all(PROF_GEMM_ALL, 2.0 * g.M * g.N * g.K, 0.0, st);
        ln2_done = gemm_tc_try<A, A, float>(g, st);
      }
    }
#endif
    if (!ln2_done && !front_done) {
      linear_fwd<A, A, float>(B_.O, WSel<A>::w(net, b.out_w), B_.Xm, R, d.D, d.inner, EPI_BIAS_RESID, P + b.out_b, st,
                              B_.Xa, nullptr, -1, ostride, xstride);
      // MLP block: x = ff(LN(x)) + x
      launch_ln_fwd<A>(B_.Xm, P + b.ln2_w, P + b.ln2_b, B_.Xn2, B_.mean2, B_.rstd2, R, d.D, st);
    }
    TagScope mlp_tag(PROF_GEMM_MLP);
    if (mlp_fused<A>(d, R, B_.Xn2, WSel<A>::w(net, b.fc1_w), WSel<A>::w(net, b.fc2_w), B_.Xm, Xnext)) {
#ifdef DGVIT_WITH_TC
      if constexpr (std::is_same<A, bf16>::value) {
        ProfScope ps(PROF_GEMM_MLP, 4.0 * R * d.D * d.M, 0.0, st);
        ProfScope ps2(PROF_MLP_FUSED, 4.0 * R * d.D * d.M + (front_done ? 2.0 * R * d.D * d.inner : 0.0), 0.0, st);
        mlp::LnFuse ln;
        if (!last) {   // the next block's LayerNorm-1 rides in this kernel's output stage
          const dgvit_block_layout& nb = L.block[l + 1];
          LayerBuf<A>& N_ = c.L[l + 1];
          ln.gamma = P + nb.ln1_w; ln.beta = P + nb.ln1_b; ln.out = (bf16*)N_.Xn1; ln.mean = N_.mean1; ln.rstd = N_.rstd1;
          ln1_done = true;
        }
        ProfScope ps3(PROF_GEMM_ALL, front_done ? 2.0 * R * d.D * d.inner : 0.0, 0.0, st);
        mlp::fwd(B_.Xn2, WSel<A>::w(net, b.fc1_w), P + b.fc1_b, WSel<A>::w(net, b.fc2_w), P + b.fc2_b, B_.Xm, d.D, Xnext,
                 d.D, R, d.M, st, ln, front_done ? fr : mlp::FrontFuse(), w_f16(net, L, b.fc2_w));
      }
#endif
    } else {
      // (forward passes without a backward do not keep the pre-activation: half of the hidden-activation traffic)
      linear_fwd<A, A, A>(B_.Xn2, WSel<A>::w(net, b.fc1_w), B_.Hpre, R, d.M, d.D, EPI_BIAS_GELU2, P + b.fc1_b, st,
                          nullptr, B_.Hact, -1, -1, -1, c.save ? 0 : 1);
      linear_fwd<A, A, float>(B_.Hact, WSel<A>::w(net, b.fc2_w), Xnext, R, d.D, d.M, EPI_BIAS_RESID, P + b.fc2_b, st,
                              B_.Xm);
    }
  }
  // c.Xout is compact [B, D] (token-0 rows of the last block); cls pooling + RMSNorm -> c.z, unless the caller's head
  // kernel does it while staging its input (fuse_rms)
  if (!fuse_rms) {
    launch_k(pool_rmsnorm_fwd_kernel, (unsigned)cdiv(d.B, 8), 256, 0, st, c.Xout, P + L.rms_g, c.z, d.B, 1, d.D,
                                                                     sqrtf((float)d.D));
    DG_LAUNCH_CHECK();
  }
}

// ------------------------------------------------------------------ trunk backward
// Only the dX launches of a block are on the dependency chain to the next block; the weight-gradient launches (MLP dW,
// dW_out, dW_qkv) and the block's multi-reduce feed nothing downstream.  In the bf16 path they are issued on a second
// stream (event edges, capturable) so that the latency-bound links of the chain (LayerNorm backward, the short GEMMs)
// run beside them instead of in series.  set_option "bwd_side" = 0 puts everything back on one stream.
struct SideState { cudaStream_t s; cudaEvent_t now, dw, mr; };
static bool g_side_enabled = true;
static bool g_fork_enabled = true;
static int g_target_fork = 1;       // set_option "target_fork": critic_target's trunk on its own stream beside policy.sample(s')
static int g_actor_s_when = 1;      // set_option "actor_s_when": 0 = policy.sample(s) forward starts at the fork, 1 = after the
                                    // policy.sample(s') forward, 2 = after the target critic forward (beside the critic backward)
static SideState& side_state() {      // one set of streams / events per device
  static SideState per_dev[64];
  static bool inited[64] = {};
  const int dev = current_device();
  SideState& f = per_dev[dev];
  bool& init = inited[dev];
  if (!init) {
    const char* pe = getenv("DGVIT_SIDE_PRIO");
    DG_CUDA(cudaStreamCreateWithPriority(&f.s, cudaStreamNonBlocking, pe ? atoi(pe) : 0));
    DG_CUDA(cudaEventCreateWithFlags(&f.now, cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&f.dw, cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&f.mr, cudaEventDisableTiming));
    init = true;
  }
  return f;
}

template <typename A>
static bool side_on(const TrunkCtx<A>& c) {
  return g_side_enabled && g_fork_enabled && std::is_same<A, bf16>::value && c.dXh2 && !prof().on;
}
// stream for work that only the end of trunk_backward (its join) waits for; sees everything issued on st so far
template <typename A>
static cudaStream_t side_fork(const TrunkCtx<A>& c, cudaStream_t st) {
  if (!side_on(c)) return st;
  SideState& sd = side_state();
  DG_CUDA(cudaEventRecord(sd.now, st));
  DG_CUDA(cudaStreamWaitEvent(sd.s, sd.now, 0));
  return sd.s;
}

// generic fork / join of the side stream for bookkeeping kernels that nothing on the caller's stream waits for
static cudaStream_t lane_fork(cudaStream_t st) {
  if (!(g_side_enabled && g_fork_enabled && !prof().on)) return st;
  SideState& sd = side_state();
  DG_CUDA(cudaEventRecord(sd.now, st));
  DG_CUDA(cudaStreamWaitEvent(sd.s, sd.now, 0));
  return sd.s;
}
static void lane_join(cudaStream_t lane, cudaStream_t st) {
  if (lane == st) return;
  SideState& sd = side_state();
  DG_CUDA(cudaEventRecord(sd.mr, lane));
  DG_CUDA(cudaStreamWaitEvent(st, sd.mr, 0));
}

// in: c.dz [B,D]; out: grads of every trunk parameter, c.dtok [B,D]
template <typename A>
static void trunk_backward(const dgvit_net& net, const dgvit_layout& L, const Dims& d, const DropDev& drop,
                           TrunkCtx<A>& c, int relu_tok, cudaStream_t st, const float* img) {
  const float* P = net.params;
  float* G = net.grads;
  const bool side = side_on(c);
  SideState* sd = side ? &side_state() : nullptr;
  cudaStream_t ss = side ? sd->s : st;
  auto side_after_main = [&]() {      // work issued on the side stream from here on sees everything main has issued so far
    if (side) { DG_CUDA(cudaEventRecord(sd->now, st)); DG_CUDA(cudaStreamWaitEvent(ss, sd->now, 0)); }
  };
  bool have_mr = false;
  bool have_dw = false;      // the MLP weight-gradient launch of the current block has been recorded on the side stream
  launch_k(pool_rmsnorm_bwd_kernel, d.B, 128, 0, st, c.Xout, P + L.rms_g, c.dz, c.dXc, c.dXch, c.dg_rows, d.B, 1, d.D,
                                                sqrtf((float)d.D));
  DG_LAUNCH_CHECK();
  // reductions outside the blocks (RMSNorm gain, position embedding, patch embedding) are queued in the tail of the
  // partial-sum workspace and done by one launch at the end
  ReduceList rl_misc(c.partial + c.partial_floats - c.misc_floats, c.misc_floats);
  const size_t block_floats = c.partial_floats - c.misc_floats;
  {  // dg = sum_b dg_rows
    const int S = std::max(1, std::min(64, d.B / 4));
    dim3 grid((unsigned)cdiv(d.D, 256), (unsigned)S);
    float* part = rl_misc.alloc((size_t)S * d.D);
    launch_k(colsum_partial_kernel<float>, grid, 128, 0, st, c.dg_rows, d.D, part, d.B, d.D, cdiv(d.B, S));
    DG_LAUNCH_CHECK();
    if (d.D % 4 == 0) rl_misc.add(part, G + L.rms_g, S, d.D, d.D);
    else {
      launch_k(reduce_partials_kernel, 1, 256, 0, st, part, G + L.rms_g, S, d.D);
      DG_LAUNCH_CHECK();
    }
  }
  for (int l = d.L - 1; l >= 0; --l) {
    const dgvit_block_layout& b = L.block[l];
    LayerBuf<A>& B_ = c.L[l];
    // the last block carries only the token-0 rows (compact [B, D]) down to its attention output
    const bool last = (l == d.L - 1);
    const int64_t R = last ? d.B : d.T;
    float* dXr = last ? c.dXc : c.dX;                  // fp32 residual-stream gradient of these rows
    const A* dxop = last ? c.dxc_op() : c.dx_op();     // operand copy of dL/dX_out
    // operand copy of dL/dX_m (after the LayerNorm-2 backward): a second buffer when the dW stream is on
    bf16* dXr_lp = side ? (last ? c.dXch2 : c.dXh2) : (last ? c.dXch : c.dXh);
    const A* dxop_m = side ? (const A*)dXr_lp : dxop;
    side_after_main();
    // every partial-sum reduction of this block is queued here and done by one launch at the end of the block
    ReduceList rl(c.partial, block_floats);
    // ---- MLP block.  dXr = dL/dX_out
    {
    TagScope mlp_tag(PROF_GEMM_MLP);
    float* Xn_out = (l + 1 < d.L) ? c.L[l + 1].Xa : c.Xout;
    bool fused = mlp_fused<A>(d, R, B_.Xn2, WSel<A>::w(net, b.fc1_w), WSel<A>::w(net, b.fc2_w), B_.Xm, Xn_out);
#ifdef DGVIT_WITH_TC
    if constexpr (std::is_same<A, bf16>::value) {
      if (fused) {
        // the fused forward saved nothing: both backward launches recompute the pre-activation on chip
        DG_REQUIRE(mlp::bwd_partial_floats(R, d.M) <= block_floats, "mlp::bwd partial buffer too small");
        ProfScope ps(PROF_GEMM_MLP, 8.0 * R * d.D * d.M, 0.0, st);
        // net.3.bias gradient = colsum(dL/dX_out): below the top block it falls out of the next block's LayerNorm-1 backward
        mlp::bwd(B_.Xn2, dxop, WSel<A>::w(net, b.fc1_w), P + b.fc1_b, WSel<A>::w(net, b.fc2_w), c.dXn, G + b.fc1_w,
                 G + b.fc1_b, G + b.fc2_w, last ? G + b.fc2_b : nullptr, c.partial, R, d.M, st, &rl, ss);
        if (side) { DG_CUDA(cudaEventRecord(sd->dw, ss)); have_dw = true; }
      }
    }
#else
    fused = false;
#endif
    if (!fused) {
      // These weight-gradient GEMMs put their split-K / column-sum partials at the START of c.partial, from the main stream,
      // right now: the previous block's multi-reduce (side stream) reads its queued partial sums from the same place and must
      // be done first.  (Found as run-to-run different LayerNorm / bias gradients of the block above whenever D != 64 --
      // this path -- ran with the side stream on: profiles/race_where.py.)
      if (side && have_mr) { DG_CUDA(cudaStreamWaitEvent(st, sd->mr, 0)); have_mr = false; }
      GemmArgs g;
      g.M = (int)R; g.N = d.M; g.K = d.D;
      g.A = dxop; g.a_sm = d.D; g.a_sk = 1;
      g.B = WSel<A>::w(net, b.fc2_w); g.b_sk = d.M; g.b_sn = 1;
      g.C = c.dH; g.ldc = d.M;
      g.epi = EPI_GELU_BWD; g.aux = B_.Hpre; g.ldaux = d.M;
      gemm<A, A, A>(g, st);
      linear_bwd_w<A, A>(dxop, B_.Hact, G + b.fc2_w, last ? G + b.fc2_b : nullptr, R, d.D, d.M, c.partial, st);
      linear_bwd_w<A, A>(c.dH, B_.Xn2, G + b.fc1_w, G + b.fc1_b, R, d.M, d.D, c.partial, st);
      linear_bwd_x<A, A, float>(c.dH, WSel<A>::w(net, b.fc1_w), c.dXn, R, d.M, d.D, EPI_NONE, nullptr, 0, st);
    }
    }
    // the previous block's multi-reduce (side stream) must have read its partial sums before this block overwrites them
    if (side && have_mr) DG_CUDA(cudaStreamWaitEvent(st, sd->mr, 0));
    // dXr becomes dL/dX_m, whose column sums are the to_out.0.bias gradient
    launch_ln_bwd(c.dXn, B_.Xm, B_.mean2, B_.rstd2, P + b.ln2_w, dXr, dXr_lp, G + b.ln2_w, G + b.ln2_b, G + b.out_b,
                  c.partial, R, d.D, st, &rl);
    // ---- attention block.  dXr = dL/dX_m
    side_after_main();
    if (!last) {
      linear_bwd_w<A, A>(dxop_m, B_.O, G + b.out_w, nullptr, d.T, d.D, d.inner, c.partial, ss, -1, -1, &rl);
      linear_bwd_x<A, A, A>(dxop_m, WSel<A>::w(net, b.out_w), c.dO, d.T, d.D, d.inner, EPI_NONE, nullptr, 0, st);
    } else {
      const int64_t ostride = (int64_t)d.N * d.inner;
      linear_bwd_w<A, A>(dxop_m, B_.O, G + b.out_w, nullptr, d.B, d.D, d.inner, c.partial, ss, -1, ostride, &rl);
      // dO is zero except on the token-0 rows (the single-query-row attention backward reads only those)
      if (!(d.dh == R0_DH && d.N <= R0_MAXN && (g_row0_mode & 2) && ((((uintptr_t)B_.QKV) | ((uintptr_t)c.dO) | ((uintptr_t)c.dQKV)) & 15) == 0))
        DG_CUDA(cudaMemsetAsync(c.dO, 0, (size_t)d.T * d.inner * sizeof(A), st));
      linear_bwd_x<A, A, A>(dxop_m, WSel<A>::w(net, b.out_w), c.dO, d.B, d.D, d.inner, EPI_NONE, nullptr, 0, st, -1,
                            nullptr, ostride);
      // (the LayerNorm-1 backward below reads the compact residual gradient directly: zero off token 0)
    }
    if (!(last && launch_attention_row0<A>(B_.QKV, (A*)nullptr, (const A*)c.dO, c.dQKV, d, st)))
      launch_attention_bwd<A>(B_.QKV, B_.O, c.dO, c.dQKV, d, st, B_.lse, c.attn_delta);
    side_after_main();
    linear_bwd_w<A, A>(c.dQKV, B_.Xn1, G + b.qkv_w, nullptr, d.T, 3 * d.inner, d.D, c.partial, ss, -1, -1, &rl);
    linear_bwd_x<A, A, float>(c.dQKV, WSel<A>::w(net, b.qkv_w), c.dXn, d.T, 3 * d.inner, d.D, EPI_NONE, nullptr, 0, st);
    // the MLP dW launch reads the incoming gradient copy that the next kernel overwrites
    // (only when this block's fused-MLP weight gradient really went to the side stream: waiting on an event this capture
    // never recorded is an error under CUDA-graph capture)
    if (side && have_dw) { DG_CUDA(cudaStreamWaitEvent(st, sd->dw, 0)); have_dw = false; }
    // c.dX becomes dL/dX_a = gradient of the previous block's output: its column sums are that block's net.3.bias gradient
    launch_ln_bwd(c.dXn, B_.Xa, B_.mean1, B_.rstd1, P + b.ln1_w, c.dX, c.dXh, G + b.ln1_w, G + b.ln1_b,
                  l > 0 ? G + L.block[l - 1].fc2_b : nullptr, c.partial, d.T, d.D, st, &rl, last ? c.dXc : nullptr, d.N);
    side_after_main();
    rl.launch(ss);
    if (side) { DG_CUDA(cudaEventRecord(sd->mr, ss)); have_mr = true; }
  }
  // ---- embedding.  c.dX = dL/dX0 (post-dropout)
  const int64_t tot = d.T * d.D;
  launch_k(embed_bwd_kernel<A>, grid1d(tot / 4), 256, 0, st, c.dX, c.tok, c.dXp, c.dtok, drop, tot, d.N, d.D, relu_tok);
  DG_LAUNCH_CHECK();
  {
    const int S = std::max(1, std::min(32, d.B / 8));
    const int64_t n = (int64_t)d.N * d.D;
    float* part = rl_misc.alloc((size_t)S * n);
    launch_k(dpos_kernel, dim3(d.N, S), 128, 0, st, c.dX, part, drop, d.B, d.N, d.D);
    DG_LAUNCH_CHECK();
    if (n % 4 == 0) rl_misc.add(part, G + L.pos, S, n, n);
    else {
      launch_k(reduce_partials_kernel, reduce_grid(n), 256, 0, st, part, G + L.pos, S, n);
      DG_LAUNCH_CHECK();
    }
  }
  // patch-weight gradient dW = dXp^T patches.  bf16 path: no patch matrix exists (the forward is im2col-free); the
  // weight-gradient kernel rebuilds the frame tiles in shared memory as its MN-major operand (patch_tc.cuh)
  bool patch_dw_done = false;
#ifdef DGVIT_WITH_TC
  if constexpr (std::is_same<A, bf16>::value) {
    if (img && patch::g_dw_enabled && patch::eligible(net.cfg, img, WSel<A>::w(net, L.patch_w), c.L[0].Xn1) && ((((uintptr_t)c.dXp) & 15) == 0)) {
      const int64_t n_tok = (int64_t)d.B * d.P;
      ProfScope ps_all(PROF_GEMM_ALL, 2.0 * n_tok * d.D * d.pd, 0.0, st);
      patch::bwd_w(net.cfg, img, (const bf16*)c.dXp, n_tok, G + L.patch_w, rl_misc, st);
      bias_grad_colsum<A>(c.dXp, d.D, G + L.patch_b, n_tok, d.D, rl_misc, st);
      patch_dw_done = true;
    } else if (!c.Pm_ext && patch::eligible(net.cfg, img, WSel<A>::w(net, L.patch_w), c.L[0].Xn1)) {
      DG_REQUIRE(img != nullptr, "trunk_backward: frames needed to rebuild the patch matrix");
      launch_patchify<A>(img, c.Pm, d.B, d, net.cfg, st);
    }
  }
#endif
  if (!patch_dw_done)
    linear_bwd_w<A, A>(c.dXp, c.Pm_ext ? c.Pm_ext : c.Pm, G + L.patch_w, G + L.patch_b, (int64_t)d.B * d.P, d.D, d.pd, c.partial, st, -1,
                       -1, &rl_misc);
  rl_misc.launch(st);
  if (side && have_mr) DG_CUDA(cudaStreamWaitEvent(st, sd->mr, 0));     // join
}

// zero the gradient ranges the reference leaves as None (so the arena is fully defined)
static void zero_unused_grads(const dgvit_net& net, const dgvit_layout& L, cudaStream_t st) {
  for (int k = 0; k < L.n_skip; ++k) {
    // the alpha-gradient slot is written by the policy-loss kernel before the actor backward
    if (net.cfg.kind == DGVIT_ACTOR && L.skip_begin[k] == L.alpha_grad_slot) continue;
    DG_CUDA(cudaMemsetAsync(net.grads + L.skip_begin[k], 0, (L.skip_end[k] - L.skip_begin[k]) * sizeof(float), st));
  }
}

// ------------------------------------------------------------------ actor
template <typename A>
struct ActorCtx {
  TrunkCtx<A> t;
  float *h1, *h2, *mean_raw, *lstd_raw;  // [B,128] [B,128] [B,na] [B,na]
  float *eps;                            // [B,na]
  float *dmean, *dlstd, *dh2, *dh1;
};
template <typename A>
static void carve_actor(Carver& cv, const Dims& d, bool save, ActorCtx<A>& c) {
  carve_trunk<A>(cv, d, save, c.t);
  c.h1 = cv.take<float>((int64_t)d.B * 128);
  c.h2 = cv.take<float>((int64_t)d.B * 128);
  c.mean_raw = cv.take<float>((int64_t)d.B * d.na);
  c.lstd_raw = cv.take<float>((int64_t)d.B * d.na);
  c.eps = cv.take<float>((int64_t)d.B * d.na);
  c.dmean = c.dlstd = c.dh2 = c.dh1 = nullptr;
  if (save) {
    c.dmean = cv.take<float>((int64_t)d.B * d.na);
    c.dlstd = cv.take<float>((int64_t)d.B * d.na);
    c.dh2 = cv.take<float>((int64_t)d.B * 128);
    c.dh1 = cv.take<float>((int64_t)d.B * 128);
  }
}

// GoTPolicy.forward + .sample (vn/got_sac_network.py:221-251)
template <typename A>
static void actor_forward(const dgvit_net& net, const dgvit_layout& L, const Dims& d, const dgvit_actor_io& io,
                          ActorCtx<A>& c, cudaStream_t st) {
  const float* P = net.params;
  DG_REQUIRE(io.img && io.pstate && io.action_scale && io.action_bias, "actor_forward: null input");
  DG_REQUIRE(io.eps || io.drop.rng_state, "actor_forward: provide eps or drop.rng_state");
  const DropDev drop = make_drop(io.drop, d, io.sample_offset);
  const GoalTok gt{io.pstate, P + L.embed_w, P + L.embed_b, d.nps, 0};      // fc_embed, no activation (got_sac_network.py:224)
  trunk_forward<A>(net, L, d, io.img, drop, c.t, st, gt, /*fuse_rms=*/true);
  {  // cls pooling + RMSNorm -> fc1 -> relu -> fc2 -> relu -> (mean_linear | log_std_linear), one launch
    heads::FwdArgs h;
    memset(&h, 0, sizeof(h));
    h.nheads = 1; h.B = d.B; h.K1 = d.D; h.K2 = 0; h.H2 = 128; h.NOa = d.na; h.NOb = d.na;
    h.x1 = c.t.z; h.xraw = c.t.Xout; h.rms_g = P + L.rms_g; h.z_out = c.t.z;
    h.w[0] = heads::HeadW{P + L.fc1_w, P + L.fc1_b, P + L.fc2_w, P + L.fc2_b, P + L.mean_w, P + L.mean_b,
                          P + L.lstd_w, P + L.lstd_b, c.h1, c.h2, c.mean_raw, c.lstd_raw};
    heads::launch_fwd(h, st);
  }
  SampleArgs s;
  s.B = d.B; s.na = d.na;
  s.mean = c.mean_raw; s.lstd_raw = c.lstd_raw;
  s.scale = io.action_scale; s.bias = io.action_bias;
  s.eps = io.eps; s.rng = io.drop.rng_state; s.stream_id = io.drop.stream_id ^ 0x5a5a0000u;
  s.sample_offset = io.sample_offset;
  s.mean_out = io.mean; s.log_std = io.log_std; s.action = io.action; s.log_prob = io.log_prob; s.mean_t = io.mean_t;
  s.eps_out = c.eps;
  s.rng_bump = (io.advance_rng && io.drop.rng_state && d.B <= 128) ? const_cast<uint64_t*>(io.drop.rng_state) : nullptr;
  launch_k(actor_sample_kernel, (unsigned)cdiv(d.B, 128), 128, 0, st, s);
  DG_LAUNCH_CHECK();
  if (io.eps_out)
    DG_CUDA(cudaMemcpyAsync(io.eps_out, c.eps, (size_t)d.B * d.na * sizeof(float), cudaMemcpyDeviceToDevice, st));
}

template <typename A>
static void actor_backward(const dgvit_net& net, const dgvit_layout& L, const Dims& d, const dgvit_actor_io& io,
                           const dgvit_actor_grad& g, const float* alpha_dev, float alpha_scale, ActorCtx<A>& c,
                           cudaStream_t st) {
  const float* P = net.params;
  float* G = net.grads;
  zero_unused_grads(net, L, side_fork(c.t, st));   // off the dependency chain; joined by trunk_backward
  SampleBwdArgs s;
  s.B = d.B; s.na = d.na;
  s.mean = c.mean_raw; s.lstd_raw = c.lstd_raw; s.eps = c.eps; s.scale = io.action_scale;
  s.d_mean = g.d_mean; s.d_log_std = g.d_log_std; s.d_action = g.d_action; s.d_log_prob = g.d_log_prob;
  s.d_mean_t = g.d_mean_t; s.d_log_prob_const = g.d_log_prob_const;
  s.d_log_prob_dev = alpha_dev; s.d_log_prob_dev_scale = alpha_scale;
  s.d_mean_out = c.dmean; s.d_lstd_out = c.dlstd;
  launch_k(actor_sample_bwd_kernel, (unsigned)cdiv((int64_t)d.B * d.na, 128), 128, 0, st, s);
  DG_LAUNCH_CHECK();
  {
    heads::BwdArgs h;
    memset(&h, 0, sizeof(h));
    h.nheads = 1; h.B = d.B; h.K0 = d.D; h.H2 = 128; h.NOa = d.na; h.NOb = d.na;
    h.h[0] = heads::BwdHead{P + L.fc1_w, P + L.fc2_w, P + L.mean_w, P + L.lstd_w, c.h1, c.h2, c.dmean, c.dlstd,
                            c.dh1, c.dh2, c.t.dz};
    heads::launch_bwd_dx(h, st);
    heads::DwList dw(d.B);
    dw.add(c.dmean, d.na, c.h2, G + L.mean_w, G + L.mean_b, d.na, 128);
    dw.add(c.dlstd, d.na, c.h2, G + L.lstd_w, G + L.lstd_b, d.na, 128);
    dw.add(c.dh2, 128, c.h1, G + L.fc2_w, G + L.fc2_b, 128, 128);
    dw.add(c.dh1, 128, c.t.z, G + L.fc1_w, G + L.fc1_b, 128, d.D);
    dw.launch(side_fork(c.t, st));      // joined at the end of trunk_backward
  }
  const DropDev drop = make_drop(io.drop, d, io.sample_offset);
  trunk_backward<A>(net, L, d, drop, c.t, /*relu_tok=*/0, st, io.img);
  {  // fc_embed: dW[D,nps] = dtok^T pstate ; db = colsum(dtok)
    heads::DwList dw(d.B);
    dw.add(c.t.dtok, d.D, io.pstate, G + L.embed_w, G + L.embed_b, d.D, d.nps);
    dw.launch(st);
  }
}

// ------------------------------------------------------------------ critic
template <typename A>
struct CriticCtx {
  TrunkCtx<A> t;
  float *xcat, *h1a, *h2a, *h1b, *h2b;  // [B,D+na] [B,128] [B,32] x2
  float *dh2, *dh1, *dh2b, *dh1b, *dxa, *dxb;
};
template <typename A>
static void carve_critic(Carver& cv, const Dims& d, bool save, CriticCtx<A>& c) {
  carve_trunk<A>(cv, d, save, c.t);
  c.xcat = cv.take<float>((int64_t)d.B * (d.D + d.na));
  c.h1a = cv.take<float>((int64_t)d.B * 128);
  c.h2a = cv.take<float>((int64_t)d.B * 32);
  c.h1b = cv.take<float>((int64_t)d.B * 128);
  c.h2b = cv.take<float>((int64_t)d.B * 32);
  c.dh2 = cv.take<float>((int64_t)d.B * 32);
  c.dh1 = cv.take<float>((int64_t)d.B * 128);
  c.dh2b = cv.take<float>((int64_t)d.B * 32);
  c.dh1b = cv.take<float>((int64_t)d.B * 128);
  c.dxa = cv.take<float>((int64_t)d.B * (d.D + d.na));
  c.dxb = cv.take<float>((int64_t)d.B * (d.D + d.na));
}

static void critic_heads_forward(const dgvit_net& net, const dgvit_layout& L, const Dims& d, float* z,
                                 const float* action, float* xcat, float* h1a, float* h2a, float* h1b, float* h2b,
                                 float* q1, float* q2, cudaStream_t st, const float* xraw = nullptr) {
  const float* P = net.params;
  heads::FwdArgs h;
  memset(&h, 0, sizeof(h));
  h.nheads = 2; h.B = d.B; h.K1 = d.D; h.K2 = d.na; h.H2 = 32; h.NOa = d.na; h.NOb = 0;
  h.x1 = z; h.x2 = action; h.xcat = xcat;
  if (xraw) { h.xraw = xraw; h.rms_g = P + L.rms_g; h.z_out = z; }   // cls pooling + RMSNorm while staging the input
  h.w[0] = heads::HeadW{P + L.fc1_w, P + L.fc1_b, P + L.fc2_w, P + L.fc2_b, P + L.fc3_w, P + L.fc3_b, nullptr, nullptr,
                        h1a, h2a, q1, nullptr};
  h.w[1] = heads::HeadW{P + L.fc11_w, P + L.fc11_b, P + L.fc21_w, P + L.fc21_b, P + L.fc31_w, P + L.fc31_b, nullptr,
                        nullptr, h1b, h2b, q2, nullptr};
  heads::launch_fwd(h, st);
}

template <typename A>
static void critic_forward(const dgvit_net& net, const dgvit_layout& L, const Dims& d, const dgvit_critic_io& io,
                           int64_t sample_offset, CriticCtx<A>& c, cudaStream_t st, cudaStream_t trunk_st = nullptr,
                           cudaEvent_t trunk_done = nullptr) {
  const float* P = net.params;
  DG_REQUIRE(io.img && io.pstate && io.action && io.q1 && io.q2, "critic_forward: null input/output");
  const DropDev drop = make_drop(io.drop, d, sample_offset);
  const GoalTok gt{io.pstate, P + L.embed_w, P + L.embed_b, d.nps, 1};      // relu(fc_embed) (got_sac_network.py:111)
  // the action enters at the heads only: the trunk may run on another stream, beside whatever produces the action
  if (trunk_st && trunk_st != st) {
    trunk_forward<A>(net, L, d, io.img, drop, c.t, trunk_st, gt, /*fuse_rms=*/true);
    DG_CUDA(cudaEventRecord(trunk_done, trunk_st));
    DG_CUDA(cudaStreamWaitEvent(st, trunk_done, 0));
  } else {
    trunk_forward<A>(net, L, d, io.img, drop, c.t, st, gt, /*fuse_rms=*/true);
  }
  critic_heads_forward(net, L, d, c.t.z, io.action, c.xcat, c.h1a, c.h2a, c.h1b, c.h2b, io.q1, io.q2, st, c.t.Xout);
}

template <typename A>
static void critic_backward(const dgvit_net& net, const dgvit_layout& L, const Dims& d, const dgvit_critic_io& io,
                            int64_t sample_offset, const float* dq1, const float* dq2, float* d_action,
                            bool param_grads, CriticCtx<A>& c, float* part_fallback, cudaStream_t st) {
  const float* P = net.params;
  float* G = net.grads;
  (void)part_fallback;
  if (param_grads) zero_unused_grads(net, L, side_fork(c.t, st));   // off the dependency chain; joined by trunk_backward
  {
    heads::BwdArgs h;
    memset(&h, 0, sizeof(h));
    h.nheads = 2; h.B = d.B; h.K0 = d.D + d.na; h.H2 = 32; h.NOa = d.na; h.NOb = 0;
    h.h[0] = heads::BwdHead{P + L.fc1_w, P + L.fc2_w, P + L.fc3_w, nullptr, c.h1a, c.h2a, dq1, nullptr, c.dh1, c.dh2, c.dxa};
    h.h[1] = heads::BwdHead{P + L.fc11_w, P + L.fc21_w, P + L.fc31_w, nullptr, c.h1b, c.h2b, dq2, nullptr, c.dh1b, c.dh2b,
                            c.dxb};
    heads::launch_bwd_dx(h, st);
    if (param_grads) {
      const int W = d.D + d.na;
      heads::DwList dw(d.B);
      dw.add(dq1, d.na, c.h2a, G + L.fc3_w, G + L.fc3_b, d.na, 32);
      dw.add(c.dh2, 32, c.h1a, G + L.fc2_w, G + L.fc2_b, 32, 128);
      dw.add(c.dh1, 128, c.xcat, G + L.fc1_w, G + L.fc1_b, 128, W);
      dw.add(dq2, d.na, c.h2b, G + L.fc31_w, G + L.fc31_b, d.na, 32);
      dw.add(c.dh2b, 32, c.h1b, G + L.fc21_w, G + L.fc21_b, 32, 128);
      dw.add(c.dh1b, 128, c.xcat, G + L.fc11_w, G + L.fc11_b, 128, W);
      dw.launch(side_fork(c.t, st));    // joined at the end of trunk_backward
    }
  }
  const int W = d.D + d.na;
  launch_k(split_dza_kernel, (unsigned)cdiv((int64_t)d.B * W, 256), 256, 0, st, c.dxa, c.dxb, param_grads ? c.t.dz : nullptr,
                                                                          d_action, d.B, d.D, d.na);
  DG_LAUNCH_CHECK();
  if (param_grads) {
    const DropDev drop = make_drop(io.drop, d, sample_offset);
    trunk_backward<A>(net, L, d, drop, c.t, /*relu_tok=*/1, st, io.img);
    heads::DwList dw(d.B);
    dw.add(c.t.dtok, d.D, io.pstate, G + L.embed_w, G + L.embed_b, d.D, d.nps);
    dw.launch(st);
  }
}

// ------------------------------------------------------------------ optimizer
// data-parallel arguments of one network's fused all-reduce + Adam pass (which: 0 critic, 1 actor); dp == null: off
static DpDev make_dp(const dgvit_dp* dp, int which, const dgvit_layout& L) {
  DpDev d;
  memset(&d, 0, sizeof(d));
  if (!dp) return d;
  d.world = dp->world; d.rank = dp->rank; d.which = which;
  d.arena_off = dp->arena_off[which];
  d.mc = dp->multicast ? dp->multicast + d.arena_off : nullptr;
  d.peers = dp->peers; d.pads = dp->pads;
  d.error_flag = dp->error_flag;
  d.reduced_out = dp->reduced_out[which];
  if (which == 1) { d.tail_out = dp->tail; d.tail_begin = L.alpha_grad_slot; d.tail_n = DGVIT_ALIGN_FLOATS; }
  return d;
}
static void dp_wait_done(const dgvit_dp* dp, int which, const dgvit_layout& L, const int64_t* step, cudaStream_t st) {
  if (!dp) return;
  launch_k(dp_wait_done_kernel, 1, 32, 0, st, make_dp(dp, which, L), step);
  DG_LAUNCH_CHECK();
}

static void fill_adam_args(AdamArgs& a, const dgvit_net& net, const dgvit_layout& L, const dgvit_adam& o) {
  memset(&a, 0, sizeof(a));
  a.p = net.params; a.g = net.grads; a.m = o.m; a.v = o.v;
  a.n = L.total; a.step = o.step;
  a.lr = o.lr; a.b1 = o.beta1; a.b2 = o.beta2; a.eps = o.eps;
  a.omb1 = (float)(1.0 - (double)o.beta1);
  a.omb2 = (float)(1.0 - (double)o.beta2);
  a.n_skip = L.n_skip;
  for (int k = 0; k < 4; ++k) { a.skip_b[k] = L.skip_begin[k]; a.skip_e[k] = L.skip_end[k]; }
}

static void adam_step(const dgvit_net& net, const dgvit_layout& L, const dgvit_adam& o, const dgvit_net* tgt,
                      float tau, bool want_shadow, cudaStream_t st, const dgvit_dp* dp = nullptr, int which = 0,
                      const float* gscale = nullptr) {
  DG_REQUIRE(o.m && o.v && o.step, "adam: null state");
  launch_k(step_bump_kernel, 1, 32, 0, st, o.step);
  DG_LAUNCH_CHECK();
  AdamArgs a;
  fill_adam_args(a, net, L, o);
  a.gscale = gscale;
  a.shadow = (want_shadow && net.shadow) ? (bf16*)net.shadow : nullptr;
  a.tgt = tgt ? tgt->params : nullptr;
  a.tgt_shadow = (tgt && want_shadow && tgt->shadow) ? (bf16*)tgt->shadow : nullptr;
  a.tau = tau;
  {  // algorithmic bytes: theta, m, v read + written, g read, bf16 shadow written (+ target read / written + its shadow)
    int64_t used = L.total;
    for (int k = 0; k < L.n_skip; ++k) used -= L.skip_end[k] - L.skip_begin[k];
    const double per = 28.0 + (a.shadow ? 4.0 : 0.0);
    const double tgt_b = a.tgt ? (double)L.total * (12.0 + (a.tgt_shadow ? 4.0 : 0.0)) : 0.0;
    ProfScope ps(PROF_ADAM, 0.0, used * per + tgt_b, st);
    if (dp) {
      DG_REQUIRE(dp->world >= 2 && dp->rank >= 0 && dp->rank < dp->world && dp->peers && dp->pads && dp->finished && dp->tail,
                 "dgvit_dp: incomplete");
      launch_k(adam_polyak_kernel<true>, 148 * 8, 256, 0, st, a, make_dp(dp, which, L), dp->finished + which);
    } else {
      launch_k(adam_polyak_kernel<false>, 148 * 8, 256, 0, st, a, DpDev(), (unsigned int*)nullptr);
    }
    DG_LAUNCH_CHECK();
    no_pdl_once().mark_strict(st);      // the next kernel on this stream may read the parameters just rewritten (bias / LayerNorm /
                                        // position loads go through the non-coherent path): it must not start before this one ends
  }
}

// ------------------------------------------------------------------ SAC.learn
template <typename A>
struct SacWs {
  ActorCtx<A> actor_s;      // saved: policy.sample(s)
  CriticCtx<A> critic_s;    // saved: critic(s, a)
  ActorCtx<A> actor_tmp;    // unsaved: policy.sample(s')
  CriticCtx<A> critic_tmp;  // unsaved: critic_target(s', a') and critic(s, pi)
  float *a2, *logp2, *mean_tmp, *lstd_tmp, *q1t, *q2t, *q1, *q2, *dq1, *dq2, *nq;
  float *pi, *logpi, *mean_pi, *lstd_pi, *q1p, *q2p, *dpi;
  float *meant_pi, *dlogp, *dmeant;   // learn_guidence: per-row gradients of the actor's (B + n_extra)-row pass
};
template <typename A>
static void carve_sac(Carver& cv, const Dims& d, const Dims& da, SacWs<A>& w) {
  carve_actor<A>(cv, da, true, w.actor_s);
  carve_critic<A>(cv, d, true, w.critic_s);
  carve_actor<A>(cv, d, false, w.actor_tmp);
  carve_critic<A>(cv, d, false, w.critic_tmp);
  const int64_t bn = (int64_t)d.B * d.na;
  w.a2 = cv.take<float>(bn); w.logp2 = cv.take<float>(d.B);
  w.mean_tmp = cv.take<float>(bn); w.lstd_tmp = cv.take<float>(bn);
  w.q1t = cv.take<float>(bn); w.q2t = cv.take<float>(bn);
  w.q1 = cv.take<float>(bn); w.q2 = cv.take<float>(bn);
  w.dq1 = cv.take<float>(bn); w.dq2 = cv.take<float>(bn); w.nq = cv.take<float>(bn);
  const int64_t an = (int64_t)da.B * d.na;
  w.pi = cv.take<float>(an); w.logpi = cv.take<float>(da.B);
  w.mean_pi = cv.take<float>(an); w.lstd_pi = cv.take<float>(an);
  w.q1p = cv.take<float>(bn); w.q2p = cv.take<float>(bn); w.dpi = cv.take<float>(an);
  w.meant_pi = cv.take<float>(an); w.dlogp = cv.take<float>(da.B); w.dmeant = cv.take<float>(an);
}

static dgvit_drop sac_drop(const dgvit_sac& s, const dgvit_noise* nz, const uint8_t* mask, uint32_t id) {
  dgvit_drop dr;
  dr.mode = nz ? nz->drop_mode : DGVIT_DROP_RNG;
  dr.p = 0.1f;
  dr.keep_mask = mask;
  dr.rng_state = s.rng_state;
  dr.stream_id = id;
  return dr;
}

// The three forward passes of the first half of the update are independent of one another
// (vn/DRL.py:388-395,404): [policy.sample(s') -> critic_target(s',a')], critic(s,a) and
// policy.sample(s).  They are issued on three streams (fork/join with events: capturable into
// one CUDA graph) so that their many small, latency-bound kernels fill each other's gaps.
struct ForkState {
  cudaStream_t aux[3];
  cudaEvent_t fork, fork2, fork3, fork4, join[3];
};
static ForkState& fork_state() {
  static ForkState per_dev[64];
  static bool inited[64] = {};
  const int dev = current_device();
  ForkState& f = per_dev[dev];
  bool& init = inited[dev];
  if (!init) {
    for (int i = 0; i < 3; ++i) {
      const char* pe = getenv("DGVIT_AUX_PRIO");
      DG_CUDA(cudaStreamCreateWithPriority(&f.aux[i], cudaStreamNonBlocking, pe ? atoi(pe) : 0));
      DG_CUDA(cudaEventCreateWithFlags(&f.join[i], cudaEventDisableTiming));
    }
    DG_CUDA(cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&f.fork2, cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&f.fork3, cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&f.fork4, cudaEventDisableTiming));
    init = true;
  }
  return f;
}

template <typename A>
static void sac_phase1(const dgvit_sac& s, const dgvit_batch& b, const dgvit_noise* nz, const dgvit_sac_out& out,
                       const Dims& d, const Dims& da, SacWs<A>& w, cudaStream_t st) {
  dgvit_layout La, Lc;
  make_layout(s.actor.cfg, La);
  make_layout(s.critic.cfg, Lc);
  ForkState& f0 = fork_state();
  ForkState f = f0;
  if (!g_fork_enabled) { f.aux[0] = st; f.aux[1] = st; f.aux[2] = st; }      // single-stream mode (per-kernel timing)
  // the five forward passes of an update see two distinct frame batches: patchify s and s' once, before the fork
  // (s' on the caller's stream, s on the first forked stream; the second forked stream starts after the latter)
  DG_CUDA(cudaEventRecord(f.fork, st));
  DG_CUDA(cudaStreamWaitEvent(f.aux[0], f.fork, 0));
  bool fused_embed = false;     // im2col-free patch embedding and patch-weight gradient: no patch matrix at all
#ifdef DGVIT_WITH_TC
  if constexpr (std::is_same<A, bf16>::value)
    fused_embed = patch::g_dw_enabled &&
                  patch::eligible(s.actor.cfg, b.next_obs, WSel<A>::w(s.actor, La.patch_w), w.actor_tmp.t.L[0].Xn1) &&
                  patch::eligible(s.actor.cfg, b.obs, WSel<A>::w(s.critic, Lc.patch_w), w.actor_s.t.L[0].Xn1);
#endif
  if (!fused_embed) launch_patchify<A>(b.obs, w.actor_s.t.Pm, da.B, da, s.actor.cfg, f.aux[0]);
  DG_CUDA(cudaEventRecord(f.fork2, f.aux[0]));
  DG_CUDA(cudaStreamWaitEvent(f.aux[1], f.fork2, 0));
  if (!fused_embed) launch_patchify<A>(b.next_obs, w.actor_tmp.t.Pm, d.B, d, s.actor.cfg, st);
  w.actor_s.t.Pm_ext = w.actor_s.t.Pm; w.critic_s.t.Pm_ext = w.actor_s.t.Pm;
  w.actor_tmp.t.Pm_ext = w.actor_tmp.t.Pm; w.critic_tmp.t.Pm_ext = w.actor_tmp.t.Pm;
  // ---- stream 0 (caller's): a', log pi' = policy.sample(s') ; q_t = critic_target(s', a')   (DRL.py:388-393)
  dgvit_actor_io ai; memset(&ai, 0, sizeof(ai));
  ai.img = b.next_obs; ai.pstate = b.next_pobs; ai.eps = nz ? nz->eps_next : nullptr;
  ai.action_scale = s.action_scale; ai.action_bias = s.action_bias;
  ai.drop = sac_drop(s, nz, nz ? nz->mask_a_next : nullptr, 1);
  ai.sample_offset = s.sample_offset;
  ai.mean = w.mean_tmp; ai.log_std = w.lstd_tmp; ai.action = w.a2; ai.log_prob = w.logp2;
  // ---- stream 2: pi, log_pi = policy.sample(s)  (DRL.py:404; the actor is not touched before :413,
  //      so this forward can run beside the whole critic update)
  dgvit_actor_io ap; memset(&ap, 0, sizeof(ap));
  ap.img = b.obs; ap.pstate = b.pobs; ap.eps = nz ? nz->eps_pi : nullptr;
  ap.action_scale = s.action_scale; ap.action_bias = s.action_bias;
  ap.drop = sac_drop(s, nz, nz ? nz->mask_a : nullptr, 4);
  ap.sample_offset = s.sample_offset;
  ap.mean = w.mean_pi; ap.log_std = w.lstd_pi; ap.action = w.pi; ap.log_prob = w.logpi; ap.mean_t = w.meant_pi;
  auto launch_actor_s = [&](int when) {      // B rows (+ the imitation rows of learn_guidence)
    if (when != g_actor_s_when) return;
    if (when > 0 && f.aux[1] != st) {        // start it later: after this point of the caller's stream
      DG_CUDA(cudaEventRecord(f.fork3, st));
      DG_CUDA(cudaStreamWaitEvent(f.aux[1], f.fork3, 0));
    }
    actor_forward<A>(s.actor, La, da, ap, w.actor_s, f.aux[1]);
    DG_CUDA(cudaEventRecord(f.join[1], f.aux[1]));
  };
  launch_actor_s(0);
  if (g_target_fork && f.aux[2] != st) {     // critic_target's trunk starts here (after the patch matrix of s', if there is one)
    DG_CUDA(cudaEventRecord(f.fork4, st));
    DG_CUDA(cudaStreamWaitEvent(f.aux[2], f.fork4, 0));
  }
  actor_forward<A>(s.actor, La, d, ai, w.actor_tmp, st);
  launch_actor_s(1);
  dgvit_critic_io ci; memset(&ci, 0, sizeof(ci));
  ci.img = b.next_obs; ci.pstate = b.next_pobs; ci.action = w.a2;
  ci.drop = sac_drop(s, nz, nz ? nz->mask_ct : nullptr, 2);
  ci.q1 = w.q1t; ci.q2 = w.q2t;
  // (its trunk does not need a': with "target_fork" it starts at the fork on a stream of its own, the heads follow policy.sample(s'))
  critic_forward<A>(s.critic_target, Lc, d, ci, s.sample_offset, w.critic_tmp, st, g_target_fork ? f.aux[2] : nullptr, f.join[2]);
  // ---- stream 1: critic(s, a)                                                   (DRL.py:395)
  dgvit_critic_io cs; memset(&cs, 0, sizeof(cs));
  cs.img = b.obs; cs.pstate = b.pobs; cs.action = b.act;
  cs.drop = sac_drop(s, nz, nz ? nz->mask_c : nullptr, 3);
  cs.q1 = w.q1; cs.q2 = w.q2;
  critic_forward<A>(s.critic, Lc, d, cs, s.sample_offset, w.critic_s, f.aux[0]);
  DG_CUDA(cudaEventRecord(f.join[0], f.aux[0]));
  launch_actor_s(2);
  // ---- losses + critic backward                                                 (DRL.py:396-401)
  DG_CUDA(cudaStreamWaitEvent(st, f.join[0], 0));
  launch_k(critic_loss_kernel, 1, 1024, 0, st, w.q1, w.q2, w.q1t, w.q2t, w.logp2, b.rew, s.alpha, s.gamma, d.B, d.na,
                                         s.global_batch, w.nq, w.dq1, w.dq2, out.losses);
  DG_LAUNCH_CHECK();
  dp_wait_done(s.dp, 0, Lc, s.critic_opt.step, st);      // (data parallel) peers have read the previous critic gradients
  critic_backward<A>(s.critic, Lc, d, cs, s.sample_offset, w.dq1, w.dq2, nullptr, true, w.critic_s, nullptr, st);
  if (out.debug) {
    const size_t bn = (size_t)d.B * d.na * sizeof(float);
    DG_CUDA(cudaMemcpyAsync(out.debug, w.nq, bn, cudaMemcpyDeviceToDevice, st));
    DG_CUDA(cudaMemcpyAsync(out.debug + d.B * d.na, w.q1, bn, cudaMemcpyDeviceToDevice, st));
    DG_CUDA(cudaMemcpyAsync(out.debug + 2 * d.B * d.na, w.q2, bn, cudaMemcpyDeviceToDevice, st));
  }
  DG_CUDA(cudaStreamWaitEvent(st, f.join[1], 0));      // join the actor forward before the phase ends
}

template <typename A>
static void sac_phase2(const dgvit_sac& s, const dgvit_batch& b, const dgvit_noise* nz, const dgvit_sac_out& out,
                       const Dims& d, const Dims& da, SacWs<A>& w, cudaStream_t st) {
  dgvit_layout La, Lc;
  make_layout(s.actor.cfg, La);
  make_layout(s.critic.cfg, Lc);
  const bool shadow = s.precision == DGVIT_BF16;
  adam_step(s.critic, Lc, s.critic_opt, nullptr, 0.f, shadow, st, s.dp, 0);      // DRL.py:402 (+ gradient all-reduce when s.dp)
  // soft update of the target (DRL.py:430-431): the critic does not change again in this update and nothing below reads
  // the target, so it runs beside the policy half instead of at the end of the chain
  cudaStream_t lane = st;
  if (s.do_polyak) {
    lane = lane_fork(st);
    launch_k(polyak_kernel, 148 * 4, 256, 0, lane, s.critic_target.params, s.critic.params,
                                           shadow ? (bf16*)s.critic_target.shadow : nullptr, s.tau, Lc.total);
    DG_LAUNCH_CHECK();
  }
  // the patch matrix of s was written by phase 1 (same workspace) into the saved actor context
  w.critic_tmp.t.Pm_ext = w.actor_s.t.Pm; w.actor_s.t.Pm_ext = w.actor_s.t.Pm; w.critic_s.t.Pm_ext = w.actor_s.t.Pm;
  // ---- q_pi = critic(s, pi) with the UPDATED critic (pi, log_pi come from phase 1)   (DRL.py:406-407)
  dgvit_actor_io ai; memset(&ai, 0, sizeof(ai));
  ai.img = b.obs; ai.pstate = b.pobs; ai.eps = nz ? nz->eps_pi : nullptr;
  ai.action_scale = s.action_scale; ai.action_bias = s.action_bias;
  ai.drop = sac_drop(s, nz, nz ? nz->mask_a : nullptr, 4);
  ai.sample_offset = s.sample_offset;
  ai.mean = w.mean_pi; ai.log_std = w.lstd_pi; ai.action = w.pi; ai.log_prob = w.logpi;
  dgvit_critic_io ci; memset(&ci, 0, sizeof(ci));
  ci.img = b.obs; ci.pstate = b.pobs; ci.action = w.pi;
  ci.drop = sac_drop(s, nz, nz ? nz->mask_c_pi : nullptr, 5);
  ci.q1 = w.q1p; ci.q2 = w.q2p;
  critic_forward<A>(s.critic, Lc, d, ci, s.sample_offset, w.critic_tmp, st);
  launch_k(policy_loss_kernel, 1, 1024, 0, st, w.q1p, w.q2p, w.logpi, s.alpha, s.log_alpha, s.target_entropy, d.B, d.na,
                                         s.global_batch, w.dq1, w.dq2, out.losses, s.actor.grads + La.alpha_grad_slot);
  DG_LAUNCH_CHECK();
  // d policy_loss / d pi through the critic heads only (the critic's own parameter gradients
  // of this backward are discarded by the reference's next zero_grad, DRL.py:399)
  critic_backward<A>(s.critic, Lc, d, ci, s.sample_offset, w.dq1, w.dq2, w.dpi, false, w.critic_tmp,
                     w.actor_s.t.partial, st);
  dgvit_actor_grad ag; memset(&ag, 0, sizeof(ag));
  ag.d_action = w.dpi;
  dp_wait_done(s.dp, 1, La, s.actor_opt.step, st);
  if (da.B == d.B) {
    actor_backward<A>(s.actor, La, d, ai, ag, s.alpha, 1.0f / (float)s.global_batch, w.actor_s, st);
  } else {
    // learn_guidence: rows >= B carry only the imitation loss  weight_r * |tanh-mean_r - target_r|^2  (DRL.py:257-278)
    DG_REQUIRE(b.extra_target && b.extra_weight, "n_extra > 0 needs extra_target / extra_weight");
    launch_k(imitation_grad_kernel, 1, 1024, 0, st, (const float*)w.meant_pi, b.extra_target, b.extra_weight, (const float*)s.alpha,
             1.0f / (float)s.global_batch, d.B, da.B, d.na, w.dpi, w.dlogp, w.dmeant, out.losses);
    DG_LAUNCH_CHECK();
    ag.d_log_prob = w.dlogp;
    ag.d_mean_t = w.dmeant;
    actor_backward<A>(s.actor, La, da, ai, ag, nullptr, 0.f, w.actor_s, st);
  }
  if (out.debug) {
    const size_t bn = (size_t)d.B * d.na * sizeof(float);
    DG_CUDA(cudaMemcpyAsync(out.debug + 3 * d.B * d.na, w.pi, bn, cudaMemcpyDeviceToDevice, st));
    DG_CUDA(cudaMemcpyAsync(out.debug + 4 * d.B * d.na, w.q1p, bn, cudaMemcpyDeviceToDevice, st));
    DG_CUDA(cudaMemcpyAsync(out.debug + 5 * d.B * d.na, w.logpi, (size_t)d.B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  // the actor backward's own join has already ordered the side stream (and the soft update on it) before this point
  if (lane != st) lane_join(lane, st);
}

template <typename A>
static void sac_phase3(const dgvit_sac& s, cudaStream_t st) {
  dgvit_layout La;
  make_layout(s.actor.cfg, La);
  const bool shadow = s.precision == DGVIT_BF16;
  // temperature step and RNG counter beside the actor's Adam pass
  cudaStream_t lane = lane_fork(st);
  adam_step(s.actor, La, s.actor_opt, nullptr, 0.f, shadow, st, s.dp, 1);        // DRL.py:413
  if (s.dp) { lane_join(lane, st); lane = st; }       // the temperature step needs the REDUCED alpha gradient the pass above wrote
  if (s.auto_alpha) {                                                            // DRL.py:416-423
    launch_k(alpha_step_kernel, 1, 32, 0, lane, s.log_alpha, s.alpha, s.alpha_m, s.alpha_v, s.alpha_step,
                                        s.dp ? s.dp->tail : s.actor.grads + La.alpha_grad_slot, s.lr_alpha, 0.9f, 0.999f,
                                        (float)(1.0 - 0.9), (float)(1.0 - 0.999), 1e-8f);
    DG_LAUNCH_CHECK();
  }
  if (s.rng_state) {
    launch_k(rng_advance_kernel, 1, 32, 0, lane, s.rng_state);
    DG_LAUNCH_CHECK();
  }
  lane_join(lane, st);
}

// ------------------------------------------------------------------ SAC.learn with the CNN twin-Q critic
// critic_type != "Transformer" (vn/DRL.py:118-121, the shipped default vn/config.yaml:61): the critic and its target are
// QNetwork arenas (dgvit_qnet_layout), cfg.kind == DGVIT_QNET.  Same statements, same three phases and the same loss /
// Adam / Polyak kernels as the Transformer-critic update; only the critic passes differ (qnet.cuh, no dropout).
template <typename A>
struct SacQWs {
  ActorCtx<A> actor_s;      // saved: policy.sample(s)
  ActorCtx<A> actor_tmp;    // unsaved: policy.sample(s')
  qnet::Ws<A> cq;           // critic(s, a) forward + backward, then critic(s, pi) forward + d/d pi
  qnet::Ws<A> ct;           // critic_target(s', a')
  float *a2, *logp2, *mean_tmp, *lstd_tmp, *q1t, *q2t, *q1, *q2, *dq1, *dq2, *nq;
  float *pi, *logpi, *mean_pi, *lstd_pi, *q1p, *q2p, *dpi, *meant_pi, *dlogp, *dmeant;
};
template <typename A>
static void carve_sac_qnet(Carver& cv, const Dims& d, const Dims& da, const qnet::Geo& g, SacQWs<A>& w) {
  carve_actor<A>(cv, da, true, w.actor_s);
  carve_actor<A>(cv, d, false, w.actor_tmp);
  qnet::carve(cv, g, w.cq);
  qnet::carve(cv, g, w.ct);
  const int64_t bn = (int64_t)d.B * d.na, an = (int64_t)da.B * d.na;
  w.a2 = cv.take<float>(bn); w.logp2 = cv.take<float>(d.B);
  w.mean_tmp = cv.take<float>(bn); w.lstd_tmp = cv.take<float>(bn);
  w.q1t = cv.take<float>(bn); w.q2t = cv.take<float>(bn);
  w.q1 = cv.take<float>(bn); w.q2 = cv.take<float>(bn);
  w.dq1 = cv.take<float>(bn); w.dq2 = cv.take<float>(bn); w.nq = cv.take<float>(bn);
  w.pi = cv.take<float>(an); w.logpi = cv.take<float>(da.B);
  w.mean_pi = cv.take<float>(an); w.lstd_pi = cv.take<float>(an);
  w.q1p = cv.take<float>(bn); w.q2p = cv.take<float>(bn); w.dpi = cv.take<float>(an);
  w.meant_pi = cv.take<float>(an); w.dlogp = cv.take<float>(da.B); w.dmeant = cv.take<float>(an);
}
// the flat QNetwork arena seen by adam_step / the data-parallel exchange: one range, nothing skipped, no 16-bit shadows
static dgvit_layout qnet_flat_layout(const dgvit_qnet_layout& Lq) {
  dgvit_layout L;
  memset(&L, 0, sizeof(L));
  L.total = Lq.total;
  return L;
}

template <typename A>
static void sac_qnet_run(const dgvit_sac& s, const dgvit_batch* bp, const dgvit_noise* nz, const dgvit_sac_out* outp,
                         const Dims& d, const Dims& da, const qnet::Geo& g, SacQWs<A>& w, cudaStream_t st, int phases) {
  dgvit_layout La;
  make_layout(s.actor.cfg, La);
  dgvit_qnet_layout Lq;
  qnet::make_layout(g.na, g.nps, Lq);
  const dgvit_layout Lc = qnet_flat_layout(Lq);
  const bool shadow = s.precision == DGVIT_BF16;
  dgvit_actor_io ap; memset(&ap, 0, sizeof(ap));       // pi, log_pi = policy.sample(s): B rows (+ the imitation rows)
  if (phases & 3) {
    ap.img = bp->obs; ap.pstate = bp->pobs; ap.eps = nz ? nz->eps_pi : nullptr;
    ap.action_scale = s.action_scale; ap.action_bias = s.action_bias;
    ap.drop = sac_drop(s, nz, nz ? nz->mask_a : nullptr, 4);
    ap.sample_offset = s.sample_offset;
    ap.mean = w.mean_pi; ap.log_std = w.lstd_pi; ap.action = w.pi; ap.log_prob = w.logpi; ap.mean_t = w.meant_pi;
  }
  if (phases & 1) {
    const dgvit_batch& b = *bp;
    const dgvit_sac_out& out = *outp;
    ForkState f = fork_state();
    if (!g_fork_enabled) { f.aux[0] = st; f.aux[1] = st; }
    DG_CUDA(cudaEventRecord(f.fork, st));
    DG_CUDA(cudaStreamWaitEvent(f.aux[0], f.fork, 0));
    DG_CUDA(cudaStreamWaitEvent(f.aux[1], f.fork, 0));
    // ---- forked stream 0: critic(s, a)                                                (DRL.py:395)
    qnet::forward<A>(s.critic.params, Lq, g, b.obs, b.pobs, b.act, w.q1, w.q2, w.cq, f.aux[0]);
    DG_CUDA(cudaEventRecord(f.join[0], f.aux[0]));
    // ---- forked stream 1: policy.sample(s) (the actor is not touched before DRL.py:413)
    actor_forward<A>(s.actor, La, da, ap, w.actor_s, f.aux[1]);
    DG_CUDA(cudaEventRecord(f.join[1], f.aux[1]));
    // ---- caller's stream: a', log pi' = policy.sample(s') ; q_t = critic_target(s', a')   (DRL.py:388-393)
    dgvit_actor_io ai; memset(&ai, 0, sizeof(ai));
    ai.img = b.next_obs; ai.pstate = b.next_pobs; ai.eps = nz ? nz->eps_next : nullptr;
    ai.action_scale = s.action_scale; ai.action_bias = s.action_bias;
    ai.drop = sac_drop(s, nz, nz ? nz->mask_a_next : nullptr, 1);
    ai.sample_offset = s.sample_offset;
    ai.mean = w.mean_tmp; ai.log_std = w.lstd_tmp; ai.action = w.a2; ai.log_prob = w.logp2;
    actor_forward<A>(s.actor, La, d, ai, w.actor_tmp, st);
    qnet::forward<A>(s.critic_target.params, Lq, g, b.next_obs, b.next_pobs, w.a2, w.q1t, w.q2t, w.ct, st);
    // ---- losses + critic backward                                                     (DRL.py:396-401)
    DG_CUDA(cudaStreamWaitEvent(st, f.join[0], 0));
    launch_k(critic_loss_kernel, 1, 1024, 0, st, w.q1, w.q2, w.q1t, w.q2t, w.logp2, b.rew, s.alpha, s.gamma, d.B, d.na,
             s.global_batch, w.nq, w.dq1, w.dq2, out.losses);
    DG_LAUNCH_CHECK();
    dp_wait_done(s.dp, 0, Lc, s.critic_opt.step, st);
    qnet::backward<A>(s.critic.params, s.critic.grads, Lq, g, b.obs, b.pobs, w.dq1, w.dq2, nullptr, true, w.cq, st);
    if (out.debug) {
      const size_t bn = (size_t)d.B * d.na * sizeof(float);
      DG_CUDA(cudaMemcpyAsync(out.debug, w.nq, bn, cudaMemcpyDeviceToDevice, st));
      DG_CUDA(cudaMemcpyAsync(out.debug + d.B * d.na, w.q1, bn, cudaMemcpyDeviceToDevice, st));
      DG_CUDA(cudaMemcpyAsync(out.debug + 2 * d.B * d.na, w.q2, bn, cudaMemcpyDeviceToDevice, st));
    }
    DG_CUDA(cudaStreamWaitEvent(st, f.join[1], 0));
  }
  if (phases & 2) {
    const dgvit_batch& b = *bp;
    const dgvit_sac_out& out = *outp;
    dgvit_net cnet = s.critic;
    cnet.shadow = nullptr;
    adam_step(cnet, Lc, s.critic_opt, nullptr, 0.f, false, st, s.dp, 0);                  // DRL.py:402
    cudaStream_t lane = st;
    if (s.do_polyak) {                                                                    // DRL.py:430-431, beside the policy half
      lane = lane_fork(st);
      launch_k(polyak_kernel, 148 * 4, 256, 0, lane, s.critic_target.params, (const float*)s.critic.params, (bf16*)nullptr, s.tau,
               Lc.total);
      DG_LAUNCH_CHECK();
    }
    // ---- q_pi = critic(s, pi) with the UPDATED critic                                 (DRL.py:406-407)
    qnet::forward<A>(s.critic.params, Lq, g, b.obs, b.pobs, w.pi, w.q1p, w.q2p, w.cq, st);
    launch_k(policy_loss_kernel, 1, 1024, 0, st, w.q1p, w.q2p, w.logpi, s.alpha, s.log_alpha, s.target_entropy, d.B, d.na,
             s.global_batch, w.dq1, w.dq2, out.losses, s.actor.grads + La.alpha_grad_slot);
    DG_LAUNCH_CHECK();
    // d policy_loss / d pi through the critic heads only (its parameter gradients are discarded, DRL.py:399)
    qnet::backward<A>(s.critic.params, nullptr, Lq, g, b.obs, b.pobs, w.dq1, w.dq2, w.dpi, false, w.cq, st);
    dgvit_actor_grad ag; memset(&ag, 0, sizeof(ag));
    ag.d_action = w.dpi;
    dp_wait_done(s.dp, 1, La, s.actor_opt.step, st);
    if (da.B == d.B) {
      actor_backward<A>(s.actor, La, d, ap, ag, s.alpha, 1.0f / (float)s.global_batch, w.actor_s, st);
    } else {
      DG_REQUIRE(b.extra_target && b.extra_weight, "n_extra > 0 needs extra_target / extra_weight");
      launch_k(imitation_grad_kernel, 1, 1024, 0, st, (const float*)w.meant_pi, b.extra_target, b.extra_weight, (const float*)s.alpha,
               1.0f / (float)s.global_batch, d.B, da.B, d.na, w.dpi, w.dlogp, w.dmeant, out.losses);
      DG_LAUNCH_CHECK();
      ag.d_log_prob = w.dlogp;
      ag.d_mean_t = w.dmeant;
      actor_backward<A>(s.actor, La, da, ap, ag, nullptr, 0.f, w.actor_s, st);
    }
    if (out.debug) {
      const size_t bn = (size_t)d.B * d.na * sizeof(float);
      DG_CUDA(cudaMemcpyAsync(out.debug + 3 * d.B * d.na, w.pi, bn, cudaMemcpyDeviceToDevice, st));
      DG_CUDA(cudaMemcpyAsync(out.debug + 4 * d.B * d.na, w.q1p, bn, cudaMemcpyDeviceToDevice, st));
      DG_CUDA(cudaMemcpyAsync(out.debug + 5 * d.B * d.na, w.logpi, (size_t)d.B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (lane != st) lane_join(lane, st);
  }
  if (phases & 4) sac_phase3<A>(s, st);        // actor Adam, temperature, RNG counter: nothing critic-specific
  (void)shadow;
}

static void check_sac(const dgvit_sac& s, int B) {
  DG_REQUIRE(B >= 1, "B must be >= 1");
  DG_REQUIRE(s.actor.cfg.kind == DGVIT_ACTOR && (s.critic.cfg.kind == DGVIT_CRITIC || s.critic.cfg.kind == DGVIT_QNET) &&
             s.critic_target.cfg.kind == s.critic.cfg.kind, "sac: wrong network kinds");
  DG_REQUIRE(s.actor.params && s.actor.grads && s.critic.params && s.critic.grads && s.critic_target.params,
             "sac: null arena");
  DG_REQUIRE(s.alpha && s.log_alpha, "sac: null alpha");
  DG_REQUIRE(s.global_batch >= B, "sac: global_batch < B");
  DG_REQUIRE(s.action_scale && s.action_bias, "sac: null action scale/bias");
  if (s.precision == DGVIT_BF16)
    DG_REQUIRE(s.actor.shadow && (s.critic.cfg.kind == DGVIT_QNET || (s.critic.shadow && s.critic_target.shadow)),
               "sac: bf16 needs shadow arenas");
}

template <typename A>
static size_t sac_ws_bytes(const dgvit_cfg& acfg, int B, int n_extra) {
  Dims d(acfg, B), da(acfg, B + n_extra);
  Carver cv(nullptr, 0, true);
  SacWs<A> w;
  carve_sac<A>(cv, d, da, w);
  return cv.off;
}

// behaviour-cloning step workspace (dgvit_bc_step)
template <typename A>
struct BcWs {
  ActorCtx<A> actor;
  float *mean, *lstd, *action, *logp, *mean_t, *d_mean_t, *part, *scale;
};
template <typename A>
static void carve_bc(Carver& cv, const Dims& d, BcWs<A>& w) {
  carve_actor<A>(cv, d, true, w.actor);
  const int64_t bn = (int64_t)d.B * d.na;
  w.mean = cv.take<float>(bn); w.lstd = cv.take<float>(bn); w.action = cv.take<float>(bn); w.logp = cv.take<float>(d.B);
  w.mean_t = cv.take<float>(bn); w.d_mean_t = cv.take<float>(bn);
  w.part = cv.take<float>(GN_BLOCKS); w.scale = cv.take<float>(4);
}
template <typename F32, typename BF>
static void by_precision(int precision, F32&& f32, BF&& bf) {
  if (precision == DGVIT_FP32) f32();
  else if (precision == DGVIT_BF16) bf();
  else fail(DGVIT_ERR_ARG, "unknown precision %d", precision);
}

}  // namespace dgvit

// =====================================================================================
// C ABI
// =====================================================================================
using namespace dgvit;

extern "C" int g_depth_strip, g_depth_skip;      // depth.cu

extern "C" {

int dgvit_version(void) { return 100; }
long long dgvit_launch_count(void) { return launch_counter(); }

int dgvit_set_option(const char* name, int value) {
  return guarded([&] {
    DG_REQUIRE(name != nullptr, "null option name");
    if (!strcmp(name, "fork_streams")) g_fork_enabled = value != 0;
    else if (!strcmp(name, "bwd_side")) g_side_enabled = value != 0;
    else if (!strcmp(name, "actor_s_when")) g_actor_s_when = value >= 0 && value <= 2 ? value : 0;
    else if (!strcmp(name, "target_fork")) g_target_fork = value != 0;
    else if (!strcmp(name, "ln_bwd_warps")) g_lnb_wpb = value >= 1 && value <= 16 ? value : 16;
    else if (!strcmp(name, "ln_bwd_blocks_per_sm")) g_lnb_bps = value >= 1 && value <= 4 ? value : 2;
    else if (!strcmp(name, "pdl")) pdl_enabled() = value != 0;
    else if (!strcmp(name, "skip")) skip_mask() = value;
    else if (!strcmp(name, "attention_row0")) g_row0_mode = value;
    else if (!strcmp(name, "depth_strip")) g_depth_strip = value;
    else if (!strcmp(name, "depth_skip")) g_depth_skip = value;
#ifdef DGVIT_WITH_TC
    else if (!strcmp(name, "tensor_cores")) tc::g_tc_enabled = value != 0;
    else if (!strcmp(name, "debug_epilogue")) tc::g_debug = value;
    else if (!strcmp(name, "attn_bwd2")) attn::g_bwd2_enabled = value != 0;
    else if (!strcmp(name, "attn_long")) attnl::g_enabled = value != 0;
    else if (!strcmp(name, "patch_fused")) patch::g_enabled = value != 0;
    else if (!strcmp(name, "patch_dw")) patch::g_dw_enabled = value != 0;
    else if (!strcmp(name, "mlp_split")) mlp::g_split_enabled = value != 0;
    else if (!strcmp(name, "mlp_front")) mlp::g_front_enabled = value != 0;
    else if (!strcmp(name, "mlp_h16")) mlp::g_h16_enabled = value != 0;
#endif
    else fail(DGVIT_ERR_ARG, "unknown option %s", name);
  });
}

#if defined(DGVIT_MLP_TRACE) && defined(DGVIT_WITH_TC)
// trace builds only (make trace -> libdgvit_trace.so): device buffer receiving the pipeline timeline of CTA 0
int dgvit_debug_set_trace(void* p) { mlp::g_trace = (long long*)p; return 0; }
#endif

int dgvit_prof_begin(int tag, int max_launches) {
  return guarded([&] {
    Prof& p = prof();
    DG_REQUIRE(max_launches > 0 && max_launches <= (1 << 20), "bad max_launches");
    while (p.ev.size() < (size_t)2 * max_launches) {
      cudaEvent_t e;
      DG_CUDA(cudaEventCreate(&e));
      p.ev.push_back(e);
    }
    p.used = 0; p.flops = 0; p.bytes = 0; p.launches = 0; p.tag = tag; p.on = true;
  });
}
int dgvit_prof_end(double* ms_total, long long* launches, double* flops, double* bytes) {
  return guarded([&] {
    Prof& p = prof();
    p.on = false;
    double ms = 0;
    for (size_t i = 0; i + 1 < p.used; i += 2) {
      DG_CUDA(cudaEventSynchronize(p.ev[i + 1]));
      float t = 0;
      DG_CUDA(cudaEventElapsedTime(&t, p.ev[i], p.ev[i + 1]));
      ms += t;
    }
    if (ms_total) *ms_total = ms;
    if (launches) *launches = p.launches;
    if (flops) *flops = p.flops;
    if (bytes) *bytes = p.bytes;
  });
}
const char* dgvit_last_error(void) { return last_error().c_str(); }

int dgvit_param_layout(const dgvit_cfg* cfg, dgvit_layout* out) {
  return guarded([&] {
    DG_REQUIRE(cfg && out, "null argument");
    make_layout(*cfg, *out);
  });
}

int dgvit_workspace_bytes(const dgvit_cfg* cfg, int B, int precision, int save, size_t* bytes) {
  return guarded([&] {
    DG_REQUIRE(cfg && bytes && B >= 1, "bad argument");
    Dims d(*cfg, B);
    Carver cv(nullptr, 0, true);
    auto run = [&](auto tag) {
      using A = decltype(tag);
      if (cfg->kind == DGVIT_ACTOR) { ActorCtx<A> c; carve_actor<A>(cv, d, save != 0, c); }
      else { CriticCtx<A> c; carve_critic<A>(cv, d, save != 0, c); }
    };
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
    *bytes = cv.off;
  });
}

int dgvit_sac_workspace_bytes(const dgvit_cfg* acfg, int B, int n_extra, int precision, size_t* bytes) {
  return guarded([&] {
    DG_REQUIRE(acfg && bytes && B >= 1 && n_extra >= 0, "bad argument");
    by_precision(precision, [&] { *bytes = sac_ws_bytes<float>(*acfg, B, n_extra); },
                 [&] { *bytes = sac_ws_bytes<bf16>(*acfg, B, n_extra); });
  });
}

int dgvit_sac_qnet_workspace_bytes(const dgvit_cfg* acfg, int B, int n_extra, int precision, size_t* bytes) {
  return guarded([&] {
    DG_REQUIRE(acfg && bytes && B >= 1 && n_extra >= 0, "bad argument");
    Dims d(*acfg, B), da(*acfg, B + n_extra);
    qnet::Geo g(B, acfg->img_h, acfg->img_w, acfg->n_act, acfg->n_pstate);
    DG_REQUIRE(g.H3 >= 1 && g.W3 >= 1, "qnet: image too small");
    Carver cv(nullptr, 0, true);
    by_precision(precision, [&] { SacQWs<float> w; carve_sac_qnet<float>(cv, d, da, g, w); },
                 [&] { SacQWs<bf16> w; carve_sac_qnet<bf16>(cv, d, da, g, w); });
    *bytes = cv.off;
  });
}

int dgvit_refresh_shadow(const dgvit_net* net, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(net ? net->params : nullptr);
    DG_REQUIRE(net && net->params && net->shadow, "null argument");
    dgvit_layout L;
    make_layout(net->cfg, L);
    launch_k(shadow_refresh_kernel, 148 * 4, 256, 0, (cudaStream_t)stream, net->params, (bf16*)net->shadow, L.total);
    DG_LAUNCH_CHECK();
  });
}

int dgvit_actor_forward(const dgvit_net* net, const dgvit_actor_io* io, int B, int precision, int save,
                        void* ws, size_t ws_bytes, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(net ? net->params : nullptr);
    DG_REQUIRE(net && io && ws && B >= 1 && net->cfg.kind == DGVIT_ACTOR && net->params, "bad argument");
    dgvit_layout L;
    make_layout(net->cfg, L);
    Dims d(net->cfg, B);
    auto run = [&](auto tag) {
      using A = decltype(tag);
      Carver cv(ws, ws_bytes);
      ActorCtx<A> c;
      carve_actor<A>(cv, d, save != 0, c);
      actor_forward<A>(*net, L, d, *io, c, (cudaStream_t)stream);
    };
    if (precision == DGVIT_BF16) DG_REQUIRE(net->shadow, "bf16 needs a shadow arena");
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}

int dgvit_actor_backward(const dgvit_net* net, const dgvit_actor_io* io, const dgvit_actor_grad* g, int B,
                         int precision, void* ws, size_t ws_bytes, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(net ? net->params : nullptr);
    DG_REQUIRE(net && io && g && ws && B >= 1 && net->cfg.kind == DGVIT_ACTOR && net->params && net->grads,
               "bad argument");
    dgvit_layout L;
    make_layout(net->cfg, L);
    Dims d(net->cfg, B);
    auto run = [&](auto tag) {
      using A = decltype(tag);
      Carver cv(ws, ws_bytes);
      ActorCtx<A> c;
      carve_actor<A>(cv, d, true, c);
      actor_backward<A>(*net, L, d, *io, *g, nullptr, 0.f, c, (cudaStream_t)stream);
    };
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}

// ---- behaviour-cloning step (vn/attention_imitating.py:48-67): one call = policy.sample -> RMSE on the clipped tanh-mean
//      -> backward -> clip_grad_norm_ -> Adam
int dgvit_bc_workspace_bytes(const dgvit_cfg* cfg, int B, int precision, size_t* bytes) {
  return guarded([&] {
    DG_REQUIRE(cfg && bytes && B >= 1 && cfg->kind == DGVIT_ACTOR, "bad argument");
    Dims d(*cfg, B);
    Carver cv(nullptr, 0, true);
    by_precision(precision, [&] { BcWs<float> w; carve_bc<float>(cv, d, w); }, [&] { BcWs<bf16> w; carve_bc<bf16>(cv, d, w); });
    *bytes = cv.off;
  });
}
int dgvit_bc_step(const dgvit_net* net, const dgvit_adam* opt, const dgvit_bc_io* io, int B, int precision, void* ws,
                  size_t ws_bytes, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(net ? net->params : nullptr);
    DG_REQUIRE(net && opt && io && ws && B >= 1 && net->cfg.kind == DGVIT_ACTOR && net->params && net->grads, "bad argument");
    DG_REQUIRE(io->img && io->pstate && io->target && io->action_scale && io->action_bias && io->loss, "bc_step: null input");
    DG_REQUIRE(io->max_action > 0.f && io->max_norm > 0.f, "bc_step: max_action and max_norm must be positive");
    if (precision == DGVIT_BF16) DG_REQUIRE(net->shadow, "bc_step: bf16 needs the shadow arena");
    dgvit_layout L;
    make_layout(net->cfg, L);
    Dims d(net->cfg, B);
    cudaStream_t st = (cudaStream_t)stream;
    auto run = [&](auto tag) {
      using A = decltype(tag);
      Carver cv(ws, ws_bytes);
      BcWs<A> w;
      carve_bc<A>(cv, d, w);
      dgvit_actor_io ai; memset(&ai, 0, sizeof(ai));
      ai.img = io->img; ai.pstate = io->pstate; ai.eps = io->eps;
      ai.action_scale = io->action_scale; ai.action_bias = io->action_bias;
      ai.drop = io->drop; ai.sample_offset = io->sample_offset;
      ai.mean = w.mean; ai.log_std = w.lstd; ai.action = w.action; ai.log_prob = w.logp; ai.mean_t = w.mean_t;
      actor_forward<A>(*net, L, d, ai, w.actor, st);                                         // :56
      const int n = d.B * d.na;
      launch_k(bc_loss_kernel, 1, 1024, 0, st, (const float*)w.mean_t, io->target, n, io->max_action, w.d_mean_t, io->loss);   // :58-60
      DG_LAUNCH_CHECK();
      dgvit_actor_grad ag; memset(&ag, 0, sizeof(ag));
      ag.d_mean_t = w.d_mean_t;
      actor_backward<A>(*net, L, d, ai, ag, nullptr, 0.f, w.actor, st);                        // :61
      AdamArgs a;
      fill_adam_args(a, *net, L, *opt);
      launch_k(sumsq_partial_kernel, GN_BLOCKS, 256, 0, st, a, w.part);                        // :62 clip_grad_norm_
      DG_LAUNCH_CHECK();
      launch_k(clip_scale_kernel, 1, 32, 0, st, (const float*)w.part, GN_BLOCKS, io->max_norm, w.scale, io->grad_norm);
      DG_LAUNCH_CHECK();
      adam_step(*net, L, *opt, nullptr, 0.f, precision == DGVIT_BF16, st, nullptr, 0, w.scale);   // :63
      if (io->advance_rng && io->drop.rng_state) {
        launch_k(rng_advance_kernel, 1, 32, 0, st, const_cast<uint64_t*>(io->drop.rng_state));
        DG_LAUNCH_CHECK();
      }
    };
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}

int dgvit_critic_forward(const dgvit_net* net, const dgvit_critic_io* io, int B, int precision, int save,
                         void* ws, size_t ws_bytes, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(net ? net->params : nullptr);
    DG_REQUIRE(net && io && ws && B >= 1 && net->cfg.kind == DGVIT_CRITIC && net->params, "bad argument");
    dgvit_layout L;
    make_layout(net->cfg, L);
    Dims d(net->cfg, B);
    auto run = [&](auto tag) {
      using A = decltype(tag);
      Carver cv(ws, ws_bytes);
      CriticCtx<A> c;
      carve_critic<A>(cv, d, save != 0, c);
      critic_forward<A>(*net, L, d, *io, 0, c, (cudaStream_t)stream);
    };
    if (precision == DGVIT_BF16) DG_REQUIRE(net->shadow, "bf16 needs a shadow arena");
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}

int dgvit_critic_backward(const dgvit_net* net, const dgvit_critic_io* io, const float* d_q1, const float* d_q2,
                          float* d_action, int param_grads, int B, int precision, void* ws, size_t ws_bytes,
                          void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(net ? net->params : nullptr);
    DG_REQUIRE(net && io && d_q1 && d_q2 && ws && B >= 1 && net->cfg.kind == DGVIT_CRITIC && net->params,
               "bad argument");
    if (param_grads) DG_REQUIRE(net->grads, "null grads");
    dgvit_layout L;
    make_layout(net->cfg, L);
    Dims d(net->cfg, B);
    auto run = [&](auto tag) {
      using A = decltype(tag);
      Carver cv(ws, ws_bytes);
      CriticCtx<A> c;
      carve_critic<A>(cv, d, true, c);
      critic_backward<A>(*net, L, d, *io, 0, d_q1, d_q2, d_action, param_grads != 0, c, nullptr,
                         (cudaStream_t)stream);
    };
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}

// GoT.forward(img, goal) -> z (vn/GoalFormer.py:156-171) on its own: patch embedding, goal token prepended, position
// embedding, dropout, the transformer blocks, token-0 pooling, RMSNorm.  `net` may be an actor or a critic arena (the
// trunk sits at the same offsets in both).
int dgvit_trunk_workspace_bytes(const dgvit_cfg* cfg, int B, int precision, int save, size_t* bytes) {
  return guarded([&] {
    DG_REQUIRE(cfg && bytes && B >= 1, "bad argument");
    Dims d(*cfg, B);
    Carver cv(nullptr, 0, true);
    by_precision(precision, [&] { TrunkCtx<float> c; carve_trunk<float>(cv, d, save != 0, c); },
                 [&] { TrunkCtx<bf16> c; carve_trunk<bf16>(cv, d, save != 0, c); });
    *bytes = cv.off;
  });
}

int dgvit_trunk_forward(const dgvit_net* net, const dgvit_trunk_io* io, int B, int precision, int save, void* ws,
                        size_t ws_bytes, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(net ? net->params : nullptr);
    DG_REQUIRE(net && io && ws && B >= 1 && net->params && io->img && io->goal && io->z, "bad argument");
    dgvit_layout L;
    make_layout(net->cfg, L);
    Dims d(net->cfg, B);
    cudaStream_t st = (cudaStream_t)stream;
    auto run = [&](auto tag) {
      using A = decltype(tag);
      Carver cv(ws, ws_bytes);
      TrunkCtx<A> c;
      carve_trunk<A>(cv, d, save != 0, c);
      const DropDev drop = make_drop(io->drop, d, io->sample_offset);
      GoalTok gt{nullptr, nullptr, nullptr, 0, 0};
      gt.direct = io->goal;
      trunk_forward<A>(*net, L, d, io->img, drop, c, st, gt, /*fuse_rms=*/false);
      DG_CUDA(cudaMemcpyAsync(io->z, c.z, (size_t)B * d.D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    };
    if (precision == DGVIT_BF16) DG_REQUIRE(net->shadow, "bf16 needs a shadow arena");
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}

int dgvit_trunk_backward(const dgvit_net* net, const dgvit_trunk_io* io, const float* d_z, float* d_goal, int B,
                         int precision, void* ws, size_t ws_bytes, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(net ? net->params : nullptr);
    DG_REQUIRE(net && io && d_z && ws && B >= 1 && net->params && net->grads, "bad argument");
    dgvit_layout L;
    make_layout(net->cfg, L);
    Dims d(net->cfg, B);
    cudaStream_t st = (cudaStream_t)stream;
    auto run = [&](auto tag) {
      using A = decltype(tag);
      Carver cv(ws, ws_bytes);
      TrunkCtx<A> c;
      carve_trunk<A>(cv, d, true, c);
      DG_CUDA(cudaMemcpyAsync(c.dz, d_z, (size_t)B * d.D * sizeof(float), cudaMemcpyDeviceToDevice, st));
      const DropDev drop = make_drop(io->drop, d, io->sample_offset);
      trunk_backward<A>(*net, L, d, drop, c, /*relu_tok=*/0, st, io->img);
      if (d_goal) DG_CUDA(cudaMemcpyAsync(d_goal, c.dtok, (size_t)B * d.D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    };
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}

static void qnet_check(int img_h, int img_w, int n_act, int n_pstate, int B);
static int sac_run(const dgvit_sac* s, const dgvit_batch* b, const dgvit_noise* nz, const dgvit_sac_out* out, int B,
                   void* ws, size_t ws_bytes, void* stream, int phases) {
  return guarded([&] {
    DeviceGuard dev_guard(s ? s->actor.params : nullptr);
    DG_REQUIRE(s && ws, "null argument");
    check_sac(*s, B);
    if (phases & 3) {
      DG_REQUIRE(b && out && out->losses, "null batch/out");
      DG_REQUIRE(b->obs && b->next_obs && b->pobs && b->next_pobs && b->act && b->rew, "null batch tensor");
      if (nz && nz->drop_mode == DGVIT_DROP_MASK) {
        DG_REQUIRE(nz->mask_a_next && nz->mask_a, "null mask");
        if (s->critic.cfg.kind != DGVIT_QNET) DG_REQUIRE(nz->mask_ct && nz->mask_c && nz->mask_c_pi, "null mask");
      }
      if (!nz || !nz->eps_next || !nz->eps_pi || nz->drop_mode == DGVIT_DROP_RNG)
        DG_REQUIRE(s->rng_state, "rng_state required when noise is not injected");
    }
    DG_REQUIRE(s->n_extra >= 0, "n_extra < 0");
    Dims d(s->actor.cfg, B), da(s->actor.cfg, B + s->n_extra);
    cudaStream_t st = (cudaStream_t)stream;
    if (s->critic.cfg.kind == DGVIT_QNET) {       // CNN twin-Q critic (vn/DRL.py:118-121)
      const dgvit_cfg& cc = s->critic.cfg;
      DG_REQUIRE(cc.img_h == s->actor.cfg.img_h && cc.img_w == s->actor.cfg.img_w && cc.n_act == s->actor.cfg.n_act &&
                 cc.n_pstate == s->actor.cfg.n_pstate, "sac: actor / critic disagree on the input shapes");
      qnet_check(cc.img_h, cc.img_w, cc.n_act, cc.n_pstate, B);
      qnet::Geo g(B, cc.img_h, cc.img_w, cc.n_act, cc.n_pstate);
      auto runq = [&](auto tag) {
        using A = decltype(tag);
        Carver cv(ws, ws_bytes);
        SacQWs<A> w;
        carve_sac_qnet<A>(cv, d, da, g, w);
        sac_qnet_run<A>(*s, b, nz, out, d, da, g, w, st, phases);
      };
      by_precision(s->precision, [&] { runq(float()); }, [&] { runq(bf16()); });
      return;
    }
    auto run = [&](auto tag) {
      using A = decltype(tag);
      Carver cv(ws, ws_bytes);
      SacWs<A> w;
      carve_sac<A>(cv, d, da, w);
      if (phases & 1) sac_phase1<A>(*s, *b, nz, *out, d, da, w, st);
      if (phases & 2) sac_phase2<A>(*s, *b, nz, *out, d, da, w, st);
      if (phases & 4) sac_phase3<A>(*s, st);
    };
    by_precision(s->precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}

int dgvit_sac_phase1(const dgvit_sac* s, const dgvit_batch* b, const dgvit_noise* nz, const dgvit_sac_out* out,
                     int B, void* ws, size_t ws_bytes, void* stream) {
  return sac_run(s, b, nz, out, B, ws, ws_bytes, stream, 1);
}
int dgvit_sac_phase2(const dgvit_sac* s, const dgvit_batch* b, const dgvit_noise* nz, const dgvit_sac_out* out,
                     int B, void* ws, size_t ws_bytes, void* stream) {
  return sac_run(s, b, nz, out, B, ws, ws_bytes, stream, 2);
}
int dgvit_sac_phase3(const dgvit_sac* s, int B, void* ws, size_t ws_bytes, void* stream) {
  return sac_run(s, nullptr, nullptr, nullptr, B, ws, ws_bytes, stream, 4);
}
int dgvit_sac_update(const dgvit_sac* s, const dgvit_batch* b, const dgvit_noise* nz, const dgvit_sac_out* out,
                     int B, void* ws, size_t ws_bytes, void* stream) {
  return sac_run(s, b, nz, out, B, ws, ws_bytes, stream, 7);
}

int dgvit_gemm_bf16(int M, int N, int K, const void* A, int64_t a_sm, int64_t a_sk, const void* B, int64_t b_sk,
                    int64_t b_sn, float* C, int64_t ldc, int splitk, float* partial, int use_tensor_cores,
                    void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(A);
    DG_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "bad argument");
    GemmArgs g;
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.a_sm = a_sm; g.a_sk = a_sk;
    g.B = B; g.b_sk = b_sk; g.b_sn = b_sn;
    g.C = C; g.ldc = ldc; g.splitk = splitk < 1 ? 1 : splitk; g.partial = partial;
    if (use_tensor_cores) {
#ifdef DGVIT_WITH_TC
      if (!gemm_tc_try<bf16, bf16, float>(g, (cudaStream_t)stream))
        fail(DGVIT_ERR_ARG, "gemm_bf16: shape/layout not eligible for the tensor-core kernel");
#else
      fail(DGVIT_ERR_ARG, "built without the tensor-core kernel");
#endif
    } else {
      gemm_simt<bf16, bf16, float>(g, (cudaStream_t)stream);
    }
  });
}

int dgvit_linear_bf16(const void* x, const void* W, void* y, int64_t rows, int N, int K, int epilogue, const float* bias,
                      const void* aux, void* y2, int weight_is_kn, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(x);
    DG_REQUIRE(x && W && y && rows > 0 && N > 0 && K > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    GemmArgs g;
    g.M = (int)rows; g.N = N; g.K = K;
    g.A = x; g.a_sm = K; g.a_sk = 1;
    g.B = W;
    if (weight_is_kn) { g.b_sk = N; g.b_sn = 1; } else { g.b_sk = 1; g.b_sn = K; }
    g.C = y; g.ldc = N; g.C2 = y2; g.bias = bias; g.aux = aux; g.ldaux = N;
    switch (epilogue) {
      case 0: g.epi = EPI_NONE; break;
      case 1: g.epi = EPI_BIAS_GELU2; DG_REQUIRE(bias && y2, "gelu2 needs bias and y2"); break;
      case 2: g.epi = EPI_GELU_BWD; DG_REQUIRE(aux, "gelu_bwd needs aux"); break;
      case 3: g.epi = EPI_GELU_BWD2; DG_REQUIRE(aux && y2, "gelu_bwd2 needs aux and y2"); break;
      default: fail(DGVIT_ERR_ARG, "unknown epilogue %d", epilogue);
    }
    gemm<bf16, bf16, bf16>(g, st);
  });
}

int64_t dgvit_mlp_partial_floats(int64_t rows, int hid) {
#ifdef DGVIT_WITH_TC
  if (rows >= 1 && hid >= mlp::HC && hid % mlp::HC == 0) return (int64_t)mlp::bwd_partial_floats(rows, hid);
#endif
  return -1;
}

int dgvit_mlp_bf16(const void* x, const void* W1, const float* b1, const void* W2, const float* b2, const float* resid,
                   float* out, const void* d_y, float* d_x, float* d_w, float* d_b2, float* partial, int64_t rows, int hid,
                   void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(x);
#ifdef DGVIT_WITH_TC
    DG_REQUIRE(x && W1 && b1 && W2 && rows > 0 && hid > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_y) {
      DG_REQUIRE(b2 && resid && out, "forward needs b2, resid, out");
      DG_REQUIRE(mlp::eligible(64, hid, rows, x, W1, W2, resid, 64, out, 64), "mlp: shape not eligible for the fused kernel");
      mlp::fwd((const bf16*)x, (const bf16*)W1, b1, (const bf16*)W2, b2, resid, 64, out, 64, rows, hid, st);
    } else {
      DG_REQUIRE(d_x && d_w && d_b2 && partial, "backward needs d_x, d_w, d_b2, partial");
      DG_REQUIRE(mlp::eligible(64, hid, rows, x, W1, W2, d_x, 64, d_x, 64) && ((uintptr_t)d_y & 15) == 0,
                 "mlp: shape not eligible for the fused kernel");
      mlp::bwd((const bf16*)x, (const bf16*)d_y, (const bf16*)W1, b1, (const bf16*)W2, d_x, d_w, d_w + (int64_t)hid * 64,
               d_w + (int64_t)hid * 65, d_b2, partial, rows, hid, st);
    }
#else
    fail(DGVIT_ERR_ARG, "built without the tensor-core kernels");
#endif
  });
}

int dgvit_mlp_fwd_f16w2(const void* x, const void* W1, const float* b1, const void* W2_f16, const float* b2, const float* resid,
                        float* out, int64_t rows, int hid, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(x);
#ifdef DGVIT_WITH_TC
    DG_REQUIRE(x && W1 && b1 && W2_f16 && b2 && resid && out && rows > 0 && hid > 0, "bad argument");
    DG_REQUIRE(mlp::eligible(64, hid, rows, x, W1, W2_f16, resid, 64, out, 64), "mlp: shape not eligible for the fused kernel");
    mlp::fwd((const bf16*)x, (const bf16*)W1, b1, nullptr, b2, resid, 64, out, 64, rows, hid, (cudaStream_t)stream, mlp::LnFuse(),
             mlp::FrontFuse(), W2_f16);
#else
    fail(DGVIT_ERR_ARG, "built without the tensor-core kernels");
#endif
  });
}

int64_t dgvit_attention_stats_floats(int B, int N, int H) { return N > 128 ? (int64_t)2 * B * H * N : 0; }

int dgvit_attention_bf16(const void* qkv, void* o, const void* d_o, void* d_qkv, int B, int N, int H, int dim_head,
                         int use_tensor_cores, float* stats, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(qkv);
    DG_REQUIRE(qkv && o && B >= 1 && N >= 1 && H >= 1, "bad argument");
    dgvit_cfg c;
    memset(&c, 0, sizeof(c));
    c.img_h = 16; c.img_w = 20 * (N - 1); c.patch_h = 16; c.patch_w = 20; c.heads = H; c.dim_head = dim_head;
    c.dim = 64; c.depth = 1; c.mlp_dim = 64; c.n_act = 2; c.n_pstate = 2;
    Dims d(c, B);
    DG_REQUIRE(d.N == N, "internal: N");
    cudaStream_t st = (cudaStream_t)stream;
#ifdef DGVIT_WITH_TC
    const bool prev = tc::g_tc_enabled;
    tc::g_tc_enabled = use_tensor_cores != 0;
    if (use_tensor_cores)
      DG_REQUIRE(attn::eligible(N, dim_head, qkv, 3 * d.inner) || (stats && attnl::eligible(N, dim_head, qkv, 3 * d.inner)),
                 "attention: shape not eligible for the tensor-core kernels (N > 128 needs the stats scratch)");
    float* lse = stats;
    float* delta = stats ? stats + (int64_t)B * H * N : nullptr;
#else
    DG_REQUIRE(!use_tensor_cores, "built without the tensor-core kernels");
#endif
    try {
      if (!d_o) launch_attention_fwd<bf16>((const bf16*)qkv, (bf16*)o, d, st, lse);
      else {
        DG_REQUIRE(d_qkv, "null d_qkv");
        launch_attention_bwd<bf16>((const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, (bf16*)d_qkv, d, st, lse, delta);
      }
    } catch (...) {
#ifdef DGVIT_WITH_TC
      tc::g_tc_enabled = prev;
#endif
      throw;
    }
#ifdef DGVIT_WITH_TC
    tc::g_tc_enabled = prev;
#endif
  });
}

// ---- CNN twin-Q critic (vn/got_sac_network.py:125-170)
int dgvit_qnet_param_layout(int n_act, int n_pstate, dgvit_qnet_layout* out) {
  return guarded([&] {
    DG_REQUIRE(out && n_act >= 1 && n_act <= 4 && n_pstate >= 1, "bad argument");
    qnet::make_layout(n_act, n_pstate, *out);
  });
}
static void qnet_check(int img_h, int img_w, int n_act, int n_pstate, int B) {
  DG_REQUIRE(B >= 1 && n_act >= 1 && n_act <= 4 && n_pstate >= 1, "qnet: bad sizes");
  qnet::Geo g(B, img_h, img_w, n_act, n_pstate);
  DG_REQUIRE(g.H3 >= 1 && g.W3 >= 1, "qnet: image %dx%d too small for three 5x5 stride-2 convolutions", img_h, img_w);
}
int dgvit_qnet_workspace_bytes(int img_h, int img_w, int n_act, int n_pstate, int B, int precision, size_t* bytes) {
  return guarded([&] {
    DG_REQUIRE(bytes != nullptr, "null bytes");
    qnet_check(img_h, img_w, n_act, n_pstate, B);
    qnet::Geo g(B, img_h, img_w, n_act, n_pstate);
    Carver cv(nullptr, 0, true);
    by_precision(precision, [&] { qnet::Ws<float> w; qnet::carve(cv, g, w); }, [&] { qnet::Ws<bf16> w; qnet::carve(cv, g, w); });
    *bytes = cv.off;
  });
}
int dgvit_qnet_forward(const float* params, const float* img, const float* pstate, const float* action, float* q1, float* q2,
                       int img_h, int img_w, int n_act, int n_pstate, int B, int precision, void* ws, size_t ws_bytes,
                       void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(params);
    DG_REQUIRE(params && img && pstate && action && q1 && q2 && ws, "qnet_forward: null argument");
    DG_REQUIRE((((uintptr_t)img) & 15) == 0 && (((uintptr_t)params) & 15) == 0, "qnet_forward: 16-byte alignment required");
    qnet_check(img_h, img_w, n_act, n_pstate, B);
    qnet::Geo g(B, img_h, img_w, n_act, n_pstate);
    dgvit_qnet_layout L;
    qnet::make_layout(n_act, n_pstate, L);
    Carver cv(ws, ws_bytes);
    auto run = [&](auto tag) {
      using T = decltype(tag);
      qnet::Ws<T> w;
      qnet::carve(cv, g, w);
      qnet::forward<T>(params, L, g, img, pstate, action, q1, q2, w, (cudaStream_t)stream);
    };
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}
int dgvit_qnet_backward(const float* params, float* grads, const float* img, const float* pstate, const float* d_q1,
                        const float* d_q2, float* d_action, int param_grads, int img_h, int img_w, int n_act, int n_pstate,
                        int B, int precision, void* ws, size_t ws_bytes, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(params);
    DG_REQUIRE(params && img && pstate && d_q1 && d_q2 && ws, "qnet_backward: null argument");
    DG_REQUIRE(!param_grads || grads, "qnet_backward: param_grads without a gradient arena");
    qnet_check(img_h, img_w, n_act, n_pstate, B);
    qnet::Geo g(B, img_h, img_w, n_act, n_pstate);
    dgvit_qnet_layout L;
    qnet::make_layout(n_act, n_pstate, L);
    Carver cv(ws, ws_bytes);
    auto run = [&](auto tag) {
      using T = decltype(tag);
      qnet::Ws<T> w;
      qnet::carve(cv, g, w);
      qnet::backward<T>(params, grads, L, g, img, pstate, d_q1, d_q2, d_action, param_grads != 0, w, (cudaStream_t)stream);
    };
    by_precision(precision, [&] { run(float()); }, [&] { run(bf16()); });
  });
}

int dgvit_adam_step(const dgvit_net* net, const dgvit_adam* opt, const dgvit_net* tgt, float tau, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(net ? net->params : nullptr);
    DG_REQUIRE(net && opt && net->params && net->grads, "null argument");
    dgvit_layout L;
    make_layout(net->cfg, L);
    adam_step(*net, L, *opt, tgt, tau, net->shadow != nullptr, (cudaStream_t)stream);
  });
}

int dgvit_polyak(const dgvit_net* target, const dgvit_net* source, float tau, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(target ? target->params : nullptr);
    DG_REQUIRE(target && source && target->params && source->params, "null argument");
    dgvit_layout L, Ls;
    make_layout(target->cfg, L);
    make_layout(source->cfg, Ls);
    DG_REQUIRE(L.total == Ls.total, "polyak: layouts differ");
    launch_k(polyak_kernel, 148 * 4, 256, 0, (cudaStream_t)stream, target->params, source->params, (bf16*)target->shadow, tau,
                                                             L.total);
    DG_LAUNCH_CHECK();
  });
}

int dgvit_polyak_flat(float* target, const float* source, int64_t n, float tau, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(target);
    DG_REQUIRE(target && source && n >= 0, "null argument");
    if (n == 0) return;
    launch_k(polyak_kernel, 148 * 4, 256, 0, (cudaStream_t)stream, target, source, (bf16*)nullptr, tau, n);
    DG_LAUNCH_CHECK();
  });
}

int dgvit_replay_gather(const dgvit_replay* s, const int64_t* idx, int B, float* obs, float* next_obs, float* pobs,
                        float* next_pobs, float* act, float* rew, float* done, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(s ? s->obs : nullptr);
    DG_REQUIRE(s && idx && B >= 0 && s->obs && s->size > 0, "bad argument");
    DG_REQUIRE(s->frame % 4 == 0, "frame size must be a multiple of 4 floats");
    DG_REQUIRE(((uintptr_t)s->obs % 16) == 0 && ((uintptr_t)obs % 16) == 0 && ((uintptr_t)next_obs % 16) == 0,
               "frame buffers must be 16-byte aligned");
    if (B == 0) return;
    cudaStream_t st = (cudaStream_t)stream;
    GatherArgs g;
    memset(&g, 0, sizeof(g));
    g.store = (const float4*)s->obs; g.idx = idx; g.size = s->size; g.frame4 = s->frame / 4;
    g.obs = (float4*)obs; g.next_obs = (float4*)next_obs;
    g.n = 5;
    g.src[0] = s->pobs; g.dst[0] = pobs; g.width[0] = s->n_pstate;
    g.src[1] = s->next_pobs; g.dst[1] = next_pobs; g.width[1] = s->n_pstate;
    g.src[2] = s->act; g.dst[2] = act; g.width[2] = s->n_act;
    g.src[3] = s->rew; g.dst[3] = rew; g.width[3] = 1;
    g.src[4] = s->done; g.dst[4] = done; g.width[4] = 1;
    DG_REQUIRE(2 * s->n_pstate + s->n_act + 2 <= 256, "too many scalar fields per transition");
    // slices per frame: every thread moves GATHER_ILP float4 per pass; one pass per block at the shipped frame size
    const int slices = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(g.frame4, 256 * GATHER_ILP), 64));
    ProfScope ps(PROF_GATHER, 0.0, (double)B * (4.0 * s->frame * 4 + (2 * s->n_pstate + s->n_act + 2) * 8.0), st);
    launch_k(replay_gather_kernel, dim3((unsigned)slices, (unsigned)B, 2), 256, 0, st, g);
    DG_LAUNCH_CHECK();
  });
}

int dgvit_replay_append(const dgvit_replay* s, float* engage_store, const float* records,
                        const int64_t* slots, int n, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(s ? s->obs : nullptr);
    DG_REQUIRE(s && records && slots && n >= 0 && s->obs && s->size > 0, "bad argument");
    DG_REQUIRE(s->frame % 4 == 0 && ((uintptr_t)s->obs % 16) == 0 && ((uintptr_t)records % 16) == 0,
               "frames must be 16-byte aligned multiples of 4 floats");
    if (n == 0) return;
    AppendArgs a;
    memset(&a, 0, sizeof(a));
    a.store = (float4*)s->obs; a.size = s->size; a.frame4 = s->frame / 4;
    a.fld[0] = (float*)s->pobs; a.width[0] = s->n_pstate;
    a.fld[1] = (float*)s->next_pobs; a.width[1] = s->n_pstate;
    a.fld[2] = (float*)s->act; a.width[2] = s->n_act;
    a.fld[3] = (float*)s->rew; a.width[3] = 1;
    a.fld[4] = (float*)s->done; a.width[4] = 1;
    a.fld[5] = engage_store; a.width[5] = 1;
    const int64_t scal = 2 * s->n_pstate + s->n_act + 3;
    a.rec = records; a.rec_floats = (2 * s->frame + scal + 3) / 4 * 4;      // records are padded to 16 bytes
    a.slot = slots;
    const int slices = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(a.frame4, 1024), 16));
    launch_k(replay_append_kernel, dim3((unsigned)slices, (unsigned)n, 2), 256, 0, (cudaStream_t)stream, a);
    DG_LAUNCH_CHECK();
  });
}
int64_t dgvit_replay_record_floats(int64_t frame, int n_pstate, int n_act) {
  return (2 * frame + 2 * n_pstate + n_act + 3 + 3) / 4 * 4;
}

// the embedding-dropout keep decisions a trunk call with `drop` makes for B samples of N tokens x D features, as a
// {0,1} mask [B, N, D] (tests: the in-kernel Philox stream must equal an injected DGVIT_DROP_MASK run)
int dgvit_debug_drop_mask(const dgvit_drop* drop, int B, int N, int D, int sample_offset, uint8_t* out, void* stream) {
  return guarded([&] {
    DeviceGuard dev_guard(out);
    DG_REQUIRE(drop && out && B >= 1 && N >= 1 && D >= 1, "bad argument");
    DropDev r;
    r.mode = drop->mode; r.p = drop->p; r.scale = 1.0f; r.mask = drop->keep_mask; r.rng = drop->rng_state;
    r.stream_id = drop->stream_id; r.elem_offset = (int64_t)sample_offset * N * D;
    const int64_t total = (int64_t)B * N * D;
    launch_k(drop_mask_dump_kernel, grid1d(total), 256, 0, (cudaStream_t)stream, r, out, total);
    DG_LAUNCH_CHECK();
  });
}

}  // extern "C"
