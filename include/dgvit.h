/*
 * dgvit.h — C ABI of libdgvit.so, the B200 (sm_100a) implementation of the DGViT
 * actor-critic hot path.
 *
 * The reference (REGRAGUIahmed/DGViT-…) is pure Python/PyTorch and has no FFI of its
 * own; the boundary a maintainer would bind is the nn.Module / SAC method surface.
 * Each entry point below names the reference interface it replaces
 * (paths relative to src/vis_nav/vis_nav/ = "vn/").  The ctypes binding is shown in
 * INTEGRATION.md and implemented in dgvit_b200/_lib.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every pointer is a DEVICE pointer unless the name ends in _host.
 *   - the library owns no memory: parameters, gradients, optimizer state, workspaces and
 *     outputs are allocated by the caller (PyTorch on the Python side).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no
 *     host synchronisation and is CUDA-graph capturable.
 *   - return value: 0 = ok, negative = error; dgvit_last_error() returns a thread-local
 *     message.  No exceptions cross the boundary.
 */
#ifndef DGVIT_H_
#define DGVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGVIT_MAX_DEPTH 16
#define DGVIT_ALIGN_FLOATS 64 /* every tensor starts on a 256-byte boundary inside an arena */

enum { DGVIT_ACTOR = 0, DGVIT_CRITIC = 1,
       DGVIT_QNET = 2 /* CNN twin-Q critic `QNetwork` inside dgvit_sac: params / grads are dgvit_qnet_layout arenas, only
                         img_h, img_w, n_act, n_pstate of the cfg are read, no shadow */ };
enum { DGVIT_FP32 = 0, DGVIT_BF16 = 1 };                 /* arithmetic of the contractions */
enum { DGVIT_DROP_NONE = 0, DGVIT_DROP_MASK = 1, DGVIT_DROP_RNG = 2 };

enum {
  DGVIT_OK = 0,
  DGVIT_ERR_ARG = -1,       /* bad argument / unsupported shape */
  DGVIT_ERR_WORKSPACE = -2, /* workspace too small */
  DGVIT_ERR_CUDA = -3       /* a CUDA call failed (message has the cudaError string) */
};

/* Shapes of one network: GoTPolicy(nb_actions, nb_pstate, block, head, l_f_size)
 * vn/got_sac_network.py:173-185 / GoTQNetwork(...) :76-88; GoT(...) vn/GoalFormer.py:124. */
typedef struct dgvit_cfg {
  int32_t kind;     /* DGVIT_ACTOR | DGVIT_CRITIC (| DGVIT_QNET, dgvit_sac.critic / critic_target only) */
  int32_t img_h, img_w, patch_h, patch_w; /* 128,160,16,20 */
  int32_t dim;      /* l_f_size */
  int32_t depth;    /* block */
  int32_t heads;    /* head */
  int32_t dim_head; /* 64 */
  int32_t mlp_dim;  /* 2048 */
  int32_t n_act;    /* nb_actions (2) */
  int32_t n_pstate; /* nb_pstate (2) */
} dgvit_cfg;

/* Offsets (in floats) of every parameter tensor inside the flat arena, in the reference's
 * registration order (= module.parameters() order, which vn/utils.py:31-37 zips). */
typedef struct dgvit_block_layout {
  int64_t ln1_w, ln1_b, qkv_w, out_w, out_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b;
} dgvit_block_layout;

typedef struct dgvit_layout {
  int64_t total; /* arena length in floats (padded) */
  int64_t pos, cls, rms_g, patch_w, patch_b;
  dgvit_block_layout block[DGVIT_MAX_DEPTH];
  int64_t mlp_head_ln_w, mlp_head_ln_b, mlp_head_w, mlp_head_b; /* constructed, never used */
  /* actor heads */
  int64_t embed_w, embed_b, fc1_w, fc1_b, fc2_w, fc2_b, mean_w, mean_b, lstd_w, lstd_b;
  /* critic heads (conv1-3 are constructed, never used) */
  int64_t conv1_w, conv1_b, conv2_w, conv2_b, conv3_w, conv3_b;
  int64_t fc3_w, fc3_b, fc11_w, fc11_b, fc21_w, fc21_b, fc31_w, fc31_b;
  /* ranges Adam skips because the reference never produces a gradient for them
   * (grad is None: cls_token, mlp_head.*, conv1-3) */
  int32_t n_skip;
  int64_t skip_begin[4], skip_end[4];
  /* actor only: one extra float at the end of the arena that carries d(alpha_loss)/d(log_alpha)
   * in the GRADIENT arena, so that a single all-reduce of actor.grads covers it (data parallel) */
  int64_t alpha_grad_slot;
} dgvit_layout;

/* One network instance: caller-owned arenas, all `layout.total` elements long. */
typedef struct dgvit_net {
  dgvit_cfg cfg;
  float* params;    /* fp32 master parameters */
  float* grads;     /* fp32 gradients (written, not accumulated, by *_backward) */
  uint16_t* shadow; /* 16-bit operand copies of params for DGVIT_BF16, 2 * layout.total elements: bf16 copy in
                       [0, total), f16 copy in [total, 2 total) (the f16 x f16 second GEMM of the fused MLP forward reads
                       net.3.weight from it); may be NULL for DGVIT_FP32 */
} dgvit_net;

/* Stochastic inputs of one trunk call (vn/GoalFormer.py:163, emb dropout p=0.1). */
typedef struct dgvit_drop {
  int32_t mode;              /* DGVIT_DROP_* */
  float p;                   /* 0.1 */
  const uint8_t* keep_mask;  /* [B, N, D] {0,1}, DGVIT_DROP_MASK */
  const uint64_t* rng_state; /* device {seed, counter}, DGVIT_DROP_RNG */
  uint32_t stream_id;        /* distinguishes the calls inside one update */
} dgvit_drop;

/* ---- GoTPolicy.forward / .sample  (vn/got_sac_network.py:221-251) ------------------ */
typedef struct dgvit_actor_io {
  const float* img;    /* [B, img_h, img_w] */
  const float* pstate; /* [B, n_pstate] */
  const float* eps;    /* [B, n_act] N(0,1) draws of rsample, or NULL -> generated from rng */
  const float* action_scale; /* [n_act] */
  const float* action_bias;  /* [n_act] */
  dgvit_drop drop;
  int32_t sample_offset; /* global index of local sample 0 (data-parallel RNG slicing) */
  /* outputs (any may be NULL except mean/log_std) */
  float* mean;     /* [B, n_act]  pre-tanh mean */
  float* log_std;  /* [B, n_act]  clamped to [-20, 2] */
  float* action;   /* [B, n_act] */
  float* log_prob; /* [B, 1] */
  float* mean_t;   /* [B, n_act]  tanh(mean)*scale+bias */
  float* eps_out;  /* [B, n_act]  the eps actually used (needed by backward) */
  int32_t advance_rng; /* 1 (B <= 128 only): the call's last kernel advances the counter of drop.rng_state once all its
                          draws are made — the batch-1 act loop replays one graph and needs a fresh stream per call */
} dgvit_actor_io;

typedef struct dgvit_actor_grad {
  const float* d_mean;     /* [B, n_act] or NULL */
  const float* d_log_std;  /* [B, n_act] or NULL */
  const float* d_action;   /* [B, n_act] or NULL */
  const float* d_log_prob; /* [B, 1] or NULL */
  const float* d_mean_t;   /* [B, n_act] or NULL */
  float d_log_prob_const;  /* added to every d_log_prob element (alpha/B in SAC) */
} dgvit_actor_grad;

/* ---- GoTQNetwork.forward (vn/got_sac_network.py:107-123) ---------------------------- */
typedef struct dgvit_critic_io {
  const float* img;    /* [B, img_h, img_w] */
  const float* pstate; /* [B, n_pstate] */
  const float* action; /* [B, n_act] */
  dgvit_drop drop;
  float* q1; /* [B, n_act]  (the reference's Q heads emit nb_actions values, :97,:103) */
  float* q2; /* [B, n_act] */
} dgvit_critic_io;

/* ---- SAC.learn (vn/DRL.py:373-437) -------------------------------------------------- */
typedef struct dgvit_adam {
  float* m;       /* [layout.total] exp_avg */
  float* v;       /* [layout.total] exp_avg_sq */
  int64_t* step;  /* device scalar, incremented by the kernel */
  float lr, beta1, beta2, eps;
} dgvit_adam;

/* ---- behaviour-cloning step on the actor (vn/attention_imitating.py:48-67), SURVEY.md section 8 row f3 ------------------------
 * One call = policy.sample([img, pstate]) -> loss = sqrt(mean((clip(tanh-mean, +-max_action) - target)^2)) -> backward ->
 * torch.nn.utils.clip_grad_norm_(parameters, max_norm) -> Adam.step (opt; parameters the loss does not reach -- cls_token,
 * mlp_head.*, log_std_linear -- keep zero moments and do not move, like `grad is None` in the reference). */
typedef struct dgvit_bc_io {
  const float* img;      /* [B, img_h, img_w] */
  const float* pstate;   /* [B, n_pstate] */
  const float* target;   /* [B, n_act] demonstrated actions */
  const float* eps;      /* [B, n_act] rsample draws or NULL (the sampled action does not enter the loss) */
  const float* action_scale; const float* action_bias;   /* [n_act] */
  dgvit_drop drop;
  int32_t sample_offset;
  int32_t advance_rng;   /* 1: advance the counter of drop.rng_state after the step */
  float max_action;      /* clip range of the predicted mean (1.0 in the reference) */
  float max_norm;        /* clip_grad_norm_ threshold (10 in the reference) */
  float* loss;           /* device [1] */
  float* grad_norm;      /* device [1] total gradient norm before clipping, or NULL */
} dgvit_bc_io;
int dgvit_bc_workspace_bytes(const dgvit_cfg* actor_cfg, int B, int precision, size_t* bytes);
int dgvit_bc_step(const dgvit_net* actor, const dgvit_adam* opt, const dgvit_bc_io* io, int B, int precision,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Data parallel without NCCL calls: the gradient all-reduce is fused into the optimizer pass.  Both gradient arenas live
 * in ONE symmetric-memory buffer per rank ([critic grads | actor grads], e.g. torch.distributed._symmetric_memory): the
 * Adam kernel of a network runs a flag barrier over the ranks' signal pads, reads the rank-sum of every gradient element
 * (multimem.ld_reduce through the NVSwitch when `multicast` != NULL, else peer loads in rank order) and updates the
 * replica; dgvit_sac_update is then the whole data-parallel step (one call, one CUDA graph).  All pointers are device
 * pointers of THIS rank's address space. */
typedef struct dgvit_dp {
  int32_t world, rank;
  const float* multicast;         /* multicast address of the symmetric buffer, or NULL */
  const float* const* peers;      /* [world] the symmetric buffer on every rank (device array) */
  uint32_t* const* pads;          /* [world] signal pads, >= 4 * world uint32 each, zero-initialised (device array) */
  int64_t arena_off[2];           /* float offset of the critic / actor gradient arena inside the buffer */
  float* tail;                    /* [64] local: reduced alpha-gradient slot and loss sums (actor arena tail) */
  float* reduced_out[2];          /* optional [layout.total] each: the reduced gradients (tests), else NULL */
  unsigned int* finished;         /* [2] local zero-initialised counters */
  int32_t* error_flag;            /* local: set to 1 when a peer did not show up within ~3 s */
} dgvit_dp;

typedef struct dgvit_sac {
  dgvit_net actor, critic, critic_target;
  dgvit_adam actor_opt, critic_opt;
  /* entropy temperature (vn/DRL.py:136-139,416-424); device scalars */
  float* log_alpha; float* alpha; float* alpha_m; float* alpha_v; int64_t* alpha_step;
  float lr_alpha; int32_t auto_alpha; float target_entropy;
  float gamma, tau;
  int32_t do_polyak;      /* itera % policy_freq == 0 (vn/DRL.py:430) */
  int32_t precision;      /* DGVIT_FP32 | DGVIT_BF16 */
  int32_t global_batch;   /* loss means are taken over this many samples (data parallel) */
  int32_t n_extra;        /* learn_guidence: imitation rows appended to the actor's batch (0 for learn) */
  int32_t sample_offset;  /* this rank's first sample inside the global batch */
  uint64_t* rng_state;    /* device {seed, counter}; counter advanced once per update */
  const float* action_scale; const float* action_bias;
  const dgvit_dp* dp;     /* NULL: single GPU, or gradients all-reduced by the caller between the phases */
} dgvit_sac;

typedef struct dgvit_batch {   /* one replay minibatch, already on the device */
  const float *obs, *next_obs;   /* [B (+ n_extra for obs), img_h, img_w] */
  const float *pobs, *next_pobs; /* [B (+ n_extra for pobs), n_pstate] */
  const float *act;              /* [B, n_act] */
  const float *rew;              /* [B, 1] */
  const float *done;             /* [B, 1]  read by the reference, unused (vn/DRL.py:393) */
  /* SAC.learn_guidence (vn/DRL.py:257-278): rows B .. B+n_extra-1 of obs/pobs are imitation rows (expert
   * minibatch / engaged rows) seen only by the actor; they add  sum_r weight[r] * |tanh-mean_r - target_r|^2
   * to the policy loss (weight = guidence_weight / (rows * n_act), resp. engage_weight / ...). */
  const float *extra_target;     /* [n_extra, n_act] or NULL */
  const float *extra_weight;     /* [n_extra] or NULL */
} dgvit_batch;

typedef struct dgvit_noise {   /* parity mode: injected stochastic inputs; all NULL = RNG */
  const float *eps_next, *eps_pi;                  /* [B, n_act] (eps_pi: B + n_extra rows) */
  const uint8_t *mask_a_next, *mask_ct, *mask_c, *mask_a, *mask_c_pi; /* [B, N, D] (mask_a: B + n_extra rows) */
  int32_t drop_mode;                               /* DGVIT_DROP_* */
} dgvit_noise;

typedef struct dgvit_sac_out {
  float* losses;   /* [4] qf1_loss, policy_loss, qf2_loss, alpha_loss (sums over the local
                      shard already divided by global_batch; all-reduce SUM across ranks) */
  float* debug;    /* optional [B*(5*n_act+1)]: nq, q1, q2, pi, (q1pi) , log_pi ; or NULL */
} dgvit_sac_out;

/* -------------------------------------------------------------------------------------- */
int dgvit_version(void);
const char* dgvit_last_error(void);

/* measurement hooks (bench.py): number of kernels this library has launched so far, and
 * CUDA-event timing of the launches of one kernel family (DGVIT_PROF_*), recorded on the
 * launching stream.  dgvit_prof_end synchronises on the recorded events. */
enum { DGVIT_PROF_NONE = 0, DGVIT_PROF_GEMM_MLP = 1, DGVIT_PROF_GEMM_ALL = 2, DGVIT_PROF_ATTENTION = 3,
       DGVIT_PROF_GATHER = 4, DGVIT_PROF_ADAM = 5, DGVIT_PROF_MLP_FUSED = 6, DGVIT_PROF_LN_BWD = 7,
       DGVIT_PROF_EMBED = 8, DGVIT_PROF_PATCH = 9 };
long long dgvit_launch_count(void);
/* runtime switches for measurement / A-B tests (results are identical for every setting up to summation order; the
 * stream switches change nothing at all, tests/test_gpu_parity.py):
 *   "fork_streams"   1: the independent forward passes of the update run on library-owned streams; 0: caller's stream only
 *   "bwd_side"       1: weight-gradient launches, per-block reductions and bookkeeping kernels on a second stream
 *   "actor_s_when"   0/1/2: policy.sample(s) forward starts at the fork / after the policy.sample(s') forward / after
 *                    the target-critic forward
 *   "tensor_cores"   0: route the bf16 contractions to the CUDA-core kernels
 *   "mlp_h16"        0: bf16 hidden tile and bf16 W2 in the fused MLP forward (1: f16 tile + f16 copy of W2)
 *   "attn_long"      0: sequences of more than 128 tokens go to the CUDA-core attention kernels
 *   "mlp_split", "mlp_front", "attention_row0", "attn_bwd2": kernel variants of the pruned last block / fused prologue /
 *                    pipelined attention backward (0 = the plain kernels)
 *   "ln_bwd_warps", "ln_bwd_blocks_per_sm": LayerNorm-backward block shape; "pdl": programmatic dependent launch;
 *   "skip": bit mask of kernel families to drop (timing attribution only, results become garbage) */
int dgvit_set_option(const char* name, int value);
int dgvit_prof_begin(int tag, int max_launches);
int dgvit_prof_end(double* ms_total, long long* launches, double* flops, double* bytes);

/* parameter arena layout for `cfg` (mirrors module.parameters() order) */
int dgvit_param_layout(const dgvit_cfg* cfg, dgvit_layout* out);

/* bytes of workspace a forward (+ saved activations when save_for_backward) needs */
int dgvit_workspace_bytes(const dgvit_cfg* cfg, int B, int precision, int save_for_backward,
                          size_t* bytes);
/* bytes of workspace dgvit_sac_* needs for a local batch of B */
int dgvit_sac_workspace_bytes(const dgvit_cfg* actor_cfg, int B, int n_extra, int precision, size_t* bytes);
/* the same when dgvit_sac.critic / critic_target are the CNN twin-Q critic (cfg.kind == DGVIT_QNET): the reference's shipped
 * default critic_type (vn/config.yaml:61, vn/DRL.py:118-121).  dgvit_sac_update / _phase1-3 then run vn/DRL.py:388-434 with
 * QNetwork passes (vn/got_sac_network.py:125-170) in place of the GoT critic's; critic_opt spans the QNetwork arena. */
int dgvit_sac_qnet_workspace_bytes(const dgvit_cfg* actor_cfg, int B, int n_extra, int precision, size_t* bytes);

/* refresh both 16-bit shadows of a parameter arena (after load_state_dict etc.) */
int dgvit_refresh_shadow(const dgvit_net* net, void* stream);

/* GoTPolicy.forward/.sample — vn/got_sac_network.py:221-251 */
int dgvit_actor_forward(const dgvit_net* net, const dgvit_actor_io* io, int B, int precision,
                        int save_for_backward, void* workspace, size_t workspace_bytes,
                        void* stream);
/* backward of the above through the saved workspace; writes net->grads */
int dgvit_actor_backward(const dgvit_net* net, const dgvit_actor_io* io,
                         const dgvit_actor_grad* g, int B, int precision, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- GoT.forward(img, goal) (vn/GoalFormer.py:156-171): the trunk on its own.  goal [B, dim] is the goal token
 * (the caller's fc_embed output), z [B, dim] the pooled, RMS-normalised token 0.  `net` may be an actor or a critic
 * arena (the trunk tensors sit at the same offsets).  backward: d_z [B, dim] -> the trunk ranges of net->grads
 * (head ranges untouched) and d_goal [B, dim] (optional). */
typedef struct dgvit_trunk_io {
  const float* img;   /* [B, img_h, img_w] */
  const float* goal;  /* [B, dim] */
  dgvit_drop drop;
  int32_t sample_offset;
  float* z;           /* [B, dim] */
} dgvit_trunk_io;
int dgvit_trunk_workspace_bytes(const dgvit_cfg* cfg, int B, int precision, int save_for_backward, size_t* bytes);
int dgvit_trunk_forward(const dgvit_net* net, const dgvit_trunk_io* io, int B, int precision, int save_for_backward,
                        void* workspace, size_t workspace_bytes, void* stream);
int dgvit_trunk_backward(const dgvit_net* net, const dgvit_trunk_io* io, const float* d_z, float* d_goal, int B,
                         int precision, void* workspace, size_t workspace_bytes, void* stream);

/* GoTQNetwork.forward — vn/got_sac_network.py:107-123 */
int dgvit_critic_forward(const dgvit_net* net, const dgvit_critic_io* io, int B, int precision,
                         int save_for_backward, void* workspace, size_t workspace_bytes,
                         void* stream);
/* backward: d_q1,d_q2 [B,n_act]; writes net->grads when param_grads != 0 and
 * d_action [B,n_act] when non-NULL */
int dgvit_critic_backward(const dgvit_net* net, const dgvit_critic_io* io, const float* d_q1,
                          const float* d_q2, float* d_action, int param_grads, int B,
                          int precision, void* workspace, size_t workspace_bytes, void* stream);

/* SAC.learn in three phases so that a data-parallel caller can all-reduce the gradient
 * arenas between them (vn/DRL.py:388-402 | :404-413 | :414-432):
 *   phase 1: TD target + critic forward/backward        -> critic.grads
 *   phase 2: critic Adam (+ Polyak target update beside the rest of the phase); critic(s,pi) with the
 *            policy sample of phase 1; policy/alpha losses; backward -> actor.grads (+ alpha grad slot)
 *   phase 3: actor Adam, alpha Adam, RNG counter advance
 * dgvit_sac_update = phases 1-3 back to back (single GPU).
 * Stream contract: each call forks library-owned streams off `stream` with events (independent forward passes,
 * weight-gradient launches, bookkeeping kernels) and joins every one of them back into `stream` before it returns, so
 * the caller sees ordinary stream-ordered work and the call is capturable into one CUDA graph. */
int dgvit_sac_phase1(const dgvit_sac* s, const dgvit_batch* b, const dgvit_noise* nz,
                     const dgvit_sac_out* out, int B, void* workspace, size_t workspace_bytes,
                     void* stream);
int dgvit_sac_phase2(const dgvit_sac* s, const dgvit_batch* b, const dgvit_noise* nz,
                     const dgvit_sac_out* out, int B, void* workspace, size_t workspace_bytes,
                     void* stream);
int dgvit_sac_phase3(const dgvit_sac* s, int B, void* workspace, size_t workspace_bytes,
                     void* stream);
int dgvit_sac_update(const dgvit_sac* s, const dgvit_batch* b, const dgvit_noise* nz,
                     const dgvit_sac_out* out, int B, void* workspace, size_t workspace_bytes,
                     void* stream);

/* bf16 GEMM primitive behind every contraction of the bf16 path (nn.Linear forward / dX / dW,
 * vn/GoalFormer.py:43,46,64,67,139): C[m,n] (fp32) = sum_k A(m,k) B(k,n) with element strides
 * A(m,k)=A[m*a_sm+k*a_sk], B(k,n)=B[k*b_sk+n*b_sn].  use_tensor_cores=1 -> tcgen05/TMA kernel
 * (fails if the shape is not eligible), 0 -> CUDA-core kernel (test cross-check). */
int dgvit_gemm_bf16(int M, int N, int K, const void* A, int64_t a_sm, int64_t a_sk, const void* B,
                    int64_t b_sk, int64_t b_sn, float* C, int64_t ldc, int splitk, float* partial,
                    int use_tensor_cores, void* stream);

/* bf16 nn.Linear with the fused epilogues of the MLP path (vn/GoalFormer.py:43-46): y[rows,N] (bf16) =
 * x[rows,K] W^T (W [N,K]; or x W with W [K,N] when weight_is_kn).  epilogue: 0 none; 1 y = acc+bias,
 * y2 = gelu(y); 2 y = acc * gelu'(aux); 3 as 2 and y2 = gelu(aux).  (unit tests, micro-benchmarks) */
int dgvit_linear_bf16(const void* x, const void* W, void* y, int64_t rows, int N, int K, int epilogue,
                      const float* bias, const void* aux, void* y2, int weight_is_kn, void* stream);

/* Attention.forward core (vn/GoalFormer.py:75-81) on bf16 QKV [B*N, 3*H*dim_head] (column =
 * which*inner + h*dim_head + d): forward when d_o == NULL (writes o [B*N, inner]), else backward
 * (reads o, d_o; writes d_qkv).  use_tensor_cores=1 -> tcgen05/TMEM kernels, 0 -> CUDA-core kernel.
 * stats: scratch of dgvit_attention_stats_floats(B, N, H) floats that the forward fills (softmax log-sum-exp) and the
 * backward of the same inputs reads; needed for N > 128 (the key-chunked kernels), may be NULL otherwise. */
int64_t dgvit_attention_stats_floats(int B, int N, int H);
int dgvit_attention_bf16(const void* qkv, void* o, const void* d_o, void* d_qkv, int B, int N, int H,
                         int dim_head, int use_tensor_cores, float* stats, void* stream);

/* FeedForward.forward + residual (vn/GoalFormer.py:39-50,104) on bf16 operands, fused tcgen05 kernels, D = 64:
 * forward  (d_y == NULL): out[rows,64] (fp32) = resid + W2 gelu(W1 x + b1) + b2, x [rows,64] bf16 (the LayerNorm output),
 *                         W1 [hid,64], W2 [64,hid] bf16.
 * backward (d_y != NULL): d_x[rows,64] (fp32) = gradient w.r.t. x; d_w = [dW1 (hid*64) | db1 (hid) | dW2 (64*hid)] fp32,
 *                         d_b2 [64] fp32; partial = scratch of dgvit_mlp_partial_floats(rows, hid) floats.
 * (unit tests, micro-benchmarks; the update calls the same kernels internally) */
int dgvit_mlp_bf16(const void* x, const void* W1, const float* b1, const void* W2, const float* b2,
                   const float* resid, float* out, const void* d_y, float* d_x, float* d_w, float* d_b2,
                   float* partial, int64_t rows, int hid, void* stream);
int64_t dgvit_mlp_partial_floats(int64_t rows, int hid);
/* the forward as the update runs it: f16 hidden tile and an f16 copy of W2 (W2_f16 [64,hid]), W1 / x bf16 */
int dgvit_mlp_fwd_f16w2(const void* x, const void* W1, const float* b1, const void* W2_f16, const float* b2,
                        const float* resid, float* out, int64_t rows, int hid, void* stream);

/* ---- CNN twin-Q critic `QNetwork` (vn/got_sac_network.py:125-170; the reference's shipped default critic_type,
 * vn/config.yaml:61, vn/DRL.py:118-121).  Parameters live in one flat fp32 arena in the reference's registration
 * order (conv1-3, fc1, fc2, fc3, fc_embed, fc11, fc21, fc31; weight then bias), each tensor 256-byte aligned. */
typedef struct dgvit_qnet_layout {
  int64_t conv_w[3], conv_b[3];          /* [16,1,5,5] [64,16,5,5] [256,64,5,5] */
  int64_t fc1_w, fc1_b, fc2_w, fc2_b, fc3_w, fc3_b;     /* [128, 256+32+n_act] [32,128] [n_act,32] */
  int64_t embed_w, embed_b;              /* fc_embed [32, n_pstate] */
  int64_t fc11_w, fc11_b, fc21_w, fc21_b, fc31_w, fc31_b;
  int64_t total;                         /* floats */
} dgvit_qnet_layout;
int dgvit_qnet_param_layout(int n_act, int n_pstate, dgvit_qnet_layout* out);
int dgvit_qnet_workspace_bytes(int img_h, int img_w, int n_act, int n_pstate, int B, int precision, size_t* bytes);
/* QNetwork.forward([img, pstate, action]) -> (q1, q2), each [B, n_act].  img [B, img_h, img_w] fp32.  The workspace
 * keeps the activations: pass the same one to dgvit_qnet_backward. */
int dgvit_qnet_forward(const float* params, const float* img, const float* pstate, const float* action, float* q1,
                       float* q2, int img_h, int img_w, int n_act, int n_pstate, int B, int precision,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Backward of the forward that last ran on `workspace`: d_action [B, n_act] (optional) and, when param_grads != 0,
 * every parameter gradient into `grads` (same layout as params, overwritten). */
int dgvit_qnet_backward(const float* params, float* grads, const float* img, const float* pstate, const float* d_q1,
                        const float* d_q2, float* d_action, int param_grads, int img_h, int img_w, int n_act,
                        int n_pstate, int B, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* torch.optim.Adam.step (+ optional fused Polyak target update, vn/utils.py:31-33, and bf16
 * shadow refresh) over a flat arena; skips layout.skip ranges */
int dgvit_adam_step(const dgvit_net* net, const dgvit_adam* opt, const dgvit_net* polyak_target,
                    float tau, void* stream);
/* soft_update / hard_update (tau = 1) over all parameters — vn/utils.py:31-37 */
int dgvit_polyak(const dgvit_net* target, const dgvit_net* source, float tau, void* stream);
/* the same over two flat fp32 arenas of n floats (modules without a dgvit_cfg: the CNN critic) */
int dgvit_polyak_flat(float* target, const float* source, int64_t n, float tau, void* stream);

/* cpprb sample() row gather + staging — vn/DRL.py:375-386.  Bit-exact copies.
 * store_obs [size, frame] with next_obs[i] = obs[(i+1)%size] (cpprb next_of="obs"). */
typedef struct dgvit_replay {
  const float* obs; int64_t size; int64_t frame; /* frame = img_h*img_w floats */
  const float *pobs, *next_pobs, *act, *rew, *done; int32_t n_pstate, n_act;
} dgvit_replay;
int dgvit_replay_gather(const dgvit_replay* store, const int64_t* idx, int B, float* obs,
                        float* next_obs, float* pobs, float* next_pobs, float* act, float* rew,
                        float* done, void* stream);

/* store_transition / initialize_expert_buffer (vn/DRL.py:449-477): n packed transition records
 *   [obs frame | next_obs frame | pobs | next_pobs | act | rew | done | engage]  (floats, record pitch
 *   dgvit_replay_record_floats(): padded to 16 bytes; device memory or pinned host memory read zero-copy)
 * are scattered into the ring store by ONE kernel: obs -> row slots[i], next_obs -> row (slots[i]+1) % size (cpprb
 * next_of="obs": when two records of one call target the same row the later one wins), scalar fields -> row slots[i].
 * engage_store [size, 1]: the one field dgvit_replay does not carry (the gather never returns it); may be NULL. */
int64_t dgvit_replay_record_floats(int64_t frame, int n_pstate, int n_act);
int dgvit_replay_append(const dgvit_replay* store, float* engage_store, const float* records, const int64_t* slots,
                        int n, void* stream);

/* test hook: the embedding-dropout keep decisions ({0,1}, [B, N, D]) a trunk call with `drop` makes for samples
 * sample_offset .. sample_offset+B-1 (DGVIT_DROP_RNG: the Philox stream the kernels evaluate in place) */
int dgvit_debug_drop_mask(const dgvit_drop* drop, int B, int N, int D, int sample_offset, uint8_t* out, void* stream);

/* depth normalise + noise + blur + resize — vn/env_lab.py:420-434,78-90,69-76,295-299.
 * raw [n, H, W] f32, noise [n, H, W] f32 (N(0,50) draws) or NULL (generated from rng),
 * out [n, H/4, W/4] f32 in [0,1]; scratch >= dgvit_depth_scratch_bytes(n,H,W). */
int dgvit_depth_scratch_bytes(int n, int H, int W, size_t* bytes);
int dgvit_depth_augment(const float* raw, const float* noise, const uint64_t* rng_state, int n,
                        int H, int W, float* out, void* scratch, size_t scratch_bytes,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGVIT_H_ */
