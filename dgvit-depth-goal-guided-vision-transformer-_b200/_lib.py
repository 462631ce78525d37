"""ctypes binding of libdgvit.so (include/dgvit.h).

There is NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  PyTorch is used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DGVIT_LIB selects another build of the same library (A/B measurements); there is still no fallback
LIB_PATH = os.environ.get("DGVIT_LIB") or os.path.join(_HERE, "libdgvit.so")

MAX_DEPTH = 16
ACTOR, CRITIC, QNET = 0, 1, 2
FP32, BF16 = 0, 1
DROP_NONE, DROP_MASK, DROP_RNG = 0, 1, 2
(PROF_NONE, PROF_GEMM_MLP, PROF_GEMM_ALL, PROF_ATTENTION, PROF_GATHER, PROF_ADAM, PROF_MLP_FUSED, PROF_LN_BWD, PROF_EMBED,
 PROF_PATCH) = range(10)

c_f_p = C.c_void_p  # device pointers travel as integers


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("kind", "img_h", "img_w", "patch_h", "patch_w", "dim", "depth", "heads", "dim_head",
                 "mlp_dim", "n_act", "n_pstate")]


class BlockLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("ln1_w", "ln1_b", "qkv_w", "out_w", "out_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class Layout(C.Structure):
    _fields_ = ([("total", C.c_int64)] +
                [(n, C.c_int64) for n in ("pos", "cls", "rms_g", "patch_w", "patch_b")] +
                [("block", BlockLayout * MAX_DEPTH)] +
                [(n, C.c_int64) for n in ("mlp_head_ln_w", "mlp_head_ln_b", "mlp_head_w", "mlp_head_b",
                                          "embed_w", "embed_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
                                          "mean_w", "mean_b", "lstd_w", "lstd_b",
                                          "conv1_w", "conv1_b", "conv2_w", "conv2_b", "conv3_w", "conv3_b",
                                          "fc3_w", "fc3_b", "fc11_w", "fc11_b", "fc21_w", "fc21_b",
                                          "fc31_w", "fc31_b")] +
                [("n_skip", C.c_int32), ("skip_begin", C.c_int64 * 4), ("skip_end", C.c_int64 * 4),
                 ("alpha_grad_slot", C.c_int64)])


class Net(C.Structure):
    _fields_ = [("cfg", Cfg), ("params", c_f_p), ("grads", c_f_p), ("shadow", c_f_p)]


class Drop(C.Structure):
    _fields_ = [("mode", C.c_int32), ("p", C.c_float), ("keep_mask", c_f_p), ("rng_state", c_f_p),
                ("stream_id", C.c_uint32)]


class ActorIO(C.Structure):
    _fields_ = [("img", c_f_p), ("pstate", c_f_p), ("eps", c_f_p), ("action_scale", c_f_p),
                ("action_bias", c_f_p), ("drop", Drop), ("sample_offset", C.c_int32),
                ("mean", c_f_p), ("log_std", c_f_p), ("action", c_f_p), ("log_prob", c_f_p),
                ("mean_t", c_f_p), ("eps_out", c_f_p), ("advance_rng", C.c_int32)]


class BcIO(C.Structure):
    _fields_ = [("img", c_f_p), ("pstate", c_f_p), ("target", c_f_p), ("eps", c_f_p), ("action_scale", c_f_p),
                ("action_bias", c_f_p), ("drop", Drop), ("sample_offset", C.c_int32), ("advance_rng", C.c_int32),
                ("max_action", C.c_float), ("max_norm", C.c_float), ("loss", c_f_p), ("grad_norm", c_f_p)]


class ActorGrad(C.Structure):
    _fields_ = [("d_mean", c_f_p), ("d_log_std", c_f_p), ("d_action", c_f_p), ("d_log_prob", c_f_p),
                ("d_mean_t", c_f_p), ("d_log_prob_const", C.c_float)]


class CriticIO(C.Structure):
    _fields_ = [("img", c_f_p), ("pstate", c_f_p), ("action", c_f_p), ("drop", Drop), ("q1", c_f_p), ("q2", c_f_p)]


class TrunkIO(C.Structure):
    _fields_ = [("img", c_f_p), ("goal", c_f_p), ("drop", Drop), ("sample_offset", C.c_int32), ("z", c_f_p)]


class Adam(C.Structure):
    _fields_ = [("m", c_f_p), ("v", c_f_p), ("step", c_f_p), ("lr", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float)]


class Dp(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("multicast", c_f_p), ("peers", c_f_p), ("pads", c_f_p),
                ("arena_off", C.c_int64 * 2), ("tail", c_f_p), ("reduced_out", c_f_p * 2), ("finished", c_f_p),
                ("error_flag", c_f_p)]


class Sac(C.Structure):
    _fields_ = [("actor", Net), ("critic", Net), ("critic_target", Net), ("actor_opt", Adam), ("critic_opt", Adam),
                ("log_alpha", c_f_p), ("alpha", c_f_p), ("alpha_m", c_f_p), ("alpha_v", c_f_p), ("alpha_step", c_f_p),
                ("lr_alpha", C.c_float), ("auto_alpha", C.c_int32), ("target_entropy", C.c_float),
                ("gamma", C.c_float), ("tau", C.c_float), ("do_polyak", C.c_int32), ("precision", C.c_int32),
                ("global_batch", C.c_int32), ("n_extra", C.c_int32), ("sample_offset", C.c_int32), ("rng_state", c_f_p),
                ("action_scale", c_f_p), ("action_bias", c_f_p), ("dp", C.POINTER(Dp))]


class Batch(C.Structure):
    _fields_ = [(n, c_f_p) for n in ("obs", "next_obs", "pobs", "next_pobs", "act", "rew", "done", "extra_target",
                                     "extra_weight")]


class Noise(C.Structure):
    _fields_ = [(n, c_f_p) for n in ("eps_next", "eps_pi", "mask_a_next", "mask_ct", "mask_c", "mask_a", "mask_c_pi")] + \
               [("drop_mode", C.c_int32)]


class SacOut(C.Structure):
    _fields_ = [("losses", c_f_p), ("debug", c_f_p)]


class Replay(C.Structure):
    _fields_ = [("obs", c_f_p), ("size", C.c_int64), ("frame", C.c_int64), ("pobs", c_f_p), ("next_pobs", c_f_p),
                ("act", c_f_p), ("rew", c_f_p), ("done", c_f_p), ("n_pstate", C.c_int32), ("n_act", C.c_int32)]


class QnetLayout(C.Structure):
    _fields_ = [("conv_w", C.c_int64 * 3), ("conv_b", C.c_int64 * 3)] + [
        (n, C.c_int64) for n in ("fc1_w", "fc1_b", "fc2_w", "fc2_b", "fc3_w", "fc3_b", "embed_w", "embed_b",
                                 "fc11_w", "fc11_b", "fc21_w", "fc21_b", "fc31_w", "fc31_b", "total")]


# every symbol include/dgvit.h declares: (name, restype, argtypes)
P = C.POINTER
SYMBOLS = {
    "dgvit_version": (C.c_int, []),
    "dgvit_last_error": (C.c_char_p, []),
    "dgvit_launch_count": (C.c_longlong, []),
    "dgvit_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "dgvit_prof_begin": (C.c_int, [C.c_int, C.c_int]),
    "dgvit_prof_end": (C.c_int, [P(C.c_double), P(C.c_longlong), P(C.c_double), P(C.c_double)]),
    "dgvit_param_layout": (C.c_int, [P(Cfg), P(Layout)]),
    "dgvit_workspace_bytes": (C.c_int, [P(Cfg), C.c_int, C.c_int, C.c_int, P(C.c_size_t)]),
    "dgvit_sac_workspace_bytes": (C.c_int, [P(Cfg), C.c_int, C.c_int, C.c_int, P(C.c_size_t)]),
    "dgvit_sac_qnet_workspace_bytes": (C.c_int, [P(Cfg), C.c_int, C.c_int, C.c_int, P(C.c_size_t)]),
    "dgvit_refresh_shadow": (C.c_int, [P(Net), C.c_void_p]),
    "dgvit_actor_forward": (C.c_int, [P(Net), P(ActorIO), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_actor_backward": (C.c_int, [P(Net), P(ActorIO), P(ActorGrad), C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_bc_workspace_bytes": (C.c_int, [P(Cfg), C.c_int, C.c_int, P(C.c_size_t)]),
    "dgvit_bc_step": (C.c_int, [P(Net), P(Adam), P(BcIO), C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_trunk_workspace_bytes": (C.c_int, [P(Cfg), C.c_int, C.c_int, C.c_int, P(C.c_size_t)]),
    "dgvit_trunk_forward": (C.c_int, [P(Net), P(TrunkIO), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_trunk_backward": (C.c_int, [P(Net), P(TrunkIO), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                       C.c_void_p]),
    "dgvit_critic_forward": (C.c_int, [P(Net), P(CriticIO), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_critic_backward": (C.c_int, [P(Net), P(CriticIO), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_sac_phase1": (C.c_int, [P(Sac), P(Batch), P(Noise), P(SacOut), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_sac_phase2": (C.c_int, [P(Sac), P(Batch), P(Noise), P(SacOut), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_sac_phase3": (C.c_int, [P(Sac), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_sac_update": (C.c_int, [P(Sac), P(Batch), P(Noise), P(SacOut), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_gemm_bf16": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "dgvit_linear_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "dgvit_attention_stats_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "dgvit_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p]),
    "dgvit_mlp_bf16": (C.c_int, [C.c_void_p] * 12 + [C.c_int64, C.c_int, C.c_void_p]),
    "dgvit_mlp_partial_floats": (C.c_int64, [C.c_int64, C.c_int]),
    "dgvit_mlp_fwd_f16w2": (C.c_int, [C.c_void_p] * 7 + [C.c_int64, C.c_int, C.c_void_p]),
    "dgvit_qnet_param_layout": (C.c_int, [C.c_int, C.c_int, P(QnetLayout)]),
    "dgvit_qnet_workspace_bytes": (C.c_int, [C.c_int] * 6 + [P(C.c_size_t)]),
    "dgvit_qnet_forward": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 6 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_qnet_backward": (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 7 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "dgvit_adam_step": (C.c_int, [P(Net), P(Adam), P(Net), C.c_float, C.c_void_p]),
    "dgvit_polyak": (C.c_int, [P(Net), P(Net), C.c_float, C.c_void_p]),
    "dgvit_polyak_flat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "dgvit_replay_gather": (C.c_int, [P(Replay), C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p]),
    "dgvit_replay_record_floats": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "dgvit_replay_append": (C.c_int, [P(Replay), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "dgvit_debug_drop_mask": (C.c_int, [P(Drop), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dgvit_depth_scratch_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, P(C.c_size_t)]),
    "dgvit_depth_augment": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


def lib():
    """Load libdgvit.so once; raise loudly if it is not built (no CPU / eager fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
                "dgvit_b200 has no fallback path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        # DGVIT_OPTS="name=value,..." applies dgvit_set_option at load time (A/B measurements of kernel variants)
        for kv in filter(None, os.environ.get("DGVIT_OPTS", "").split(",")):
            k, v = kv.split("=")
            if l.dgvit_set_option(k.strip().encode(), int(v)) != 0:
                raise RuntimeError(f"DGVIT_OPTS: {l.dgvit_last_error().decode()}")
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().dgvit_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libdgvit {what} failed ({rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def layout_of(cfg: Cfg) -> Layout:
    out = Layout()
    check(lib().dgvit_param_layout(C.byref(cfg), C.byref(out)), "param_layout")
    return out
