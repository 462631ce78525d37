"""SAC agent with the reference's method surface (vn/DRL.py:34-510), fused on the GPU.

``SAC.learn`` runs the whole update (replay gather -> TD target -> critic fwd/bwd/Adam ->
actor fwd + critic(s,pi) -> policy/alpha losses -> actor bwd/Adam -> alpha Adam -> Polyak)
as hand-written CUDA kernels behind three C calls, optionally replayed from one CUDA graph.
Data parallel: one process per GPU, gradient arenas all-reduced (NCCL) between the phases.
"""
from __future__ import annotations

import os
import copy
import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib as L
from .modules import GoTPolicy, GoTQNetwork, QNetwork, set_seed, _stream
from .ops import hard_update, soft_update
from .parallel import allreduce_sum_


# --------------------------------------------------------------------------- replay store
class ReplayStore:
    """Device-resident ring store with the fields and ``next_of="obs"`` aliasing of the
    reference's cpprb buffer (vn/DRL.py:80-89).  Index *selection* stays on the host side of
    the boundary (the reference never updates priorities, so PER sampling is uniform);
    the row gather is a bit-exact CUDA kernel (``dgvit_replay_gather``) and transitions enter the
    store through one packed host record + one scatter kernel (``dgvit_replay_append``).

    Ring layout: ``cap = size + 1`` frame rows; transition at row ``i`` has its ``next_obs`` in row
    ``(i + 1) % cap``.  ``head`` is the row the next transition's ``obs`` goes to; it always holds the
    newest ``next_obs`` and is never a sampleable transition, so a stored transition can not be paired
    with a frame of a later episode once the ring has wrapped (cpprb keeps that frame in a side cache)."""

    STAGE_RECORDS = 64          # transitions per packed staging buffer

    def __init__(self, size: int, obs_shape=(128, 160), action_dim=2, pstate_dim=2, device="cuda", seed=0):
        self.size = int(size)
        self.obs_shape = tuple(obs_shape)
        self.device = torch.device(device)
        f = obs_shape[0] * obs_shape[1]
        self.cap = self.size + 1
        self.obs = torch.zeros(self.cap, f, dtype=torch.float32, device=self.device)
        self.pobs = torch.zeros(self.cap, pstate_dim, dtype=torch.float32, device=self.device)
        self.next_pobs = torch.zeros(self.cap, pstate_dim, dtype=torch.float32, device=self.device)
        self.act = torch.zeros(self.cap, action_dim, dtype=torch.float32, device=self.device)
        self.rew = torch.zeros(self.cap, 1, dtype=torch.float32, device=self.device)
        self.done = torch.zeros(self.cap, 1, dtype=torch.float32, device=self.device)
        self.engage = torch.zeros(self.cap, 1, dtype=torch.float32, device=self.device)
        self.engage_host = np.zeros(self.cap, dtype=np.float32)     # host mirror: engaged rows are selected on the host
        self.stored = 0
        self.head = 0
        self.action_dim, self.pstate_dim = action_dim, pstate_dim
        self._gen = torch.Generator().manual_seed(seed)
        self._rec = int(L.lib().dgvit_replay_record_floats(f, pstate_dim, action_dim))
        self._stage = None
        self._stage_slot = 0

    def get_stored_size(self):
        return self.stored

    def _struct(self) -> L.Replay:
        return L.Replay(obs=self.obs.data_ptr(), size=self.cap, frame=self.obs.shape[1], pobs=self.pobs.data_ptr(),
                        next_pobs=self.next_pobs.data_ptr(), act=self.act.data_ptr(), rew=self.rew.data_ptr(),
                        done=self.done.data_ptr(), n_pstate=self.pstate_dim, n_act=self.action_dim)

    # ---- write path (vn/DRL.py:449-477)
    def add(self, obs, act, pobs, next_pobs, rew, next_obs, engage=0.0, done=0.0):
        """One transition, or a batch when ``obs`` carries a leading transition axis (cpprb ``add`` accepts both;
        ``initialize_expert_buffer`` is called with whole demonstration datasets, vn/main.py:264-266)."""
        f = self.obs.shape[1]
        o = np.asarray(obs, dtype=np.float32)
        H, W = self.obs_shape
        ok = (o.shape[-2:] == (H, W) and o.ndim in (2, 3)) or (o.shape[-3:] == (H, W, 1) and o.ndim in (3, 4)) or \
             (o.shape[-1] == f and o.ndim in (1, 2))
        if not ok:      # e.g. the 4-channel frame-stacked demonstrations of the legacy pipeline
            raise ValueError(f"obs of shape {o.shape} is not [n,]{H}x{W}[x1] depth frames")
        n = o.size // f
        col = lambda v, w: np.broadcast_to(np.asarray(v, dtype=np.float32).reshape(-1, w) if np.ndim(v) else
                                           np.full((1, w), float(v), np.float32), (n, w))
        self.add_batch(o.reshape(n, f), np.asarray(next_obs, dtype=np.float32).reshape(n, f), col(pobs, self.pstate_dim),
                       col(next_pobs, self.pstate_dim), col(act, self.action_dim), col(rew, 1), col(done, 1), col(engage, 1))

    def add_batch(self, obs, next_obs, pobs, next_pobs, act, rew, done, engage):
        """n transitions (2-D float32 arrays, oldest first) -> packed pinned records -> one H2D copy and one scatter
        kernel per ``STAGE_RECORDS`` transitions."""
        n, f, R = obs.shape[0], self.obs.shape[1], self._rec
        if self._stage is None:
            K = self.STAGE_RECORDS
            self._stage = [dict(host=torch.zeros(K, R, dtype=torch.float32).pin_memory(),
                                slots=torch.zeros(K, dtype=torch.int64).pin_memory(),
                                dev=torch.zeros(K, R, dtype=torch.float32, device=self.device),
                                dslots=torch.zeros(K, dtype=torch.int64, device=self.device),
                                ev=torch.cuda.Event()) for _ in range(2)]
        st = self._struct()
        # at most `size` records per launch: a longer run would wrap onto rows written by the same launch
        step = min(self.STAGE_RECORDS, self.size)
        for a in range(0, n, step):
            b = min(n, a + step)
            k = b - a
            sg = self._stage[self._stage_slot]
            self._stage_slot ^= 1
            sg["ev"].synchronize()                     # the copy that last read this pinned buffer has completed
            rec = sg["host"].numpy()
            rec[:k, :f] = obs[a:b]
            rec[:k, f:2 * f] = next_obs[a:b]
            c = 2 * f
            for arr, w in ((pobs, self.pstate_dim), (next_pobs, self.pstate_dim), (act, self.action_dim), (rew, 1),
                           (done, 1), (engage, 1)):
                rec[:k, c:c + w] = arr[a:b]
                c += w
            slots = sg["slots"].numpy()
            for i in range(k):
                slots[i] = self.head
                self.engage_host[self.head] = engage[a + i, 0]
                self.head = (self.head + 1) % self.cap
                self.stored = min(self.stored + 1, self.size)
            stream = torch.cuda.current_stream(self.device)
            sg["dev"][:k].copy_(sg["host"][:k], non_blocking=True)
            sg["dslots"][:k].copy_(sg["slots"][:k], non_blocking=True)
            sg["ev"].record(stream)
            L.check(L.lib().dgvit_replay_append(C.byref(st), self.engage.data_ptr(), sg["dev"].data_ptr(),
                                                sg["dslots"].data_ptr(), k, stream.cuda_stream), "replay_append")

    def fill_synthetic(self, n: int, seed: int = 3407):
        """Synthetic transitions written straight on the device."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        n = min(n, self.size)
        self.obs[: n + 1] = torch.rand(n + 1, self.obs.shape[1], device=self.device, generator=g)
        self.pobs[:n, 0] = torch.rand(n, device=self.device, generator=g)
        self.pobs[:n, 1] = torch.rand(n, device=self.device, generator=g) * 2 - 1
        self.next_pobs[:n, 0] = torch.rand(n, device=self.device, generator=g)
        self.next_pobs[:n, 1] = torch.rand(n, device=self.device, generator=g) * 2 - 1
        self.act[:n] = torch.rand(n, self.action_dim, device=self.device, generator=g) * 2 - 1
        self.rew[:n] = (torch.randn(n, 1, device=self.device, generator=g) * 20).clamp(-200, 500)
        self.done[:n] = (torch.rand(n, 1, device=self.device, generator=g) < 0.01).float()
        self.stored, self.head = n, n % self.cap

    # ---- read path (vn/DRL.py:375-386)
    def live_rows(self) -> np.ndarray:
        """Store rows of the live transitions, oldest first."""
        return (self.head - self.stored + np.arange(self.stored)) % self.cap

    def sample_indexes(self, batch_size: int) -> torch.Tensor:
        k = torch.randint(0, max(self.stored, 1), (batch_size,), generator=self._gen, dtype=torch.int64)
        first = (self.head - self.stored) % self.cap
        return k if first == 0 else (k + first) % self.cap

    def gather(self, idx: torch.Tensor, out: Dict[str, torch.Tensor]):
        """out: dict of preallocated device tensors obs,next_obs,pobs,next_pobs,act,rew,done."""
        B = idx.numel()
        st = self._struct()
        L.check(L.lib().dgvit_replay_gather(C.byref(st), idx.data_ptr(), B, out["obs"].data_ptr(),
                                            out["next_obs"].data_ptr(), out["pobs"].data_ptr(),
                                            out["next_pobs"].data_ptr(), out["act"].data_ptr(), out["rew"].data_ptr(),
                                            out["done"].data_ptr(), _stream(self.device)), "replay_gather")

    # ---- on-disk format (vn/DRL.py:505-510 call cpprb's save_transitions / load_transitions; cpprb is absent here, so the
    #      file is this store's own .npz: one array per field, oldest transition first)
    def save_transitions(self, file: str):
        rows = torch.as_tensor(self.live_rows(), device=self.device)
        nxt = (rows + 1) % self.cap
        cpu = lambda t: t.cpu().numpy()
        np.savez(file if str(file).endswith(".npz") else str(file) + ".npz",
                 obs=cpu(self.obs[rows]).reshape(-1, *self.obs_shape), next_obs=cpu(self.obs[nxt]).reshape(-1, *self.obs_shape),
                 pobs=cpu(self.pobs[rows]), next_pobs=cpu(self.next_pobs[rows]), act=cpu(self.act[rows]),
                 rew=cpu(self.rew[rows]), done=cpu(self.done[rows]), engage=cpu(self.engage[rows]))

    def load_transitions(self, file: str):
        d = np.load(file)
        n = d["obs"].shape[0]
        eng = d["engage"] if "engage" in d.files else np.zeros((n, 1), np.float32)
        self.add(d["obs"], d["act"], d["pobs"], d["next_pobs"], d["rew"], d["next_obs"], engage=eng, done=d["done"])


# --------------------------------------------------------------------------- agent
class SAC(object):
    """Reference-compatible constructor (vn/DRL.py:35-39) plus keyword-only backend options."""

    def __init__(self, action_dim, pstate_dim, policy_type, critic_type, policy_attention_fix,
                 critic_attention_fix, pre_buffer, seed, LR_C=1e-3, LR_A=1e-3, LR_ALPHA=1e-4,
                 BUFFER_SIZE=int(2e5), TAU=5e-3, POLICY_FREQ=2, GAMMA=0.99, ALPHA=0.05, block=2, head=4,
                 l_f_size=32, buffer_size_expert=10816, automatic_entropy_tuning=True, *,
                 precision="bf16", device=None, image_size=(128, 160), mlp_dim=2048,
                 distributed=False, use_cuda_graph=False):
        if policy_type != "GaussianTransformer":
            raise NotImplementedError(f"policy_type={policy_type!r}: only the DGViT actor ('GaussianTransformer') "
                                      "is on the accelerated path")
        # any other critic_type selects the CNN twin-Q critic, as in vn/DRL.py:118-121
        self._cnn_critic = critic_type != "Transformer"
        if policy_attention_fix or critic_attention_fix:
            raise NotImplementedError("attention_fix (frozen trunk) variants are not on the accelerated path")
        if not torch.cuda.is_available():
            raise RuntimeError("dgvit_b200.SAC needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        self.gamma, self.tau, self.alpha = GAMMA, TAU, ALPHA
        self.pstate_dim, self.action_dim = pstate_dim, action_dim
        self.itera = 0
        self.policy_type, self.critic_type = policy_type, critic_type
        self.policy_freq = POLICY_FREQ
        self.automatic_entropy_tuning = automatic_entropy_tuning
        self.pre_buffer = pre_buffer
        self.seed = int(seed)
        self.block, self.head, self.l_f_size = block, head, l_f_size
        self.precision = precision
        self.lr_a, self.lr_c, self.lr_alpha = LR_A, LR_C, LR_ALPHA
        self.distributed = bool(distributed)
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graph_nccl = os.environ.get("DGVIT_GRAPH_NCCL", "0") == "1"

        torch.manual_seed(self.seed)
        torch.cuda.manual_seed(self.seed)
        np.random.seed(self.seed)
        set_seed(self.seed)

        # data-parallel ranks draw different minibatch indexes (same model seed, rank-offset sampler seed)
        rk = self.rank if (self.distributed and torch.distributed.is_initialized()) else 0
        self.replay_buffer = ReplayStore(BUFFER_SIZE, image_size, action_dim, pstate_dim, self.device,
                                         self.seed + 7919 * rk)
        self.buffer_size_expert = buffer_size_expert + 1                    # vn/DRL.py:53
        self.guidence_weight, self.engage_weight, self.batch_expert = 1.0, 1.0, 0          # vn/DRL.py:51-52,54
        self.replay_buffer_expert = (ReplayStore(self.buffer_size_expert, image_size, action_dim, pstate_dim, self.device,
                                                 self.seed + 1 + 7919 * rk) if pre_buffer else None)    # vn/DRL.py:91-100
        self._gbuf = {}

        # construction order == reference (critic, critic_target, policy): same seed -> same weights
        kw = dict(image_size=image_size, mlp_dim=mlp_dim)
        if self._cnn_critic:      # vn/DRL.py:118-121 (the shipped default, vn/config.yaml:61): CNN twin-Q critic
            self.critic = QNetwork(action_dim, pstate_dim, image_size=image_size).to(self.device)
            self.critic_target = QNetwork(action_dim, pstate_dim, image_size=image_size).to(self.device)
        else:
            self.critic = GoTQNetwork(action_dim, pstate_dim, block, head, l_f_size, **kw).to(self.device)
            self.critic_target = GoTQNetwork(action_dim, pstate_dim, block, head, l_f_size, **kw).to(self.device)
        self.target_entropy = -float(action_dim)
        self.policy = GoTPolicy(action_dim, pstate_dim, block, head, l_f_size, **kw).to(self.device)
        for m in (self.critic, self.critic_target, self.policy):
            m.precision = precision
            m.bind()
        self.critic_target._arena.copy_(self.critic._arena)          # hard_update (vn/DRL.py:123)
        self.target_policy = copy.deepcopy(self.policy)               # vn/DRL.py:169 (unused afterwards)

        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.log_alpha = torch.zeros(1, **f32)
        self._alpha = torch.full((1,), float(ALPHA), **f32)
        self._alpha_m = torch.zeros(1, **f32)
        self._alpha_v = torch.zeros(1, **f32)
        self._alpha_step = torch.zeros(1, dtype=torch.int64, device=dev)
        self._opt = {}
        for name, mod in (("actor", self.policy), ("critic", self.critic)):
            n = mod.layout().total
            self._opt[name] = dict(m=torch.zeros(n, **f32), v=torch.zeros(n, **f32),
                                   step=torch.zeros(1, dtype=torch.int64, device=dev))
        self._rng = torch.tensor([self.seed, 0], dtype=torch.int64, device=dev)
        self._dp = None
        if self.distributed and self.world > 1 and os.environ.get("DGVIT_DP_FUSED", "1") == "1":
            self._setup_fused_dp(keep_reduced=os.environ.get("DGVIT_DP_KEEP_REDUCED", "0") == "1")
        na = action_dim
        self._scale = self.policy.action_scale.to(dev, torch.float32).expand(na).contiguous()
        self._bias = self.policy.action_bias.to(dev, torch.float32).expand(na).contiguous()
        self._losses = torch.zeros(4, **f32)
        self._ws = None
        self._batch = None
        self._graphs = {}
        for m in (self.critic, self.critic_target, self.policy):
            if hasattr(m, "refresh_shadow"):
                m.refresh_shadow()

    # ------------------------------------------------------------------ CNN critic (vn/DRL.py:118-121)
    def _learn_cnn(self, batch: Dict[str, torch.Tensor], noise: Optional[Dict[str, torch.Tensor]] = None,
                   extra: Optional[Dict[str, torch.Tensor]] = None):
        """critic_type != "Transformer": vn/DRL.py:388-434 (and :237-299 when ``extra`` is given) with the ``QNetwork``
        critic, through the same single C call as the Transformer critic (``dgvit_sac_update`` with cfg.kind ==
        DGVIT_QNET for the critic slots).  ``noise`` (tests): eps_next / eps_pi rsample draws and mask_a_next / mask_a
        keep-masks of the two actor passes (+ eps_x / mask_x for the imitation rows; the CNN critic has no dropout).
        ``extra``: obs, pobs, target, weight of the ``learn_guidence`` imitation rows."""
        view = lambda t: t.reshape(t.shape[0], -1)
        b = {k: view(batch[k]) for k in ("obs", "next_obs", "pobs", "next_pobs", "act", "rew")}
        if "done" in batch:
            b["done"] = batch["done"]
        nz = None
        if noise is not None:
            nz = {k: noise.get(k) for k in ("eps_next", "eps_pi", "mask_a_next", "mask_a")}
            nz["mask_c"] = nz["mask_a"]        # (selects the injected-mask mode; the CNN critic itself draws nothing)
        ex = None
        if extra is not None:
            b["obs"] = torch.cat([b["obs"], view(extra["obs"])])
            b["pobs"] = torch.cat([b["pobs"], extra["pobs"]])
            ex = dict(target=extra["target"], weight=extra["weight"])
            if nz is not None:
                nz["eps_pi"] = torch.cat([nz["eps_pi"], noise["eps_x"]])
                nz["mask_a"] = nz["mask_c"] = torch.cat([nz["mask_a"], noise["mask_x"]])
        if nz is not None:
            nz = {k: (v.to(torch.uint8) if k.startswith("mask") else v).contiguous() for k, v in nz.items()}
        b = {k: v.contiguous() for k, v in b.items()}
        losses = self.update_from_batch(b, nz, extra=ex).tolist()
        self.alpha = float(self._alpha.item()) if self.automatic_entropy_tuning else self.alpha
        return losses[0], losses[1]

    # ------------------------------------------------------------------ data parallel without NCCL calls
    def _setup_fused_dp(self, keep_reduced: bool = False):
        """Both gradient arenas move into one symmetric-memory buffer ([critic | actor]); the Adam kernels then read the
        rank-sum of the gradients themselves (NVLS multimem.ld_reduce, or peer loads) behind a flag barrier: the whole
        data-parallel update is one C call / one CUDA graph (include/dgvit.h: dgvit_dp).  Falls back to the NCCL path
        (all-reduce between the phases) when symmetric memory is not available."""
        try:
            import torch.distributed._symmetric_memory as symm
            dev = self.device
            Lc, La = int(self.critic.layout().total), int(self.policy.layout().total)
            buf = symm.empty(Lc + La, dtype=torch.float32, device=dev)
            hdl = symm.rendezvous(buf, torch.distributed.group.WORLD.group_name)
            buf.zero_()
            mc = 0
            try:
                mc = int(hdl.multicast_ptr or 0)
            except Exception:
                mc = 0
            if os.environ.get("DGVIT_DP_MULTICAST", "1") != "1":
                mc = 0
            i32 = dict(dtype=torch.int32, device=dev)
            st = dict(buf=buf, hdl=hdl, Lc=Lc, La=La, tail=torch.zeros(64, device=dev), finished=torch.zeros(2, **i32),
                      err=torch.zeros(1, **i32), multicast=mc,
                      reduced=[torch.zeros(Lc, device=dev), torch.zeros(La, device=dev)] if keep_reduced else None)
            st["struct"] = L.Dp(world=hdl.world_size, rank=hdl.rank, multicast=mc or None, peers=int(hdl.buffer_ptrs_dev),
                                pads=int(hdl.signal_pad_ptrs_dev), arena_off=(C.c_int64 * 2)(0, Lc), tail=st["tail"].data_ptr(),
                                reduced_out=(L.c_f_p * 2)(*( [t.data_ptr() for t in st["reduced"]] if keep_reduced else [None, None])),
                                finished=st["finished"].data_ptr(), error_flag=st["err"].data_ptr())
            torch.distributed.barrier()
            self._dp = st
            self._point_grads_at_symmetric_buffer()
        except Exception as e:       # no symmetric memory on this system: NCCL all-reduce between the phases
            import sys
            print(f"dgvit_b200: fused data-parallel optimizer unavailable ({e!r}); using NCCL all-reduce", file=sys.stderr)
            self._dp = None

    def _point_grads_at_symmetric_buffer(self):
        st = self._dp
        for mod, a, b in ((self.critic, 0, st["Lc"]), (self.policy, st["Lc"], st["Lc"] + st["La"])):
            mod.bind()
            if mod._garena.data_ptr() != st["buf"].data_ptr() + 4 * a:
                mod._garena = st["buf"][a:b]

    def reduced_gradients(self):
        """(critic, actor) gradient arenas summed over the ranks as the last update saw them."""
        if self._dp is not None:
            if self._dp["reduced"] is None:
                raise RuntimeError("construct the fused data-parallel state with keep_reduced=True to read reduced gradients")
            return self._dp["reduced"][0], self._dp["reduced"][1]
        return self.critic._garena, self.policy._garena

    # ------------------------------------------------------------------ plumbing
    @property
    def world(self):
        return torch.distributed.get_world_size() if self.distributed else 1

    @property
    def rank(self):
        return torch.distributed.get_rank() if self.distributed else 0

    def _sac_struct(self, B_local: int, B_global: int, offset: int, n_extra: int = 0) -> L.Sac:
        def adam(name, lr):
            o = self._opt[name]
            return L.Adam(m=o["m"].data_ptr(), v=o["v"].data_ptr(), step=o["step"].data_ptr(), lr=lr, beta1=0.9,
                          beta2=0.999, eps=1e-8)
        if self._dp is not None:
            self._point_grads_at_symmetric_buffer()
        return L.Sac(dp=C.pointer(self._dp["struct"]) if self._dp is not None else None,
                     actor=self.policy.net_struct(), critic=self.critic.net_struct(),
                     critic_target=self.critic_target.net_struct(), actor_opt=adam("actor", self.lr_a),
                     critic_opt=adam("critic", self.lr_c), log_alpha=self.log_alpha.data_ptr(),
                     alpha=self._alpha.data_ptr(), alpha_m=self._alpha_m.data_ptr(), alpha_v=self._alpha_v.data_ptr(),
                     alpha_step=self._alpha_step.data_ptr(), lr_alpha=self.lr_alpha,
                     auto_alpha=int(self.automatic_entropy_tuning), target_entropy=self.target_entropy,
                     gamma=self.gamma, tau=self.tau, do_polyak=int(self.itera % self.policy_freq == 0),
                     precision={"fp32": L.FP32, "bf16": L.BF16}[self.precision], global_batch=B_global,
                     n_extra=n_extra, sample_offset=offset, rng_state=self._rng.data_ptr(), action_scale=self._scale.data_ptr(),
                     action_bias=self._bias.data_ptr())

    def _workspace(self, B: int, n_extra: int = 0) -> torch.Tensor:
        n = C.c_size_t()
        prec = {"fp32": L.FP32, "bf16": L.BF16}[self.precision]
        query = L.lib().dgvit_sac_qnet_workspace_bytes if self._cnn_critic else L.lib().dgvit_sac_workspace_bytes
        L.check(query(C.byref(self.policy._cfg), B, n_extra, prec, C.byref(n)), "sac_workspace")
        if self._ws is None or self._ws.numel() < n.value:
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=self.device)
        return self._ws

    def _batch_buffers(self, B: int) -> Dict[str, torch.Tensor]:
        """Minibatch, index and pinned index-staging buffers of one batch size.  They are kept per batch size for the
        life of the agent: a captured CUDA graph has their addresses baked in, so alternating batch sizes must find the
        buffers their graphs were captured with."""
        if self._batch is None:
            self._batch = {}
        ent = self._batch.get(B)
        if ent is None:
            f = self.replay_buffer.obs.shape[1]
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
            ent = dict(obs=z(B, f), next_obs=z(B, f), pobs=z(B, self.pstate_dim), next_pobs=z(B, self.pstate_dim),
                       act=z(B, self.action_dim), rew=z(B, 1), done=z(B, 1))
            # ring of pinned staging buffers: the host may run several steps ahead of the GPU, and a slot is rewritten
            # only after the H2D copy that last read it has completed
            ent["_idx"] = torch.zeros(B, dtype=torch.int64, device=self.device)
            ent["_ring"] = [(torch.zeros(B, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            ent["_slot"] = 0
            self._batch[B] = ent
        return ent

    # ------------------------------------------------------------------ the update
    def update_from_batch(self, batch: Dict[str, torch.Tensor], noise: Optional[Dict[str, torch.Tensor]] = None,
                          debug: Optional[torch.Tensor] = None, global_batch: Optional[int] = None,
                          sample_offset: Optional[int] = None, _phases=None,
                          extra: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """One SAC update on device tensors (no host sync).  ``noise`` injects the stochastic
        inputs (parity tests): eps_next, eps_pi [B,na] and keep-masks mask_* [B,N,D] uint8.
        ``extra`` (learn_guidence): dict(target [n,na], weight [n]); then batch["obs"] / ["pobs"] carry
        B + n rows, the last n being imitation rows seen only by the actor.
        Returns the device tensor [qf1_loss, policy_loss, qf2_loss, alpha_loss]."""
        B = batch["next_obs"].shape[0]
        n_extra = 0 if extra is None else int(extra["target"].shape[0])
        assert batch["obs"].shape[0] == B + n_extra and batch["pobs"].shape[0] == B + n_extra
        Bg = global_batch if global_batch is not None else B * self.world
        if sample_offset is None:
            sample_offset = self.rank * B
        ws = self._workspace(B, n_extra)
        s = self._sac_struct(B, Bg, sample_offset, n_extra)
        bt = L.Batch(**{k: batch[k].data_ptr() for k in ("obs", "next_obs", "pobs", "next_pobs", "act", "rew")},
                     done=batch["done"].data_ptr() if "done" in batch else None,
                     extra_target=None if extra is None else extra["target"].data_ptr(),
                     extra_weight=None if extra is None else extra["weight"].data_ptr())
        nz = None
        if noise is not None:
            nz = L.Noise(**{k: L.ptr(noise.get(k)) for k in ("eps_next", "eps_pi", "mask_a_next", "mask_ct", "mask_c",
                                                              "mask_a", "mask_c_pi")},
                         drop_mode=L.DROP_MASK if noise.get("mask_c") is not None else
                         (L.DROP_NONE if noise.get("no_dropout") else L.DROP_RNG))
        losses = self._loss_buffer()
        out = L.SacOut(losses=self._loss_out().data_ptr(), debug=L.ptr(debug))
        lib = L.lib()
        nzp = C.byref(nz) if nz is not None else None
        keep = (s, bt, nz, out, ws, batch, noise, debug, extra)      # ctypes structs must outlive the calls

        def phase(which):
            st = _stream(self.device)
            if which == 0:
                L.check(lib.dgvit_sac_update(C.byref(s), C.byref(bt), nzp, C.byref(out), B, ws.data_ptr(), ws.numel(), st),
                        "sac_update")
            elif which == 1:
                L.check(lib.dgvit_sac_phase1(C.byref(s), C.byref(bt), nzp, C.byref(out), B, ws.data_ptr(), ws.numel(), st),
                        "sac_phase1")
            elif which == 2:
                L.check(lib.dgvit_sac_phase2(C.byref(s), C.byref(bt), nzp, C.byref(out), B, ws.data_ptr(), ws.numel(), st),
                        "sac_phase2")
            else:
                L.check(lib.dgvit_sac_phase3(C.byref(s), B, ws.data_ptr(), ws.numel(), st), "sac_phase3")
            return keep

        if _phases is not None:          # caller (graph capture of the data-parallel path) drives the phases itself
            return phase
        if not self.distributed or self.world == 1 or self._dp is not None:
            phase(0)        # (fused data parallel: the Adam kernels all-reduce the gradients themselves)
        else:
            phase(1)
            allreduce_sum_(self.critic._garena)
            phase(2)
            allreduce_sum_(self.policy._garena)      # carries d(alpha_loss)/d(log_alpha) and the four loss sums in its tail
            phase(3)
        self.itera += 1
        return losses

    def _loss_buffer(self) -> torch.Tensor:
        """[qf1_loss, policy_loss, qf2_loss, alpha_loss].  Data parallel: four floats in the padded tail of the ACTOR
        gradient arena (next to the alpha-gradient slot; Adam and the unused-gradient memsets skip that range), so the
        actor's gradient all-reduce also sums the per-rank loss shares: no separate 16-byte collective."""
        if self._dp is not None:
            return self._dp["tail"][1:5]          # written by the actor's fused all-reduce + Adam pass
        return self._loss_out()

    def _loss_out(self) -> torch.Tensor:
        """Where the loss kernels write their per-rank shares."""
        if self.distributed and self.world > 1:
            if self._dp is not None:
                self._point_grads_at_symmetric_buffer()
            self.policy.bind()
            slot = int(self.policy.layout().alpha_grad_slot)
            return self.policy._garena[slot + 1: slot + 5]
        return self._losses

    def close(self):
        """Drop captured CUDA graphs (call before ``torch.distributed.destroy_process_group``: a graph that holds
        NCCL kernels must not outlive its communicator)."""
        self._graphs = {}
        if torch.cuda.is_available():
            torch.cuda.synchronize(self.device)

    def learn(self, batch_size=64):
        """vn/DRL.py:373-437 — returns (qf1_loss, policy_loss) python floats (one D2H read)."""
        losses = self.learn_async(batch_size)
        l = losses.tolist()
        self.alpha = float(self._alpha.item()) if self.automatic_entropy_tuning else self.alpha
        return l[0], l[1]

    def learn_async(self, batch_size=64) -> torch.Tensor:
        """Same update without the host read-back (losses stay on the device)."""
        B = int(batch_size)
        idx_host = self.replay_buffer.sample_indexes(B)
        batch = self._batch_buffers(B)
        pin, ev = batch["_ring"][batch["_slot"]]
        batch["_slot"] = (batch["_slot"] + 1) % len(batch["_ring"])
        ev.synchronize()
        pin.copy_(idx_host)
        batch["_idx"].copy_(pin, non_blocking=True)
        ev.record(torch.cuda.current_stream(self.device))
        return self._run(("learn", B), batch, gather=True)

    def update_from_batch_graphed(self, batch: Dict[str, torch.Tensor], key) -> torch.Tensor:
        """``update_from_batch`` replayed from CUDA graph(s) keyed by ``key`` (the batch tensors must be
        the same device buffers on every call with that key)."""
        return self._run(("batch", key, batch["obs"].data_ptr()), batch, gather=False)

    def _capture_stream(self):
        """Capture stream of the update graph.  Kernel nodes inherit its priority; measured (profiles/r1_14): giving
        the dependency chain a HIGHER priority than the library's forked streams loses all of their overlap (119.8 k vs
        126.8 k samples/s), so the default is the same priority for all."""
        if getattr(self, "_cap_stream", None) is None:
            prio = int(os.environ.get("DGVIT_GRAPH_PRIO", "0"))
            self._cap_stream = torch.cuda.Stream(device=self.device, priority=prio)
        return self._cap_stream

    def _run(self, key, batch, gather, extra=None) -> torch.Tensor:
        """Eager on the first call with a key (warms lazily initialised state), captured on the second,
        replayed afterwards.  Every launch, tensor map and device pointer of the update is frozen in the
        graph; per-step state (Adam step counts, alpha, RNG counter, sampled indexes) lives in device
        memory.  Single GPU: one graph for gather + update.  Data parallel: one graph per phase with the
        two NCCL gradient all-reduces issued between them.  ``gather``: True = the replay gather of ``learn``, a callable =
        the minibatch assembly of ``learn_guidence`` (gathers + row copies); either is part of the graph."""
        dp = self.distributed and self.world > 1 and self._dp is None
        if gather is True:
            pre = lambda: self.replay_buffer.gather(batch["_idx"], batch)
        else:
            pre = gather if callable(gather) else (lambda: None)
        if not self.use_cuda_graph:
            pre()
            return self.update_from_batch(batch, extra=extra)
        # a graph freezes every device pointer: re-bound arenas (.to(), load_state_dict(assign=True)) start new entries
        key = key + (int(self.itera % self.policy_freq == 0), self.policy.net_struct().params,
                     self.critic.net_struct().params, self.critic_target.net_struct().params)
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs[key] = "warm"
            pre()
            return self.update_from_batch(batch, extra=extra)
        if ent == "warm":
            torch.cuda.synchronize(self.device)
            phase = self.update_from_batch(batch, _phases=True, extra=extra)
            graphs = []
            # Data parallel: either one graph per phase with the NCCL all-reduces issued eagerly between them, or
            # (DGVIT_GRAPH_NCCL=1) the collectives captured too, one graph per update; `close()` must then run before the
            # process group is destroyed.
            one = dp and self._graph_nccl
            for i, which in enumerate((1, 2, 3) if (dp and not one) else (0,)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._capture_stream()):
                    if i == 0:
                        pre()
                    if one:
                        keep = phase(1)
                        allreduce_sum_(self.critic._garena)
                        phase(2)
                        allreduce_sum_(self.policy._garena)
                        phase(3)
                    else:
                        keep = phase(which)
                graphs.append(g)
            ent = self._graphs[key] = (graphs, keep, self._loss_buffer())
        graphs = ent[0]
        if dp and len(graphs) == 3:
            graphs[0].replay()
            allreduce_sum_(self.critic._garena)
            graphs[1].replay()
            allreduce_sum_(self.policy._garena)
            graphs[2].replay()
        else:
            graphs[0].replay()
        self.itera += 1
        return ent[2]

    def learn_guidence(self, engage, batch_size=64):
        """vn/DRL.py:187-301 — the update the shipped config runs (PRE_BUFFER): agent + expert minibatch
        for the critic / policy losses, plus the guidance (expert rows) and engage (rows with engage == 1)
        imitation losses on the actor's tanh-mean.  All imitation rows ride in the same actor pass as
        extra rows with per-row loss weights, so the update stays one fused call."""
        B = int(batch_size)
        rb, re = self.replay_buffer, (self.replay_buffer_expert if self.pre_buffer else None)
        Be = 0
        if re is not None and rb.get_stored_size() > 0:
            Be = int(min(np.floor(re.get_stored_size() / rb.get_stored_size() * B), B))      # :193-196
        idx_a = rb.sample_indexes(B)
        eng_rows = np.nonzero(rb.engage_host[idx_a.numpy()] == 1)[0]                         # :266
        if self.use_cuda_graph:
            return self._learn_guidence_graphed(B, Be, idx_a, eng_rows)
        Bc, n_extra = B + Be, Be + len(eng_rows)
        f = rb.obs.shape[1]
        dev = self.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        key = (Bc, n_extra)
        buf = self._gbuf.get(key)
        if buf is None:
            buf = dict(obs=z(Bc + n_extra, f), next_obs=z(Bc, f), pobs=z(Bc + n_extra, self.pstate_dim),
                       next_pobs=z(Bc, self.pstate_dim), act=z(Bc, self.action_dim), rew=z(Bc, 1), done=z(Bc, 1),
                       target=z(max(n_extra, 1), self.action_dim), weight=z(max(n_extra, 1)))
            self._gbuf = {key: buf}
        rows = lambda t, a, b: t[a:b]
        rb.gather(idx_a.to(dev), {k: rows(buf[k], 0, B) for k in ("obs", "next_obs", "pobs", "next_pobs", "act", "rew", "done")})
        if Be > 0:
            idx_e = re.sample_indexes(Be).to(dev)
            re.gather(idx_e, {k: rows(buf[k], B, Bc) for k in ("obs", "next_obs", "pobs", "next_pobs", "act", "rew", "done")})
            # guidance rows = the expert minibatch again (its own dropout / rsample draws), :259-263
            buf["obs"][Bc:Bc + Be].copy_(buf["obs"][B:Bc])
            buf["pobs"][Bc:Bc + Be].copy_(buf["pobs"][B:Bc])
            buf["target"][:Be].copy_(buf["act"][B:Bc])
            buf["weight"][:Be].fill_(self.guidence_weight / (Be * self.action_dim * self.world))
        if len(eng_rows) > 0:                                                                # :267-273
            er = torch.as_tensor(eng_rows, device=dev)
            buf["obs"][Bc + Be:].copy_(buf["obs"][er])
            buf["pobs"][Bc + Be:].copy_(buf["pobs"][er])
            buf["target"][Be:n_extra].copy_(buf["act"][er])
            buf["weight"][Be:n_extra].fill_(self.engage_weight / (len(eng_rows) * self.action_dim * self.world))
        extra = None if n_extra == 0 else dict(target=buf["target"][:n_extra], weight=buf["weight"][:n_extra])
        losses = self.update_from_batch({k: buf[k] for k in ("obs", "next_obs", "pobs", "next_pobs", "act", "rew", "done")},
                                        extra=extra).tolist()
        self.alpha = float(self._alpha.item()) if self.automatic_entropy_tuning else self.alpha
        return losses[0], losses[1]

    def _learn_guidence_graphed(self, B: int, Be: int, idx_a: torch.Tensor, eng_rows: np.ndarray):
        """``learn_guidence`` replayed from a CUDA graph.  The number of engaged rows changes from call to call; a graph
        needs fixed shapes, so the engage rows are padded to the next multiple of 32 with rows of weight 0 (a zero-weight
        imitation row contributes exact zeros to the loss and to every gradient).  One graph per (B, Be, padded count):
        sampled indexes, engaged-row indexes and the per-row weights arrive through pinned staging buffers; the two
        gathers, the row copies and the update are graph nodes."""
        rb, re = self.replay_buffer, self.replay_buffer_expert
        dev = self.device
        n_e = len(eng_rows)
        n_pad = 0 if n_e == 0 else min(B, (n_e + 31) // 32 * 32)
        Bc, n_extra = B + Be, Be + n_pad
        key = ("guidence", B, Be, n_pad)
        st = self._gbuf.get(key)
        if st is None:
            f = rb.obs.shape[1]
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
            i64 = lambda n: torch.zeros(max(n, 1), dtype=torch.int64, device=dev)
            st = dict(obs=z(Bc + n_extra, f), next_obs=z(Bc, f), pobs=z(Bc + n_extra, self.pstate_dim),
                      next_pobs=z(Bc, self.pstate_dim), act=z(Bc, self.action_dim), rew=z(Bc, 1), done=z(Bc, 1),
                      target=z(max(n_extra, 1), self.action_dim), weight=z(max(n_extra, 1)),
                      idx_a=i64(B), idx_e=i64(Be), er=i64(n_pad),
                      ring=[dict(idx_a=torch.zeros(B, dtype=torch.int64).pin_memory(),
                                 idx_e=torch.zeros(max(Be, 1), dtype=torch.int64).pin_memory(),
                                 er=torch.zeros(max(n_pad, 1), dtype=torch.int64).pin_memory(),
                                 weight=torch.zeros(max(n_extra, 1)).pin_memory(), ev=torch.cuda.Event()) for _ in range(4)],
                      slot=0)
            self._gbuf[key] = st
            # Be follows the fill ratio of the two stores and the padded count the data: keep the eight most recent shapes
            gk = [k for k in self._gbuf if isinstance(k, tuple) and k and k[0] == "guidence"]
            for old_key in gk[:-8]:
                torch.cuda.synchronize(dev)
                del self._gbuf[old_key]
                for gkey in [k for k in self._graphs if k[:len(old_key)] == old_key]:
                    del self._graphs[gkey]
        pin = st["ring"][st["slot"]]
        st["slot"] = (st["slot"] + 1) % len(st["ring"])
        pin["ev"].synchronize()                      # the copies that last read this staging slot have completed
        pin["idx_a"].copy_(idx_a)
        if Be > 0:
            pin["idx_e"][:Be].copy_(re.sample_indexes(Be))
        w = pin["weight"]
        w.zero_()
        if Be > 0:
            w[:Be] = self.guidence_weight / (Be * self.action_dim * self.world)
        if n_e > 0:
            pin["er"].zero_()
            pin["er"][:n_e].copy_(torch.from_numpy(np.ascontiguousarray(eng_rows, dtype=np.int64)))
            w[Be:Be + n_e] = self.engage_weight / (n_e * self.action_dim * self.world)
        st["idx_a"].copy_(pin["idx_a"], non_blocking=True)
        if Be > 0:
            st["idx_e"][:Be].copy_(pin["idx_e"][:Be], non_blocking=True)
        if n_pad > 0:
            st["er"][:n_pad].copy_(pin["er"][:n_pad], non_blocking=True)
        st["weight"].copy_(pin["weight"], non_blocking=True)
        pin["ev"].record(torch.cuda.current_stream(dev))
        fields = ("obs", "next_obs", "pobs", "next_pobs", "act", "rew", "done")

        def assemble():
            rb.gather(st["idx_a"], {k: st[k][0:B] for k in fields})
            if Be > 0:
                re.gather(st["idx_e"][:Be], {k: st[k][B:Bc] for k in fields})
                st["obs"][Bc:Bc + Be].copy_(st["obs"][B:Bc])            # guidance rows = the expert minibatch again, :259-263
                st["pobs"][Bc:Bc + Be].copy_(st["pobs"][B:Bc])
                st["target"][:Be].copy_(st["act"][B:Bc])
            if n_pad > 0:                                                # engaged rows, :267-273 (padding rows: row 0, weight 0)
                torch.index_select(st["obs"][:Bc], 0, st["er"][:n_pad], out=st["obs"][Bc + Be:])
                torch.index_select(st["pobs"][:Bc], 0, st["er"][:n_pad], out=st["pobs"][Bc + Be:])
                torch.index_select(st["act"][:Bc], 0, st["er"][:n_pad], out=st["target"][Be:n_extra])

        extra = None if n_extra == 0 else dict(target=st["target"][:n_extra], weight=st["weight"][:n_extra])
        losses = self._run(key, {k: st[k] for k in fields}, assemble, extra=extra).tolist()
        self.alpha = float(self._alpha.item()) if self.automatic_entropy_tuning else self.alpha
        return losses[0], losses[1]

    def initialize_expert_buffer(self, s, a_exp, ps, ps_, r, s_, d=0):
        """vn/DRL.py:469-477."""
        self.replay_buffer_expert.add(obs=s, act=a_exp, pobs=ps, next_pobs=ps_, rew=r, next_obs=s_, done=d)

    def load_demonstrations(self, files):
        """vn/main.py:232-266: concatenate the demonstration episodes (``.npz`` written by vn/demonstration.py:237-245:
        obs, act, goal, reward, next_obs, next_goal, done) and put them into the expert buffer in one batched append."""
        keys = ("obs", "act", "goal", "reward", "next_obs", "next_goal", "done")
        cols = {k: [] for k in keys}
        for fn in files:
            with np.load(fn) as d:
                for k in keys:
                    cols[k].append(np.array(d[k]))
        c = {k: np.concatenate(v, axis=0) for k, v in cols.items()}
        if self.replay_buffer_expert is None or self.replay_buffer_expert.size < c["obs"].shape[0]:
            self.buffer_size_expert = c["obs"].shape[0] + 1
            self.replay_buffer_expert = ReplayStore(self.buffer_size_expert, self.replay_buffer.obs_shape, self.action_dim,
                                                    self.pstate_dim, self.device, self.seed + 1)
            self.pre_buffer = True
        self.initialize_expert_buffer(c["obs"], c["act"], c["goal"][:, :2], c["next_goal"][:, :2], c["reward"],
                                      c["next_obs"], c["done"])
        return c["obs"].shape[0]

    # ------------------------------------------------------------------ act
    def choose_action(self, istate, pstate, evaluate=False):
        """vn/DRL.py:170-185."""
        return self.policy.choose_action(istate, pstate, evaluate)

    # ------------------------------------------------------------------ replay write path
    def store_transition(self, s, a, ps, ps_, r, s_, engage, a_exp, d=0):
        """vn/DRL.py:449-467."""
        self.replay_buffer.add(obs=s, act=a if a is not None else a_exp, pobs=ps, next_pobs=ps_, rew=r, next_obs=s_,
                               engage=engage, done=d)

    def save_transition(self, output, timeend=0):
        """vn/DRL.py:505-506."""
        self.replay_buffer.save_transitions(file="{}/{}".format(output, timeend))

    def load_transition(self, output):
        """vn/DRL.py:508-510."""
        if output is None:
            return
        self.replay_buffer.load_transitions("{}.npz".format(output))

    # ------------------------------------------------------------------ checkpoints (vn/DRL.py:480-503)
    def load_model(self, output):
        if output is None:
            return
        self.policy.load_state_dict(torch.load("{}/actor.pkl".format(output)))
        self.critic.load_state_dict(torch.load("{}/critic.pkl".format(output)))
        self._after_load()

    def save_model(self, output):
        torch.save(self.policy.state_dict(), "{}/actor.pkl".format(output))
        torch.save(self.critic.state_dict(), "{}/critic.pkl".format(output))

    def save(self, filename, directory, reward, seed, nb_col=100):
        torch.save(self.policy.state_dict(), "%s/%s_reward_%s_nbCol_%s_seed_%s_actor.pth" % (directory, filename, reward, nb_col, seed))
        torch.save(self.critic.state_dict(), "%s/%s_reward_%s_nbCol_%s_seed_%s_critic.pth" % (directory, filename, reward, nb_col, seed))

    def load(self, filename, directory):
        self.policy.load_state_dict(torch.load("%s/%s_actor.pth" % (directory, filename)))
        self.critic.load_state_dict(torch.load("%s/%s_critic.pth" % (directory, filename)))
        self._after_load()

    def load_target(self):
        self.critic_target.bind()
        self.critic_target._arena.copy_(self.critic._arena)
        if hasattr(self.critic_target, "refresh_shadow"):
            self.critic_target.refresh_shadow()

    def load_actor(self, filename, directory):
        self.policy.load_state_dict(torch.load("%s/%s_actor.pth" % (directory, filename)))
        self._after_load()

    def _after_load(self):
        for m in (self.critic, self.critic_target, self.policy):
            m.bind()
            if hasattr(m, "refresh_shadow"):
                m.refresh_shadow()
