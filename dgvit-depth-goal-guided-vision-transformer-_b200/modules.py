"""nn.Module surface of the reference networks, backed by libdgvit.so.

Mirrors (same constructor signatures, attribute names, parameter registration
order and ``state_dict`` keys):

* ``GoT``          vn/GoalFormer.py:123-171
* ``GoTPolicy``    vn/got_sac_network.py:172-256
* ``GoTQNetwork``  vn/got_sac_network.py:75-123

The sub-modules (``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Conv2d``) are parameter
containers only: they are created in the reference's order so that the same
``torch.manual_seed`` gives bit-identical initial weights, then every parameter is
re-pointed (``p.data``) into one flat fp32 arena whose geometry the C library
defines (``dgvit_param_layout``).  All arithmetic happens in hand-written CUDA
kernels reached through the C ABI; there is no PyTorch / CPU fallback.
"""
from __future__ import annotations

import os
import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L

LOG_SIG_MAX = 2
LOG_SIG_MIN = -20
epsilon = 1e-6


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


def set_seed(seed):
    """vn/got_sac_network.py:22-27."""
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def weights_init_(m):
    """vn/got_sac_network.py:30-33."""
    if isinstance(m, nn.Linear):
        torch.nn.init.xavier_uniform_(m.weight, gain=1)


# --------------------------------------------------------------------------- containers
class _Patchify(nn.Module):
    """Placeholder for einops ``Rearrange('b (h p1) (w p2) -> b (h w) (p1 p2)')`` so that the
    patch Linear keeps the key ``to_patch_embedding.1`` (vn/GoalFormer.py:137-139).  In the bf16 path the
    rearrangement is the shared-memory address of the patch-embedding kernels (``csrc/patch_tc.cuh``: frame rows ->
    swizzled UMMA tiles, forward and weight gradient; no patch matrix in HBM); the fp32 parity path materialises the
    patch matrix with ``patchify_kernel``."""

    def __init__(self, p1, p2):
        super().__init__()
        self.p1, self.p2 = p1, p2


class RMSNorm(nn.Module):
    def __init__(self, dim, unit_offset=False):
        super().__init__()
        self.unit_offset = unit_offset
        self.scale = dim ** 0.5
        self.g = nn.Parameter(torch.zeros(dim))
        nn.init.constant_(self.g, 1.0 - float(unit_offset))


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner_dim = dim_head * heads
        if heads == 1 and dim_head == dim:
            # the reference drops the output projection there (nn.Identity, vn/GoalFormer.py:55,67-70): another network
            # (no to_out parameters, other state_dict keys, other RNG consumption at init) than the kernels implement
            raise NotImplementedError("heads == 1 with dim_head == dim (no output projection) is not on the accelerated path")
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout)),
            ]))


class GoT(nn.Module):
    """The DGViT trunk (vn/GoalFormer.py:123-171).  Inside ``GoTPolicy`` / ``GoTQNetwork`` the owner runs it fused with
    its heads (one C call); ``GoT.forward(img, goal)`` on its own goes through ``dgvit_trunk_forward`` on the owner's
    arena (or a private arena for a stand-alone trunk)."""

    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, pool="cls",
                 channels=3, dim_head=64, dropout=0.0, emb_dropout=0.1):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, \
            "Image dimensions must be divisible by the patch size."
        assert pool in {"cls", "mean"}
        if pool != "cls":
            raise NotImplementedError("dgvit_b200 implements pool='cls' (the only mode the reference uses)")
        if dropout != 0.0:
            raise NotImplementedError("dgvit_b200 implements dropout=0 inside the blocks (reference default)")
        self.image_size = (image_height, image_width)
        self.patch_size = (patch_height, patch_width)
        self.dim, self.depth, self.heads, self.dim_head, self.mlp_dim = dim, depth, heads, dim_head, mlp_dim
        self.emb_dropout = emb_dropout
        num_patches = (image_height // patch_height) * (image_width // patch_width)
        self.layer_norm = RMSNorm(dim)
        # the reference hard-codes p1=16, p2=20 and Linear(320, dim)
        self.to_patch_embedding = nn.Sequential(_Patchify(patch_height, patch_width),
                                                nn.Linear(patch_height * patch_width, dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))

    def _set_owner(self, owner):
        import weakref
        object.__setattr__(self, "_owner_ref", weakref.ref(owner))

    def _backend(self):
        """The arena module that runs this trunk: the owning GoTPolicy / GoTQNetwork, else a private one."""
        ref = self.__dict__.get("_owner_ref")
        owner = ref() if ref is not None else None
        if owner is None:
            owner = self.__dict__.get("_standalone")
            if owner is None:
                owner = _StandaloneTrunk(self)
                object.__setattr__(self, "_standalone", owner)      # not a registered sub-module (no cycle)
        return owner

    def forward(self, img, goal):
        """vn/GoalFormer.py:156-171: img [B,H,W] + goal token [B,dim] -> z [B,dim] (``dgvit_trunk_forward``)."""
        be = self._backend()
        be.bind()
        _require_cuda(be._arena)
        img = be._check_img(img)
        goal = goal.to(img.device, torch.float32)
        if goal.dim() != 2 or goal.shape[0] != img.shape[0] or goal.shape[1] != self.dim:
            raise ValueError(f"goal must be [B,{self.dim}], got {tuple(goal.shape)}")
        ps_ = list(self.parameters())
        need_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in ps_) or goal.requires_grad)
        return _TrunkFn.apply(be, self, need_grad, img, goal, *ps_)


# --------------------------------------------------------------------------- arena binding
def _trunk_offsets(lay: L.Layout, depth: int, pre: str = "trans.") -> List[Tuple[str, int]]:
    out = [(pre + "pos_embedding", lay.pos), (pre + "cls_token", lay.cls), (pre + "layer_norm.g", lay.rms_g),
           (pre + "to_patch_embedding.1.weight", lay.patch_w), (pre + "to_patch_embedding.1.bias", lay.patch_b)]
    for l in range(depth):
        b = lay.block[l]
        p = f"{pre}transformer.layers.{l}."
        out += [(p + "0.norm.weight", b.ln1_w), (p + "0.norm.bias", b.ln1_b), (p + "0.fn.to_qkv.weight", b.qkv_w),
                (p + "0.fn.to_out.0.weight", b.out_w), (p + "0.fn.to_out.0.bias", b.out_b),
                (p + "1.norm.weight", b.ln2_w), (p + "1.norm.bias", b.ln2_b),
                (p + "1.fn.net.0.weight", b.fc1_w), (p + "1.fn.net.0.bias", b.fc1_b),
                (p + "1.fn.net.3.weight", b.fc2_w), (p + "1.fn.net.3.bias", b.fc2_b)]
    out += [(pre + "mlp_head.0.weight", lay.mlp_head_ln_w), (pre + "mlp_head.0.bias", lay.mlp_head_ln_b),
            (pre + "mlp_head.1.weight", lay.mlp_head_w), (pre + "mlp_head.1.bias", lay.mlp_head_b)]
    return out


class _ArenaModule(nn.Module):
    """Shared machinery: flat arenas, C structs, workspaces, noise injection."""

    KIND = -1

    def _init_backend(self, nb_actions, nb_pstate, block, head, l_f_size, image_size=(128, 160),
                      patch_size=(16, 20), mlp_dim=2048, dim_head=64):
        self._cfg = L.Cfg(kind=self.KIND, img_h=image_size[0], img_w=image_size[1], patch_h=patch_size[0],
                          patch_w=patch_size[1], dim=l_f_size, depth=block, heads=head, dim_head=dim_head,
                          mlp_dim=mlp_dim, n_act=nb_actions, n_pstate=nb_pstate)
        self._layout: Optional[L.Layout] = None
        self._arena = None       # fp32 params
        self._garena = None      # fp32 grads
        self._shadow = None      # bf16 params
        self._ws_cache: Dict = {}
        self._noise_fifo: List[dict] = []
        self._rng_state = None
        self._act_state = None
        self.precision = "fp32"  # "fp32" | "bf16"

    # ---- layout
    def layout(self) -> L.Layout:
        if self._layout is None:
            self._layout = L.layout_of(self._cfg)
        return self._layout

    def _named_offsets(self) -> List[Tuple[str, int]]:
        raise NotImplementedError

    def _unused(self, name: str) -> bool:
        return (name.endswith("cls_token") or "mlp_head." in name or name.startswith("conv"))

    # ---- binding
    def _bound(self) -> bool:
        """Are the first and last parameters still views of the arena?  (`.to()`, `load_state_dict(assign=True)` etc. re-point
        `p.data`.)  This runs on every batch-1 `choose_action`, so it probes two cached Parameter objects instead of walking
        the module tree (that walk was ~100 us of the 0.24 ms act latency)."""
        if self._arena is None:
            return False
        probe = getattr(self, "_bind_probe", None)
        if probe is None or probe[0][0] is not next(self.parameters()):
            ps = dict(self.named_parameters())
            offs = self._named_offsets()
            probe = self._bind_probe = [(ps[offs[0][0]], offs[0][1]), (ps[offs[-1][0]], offs[-1][1])]
        base = self._arena.data_ptr()
        for p, off in probe:
            if p.device != self._arena.device or p.data_ptr() != base + 4 * off:
                return False
        return True

    def bind(self, force: bool = False):
        """Alias every parameter into the flat arena on its current device."""
        if not force and self._bound():
            return self
        lay = self.layout()
        ps = dict(self.named_parameters())
        dev = next(iter(ps.values())).device
        offs = self._named_offsets()
        assert [n for n, _ in offs] == list(ps.keys()), "parameter registration order differs from the C layout"
        arena = torch.zeros(lay.total, dtype=torch.float32, device=dev)
        for name, off in offs:
            p = ps[name]
            n = p.numel()
            assert p.dtype == torch.float32
            arena[off:off + n].copy_(p.data.reshape(-1))
            p.data = arena[off:off + n].view(p.shape)
        self._arena = arena
        self._bind_probe = None
        self._garena = torch.zeros_like(arena)
        self._shadow = torch.zeros(2 * lay.total, dtype=torch.bfloat16, device=dev)    # bf16 copy | f16 copy (bits)
        self._ws_cache = {}
        self._rng_state = None
        self._act_state = None
        return self

    def net_struct(self) -> L.Net:
        self.bind()
        return L.Net(cfg=self._cfg, params=self._arena.data_ptr(), grads=self._garena.data_ptr(),
                     shadow=self._shadow.data_ptr())

    def refresh_shadow(self):
        net = self.net_struct()
        L.check(L.lib().dgvit_refresh_shadow(C.byref(net), _stream(self._arena.device)), "refresh_shadow")

    def _precision_code(self) -> int:
        return {"fp32": L.FP32, "bf16": L.BF16}[self.precision]

    def _workspace(self, B: int, save: bool) -> torch.Tensor:
        nbytes = C.c_size_t()
        L.check(L.lib().dgvit_workspace_bytes(C.byref(self._cfg), B, self._precision_code(), int(save),
                                              C.byref(nbytes)), "workspace_bytes")
        dev = self._arena.device
        if save:  # lives until its backward has run
            return torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        key = (B, self.precision)
        ws = self._ws_cache.get(key)
        if ws is None or ws.numel() < nbytes.value:
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            self._ws_cache[key] = ws
        return ws

    # ---- stochastic inputs
    def inject_noise(self, mask: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None):
        """Queue the stochastic inputs of the NEXT call (parity tests): ``mask`` is the {0,1}
        keep-mask [B, N, D] of the embedding dropout, ``eps`` the rsample draw [B, n_act]."""
        self._noise_fifo.append(dict(mask=mask, eps=eps))

    def _drop_struct(self, B: int, keep: List, training: Optional[bool] = None) -> Tuple[L.Drop, Optional[torch.Tensor]]:
        dev = self._arena.device
        inj = self._noise_fifo.pop(0) if self._noise_fifo else None
        eps = None
        if inj is not None and inj.get("eps") is not None:
            eps = inj["eps"].to(dev, torch.float32).contiguous()
        if self._rng_state is None:
            self._rng_state = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
        d = L.Drop(mode=L.DROP_NONE, p=float(self.trans.emb_dropout), keep_mask=None,
                   rng_state=self._rng_state.data_ptr(), stream_id=0)
        if inj is not None and inj.get("mask") is not None:
            m = inj["mask"].to(dev).to(torch.uint8).contiguous()
            assert m.numel() == B * self.layout_tokens() * self._cfg.dim, "mask shape"
            keep.append(m)
            d.mode, d.keep_mask = L.DROP_MASK, m.data_ptr()
        elif (self.training if training is None else training) and self.trans.emb_dropout > 0 and inj is None:
            d.mode = L.DROP_RNG
            self._rng_state[1] += 1          # host-ordered counter bump (plumbing, not arithmetic)
            # The backward kernels regenerate the keep mask from {seed, counter}: bind the counter to THIS call (a
            # snapshot kept alive with the saved workspace), so a later forward on the same module cannot change the mask
            # an earlier, not yet back-propagated pass used (learn_guidence with the CNN critic: two samples, one backward).
            call_state = self._rng_state.clone()
            keep.append(call_state)
            d.rng_state = call_state.data_ptr()
        return d, eps

    def layout_tokens(self) -> int:
        c = self._cfg
        return (c.img_h // c.patch_h) * (c.img_w // c.patch_w) + 1

    def _check_img(self, istate: torch.Tensor) -> torch.Tensor:
        c = self._cfg
        if istate.dim() != 3 or istate.shape[1] != c.img_h or istate.shape[2] != c.img_w:
            raise ValueError(f"istate must be [B,{c.img_h},{c.img_w}], got {tuple(istate.shape)}")
        return istate.to(self._arena.device, torch.float32).contiguous()

    def __deepcopy__(self, memo):
        # CUDA graphs / events / cached workspaces are per-instance runtime state, not model state
        import copy as _copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        skip = {"_act_state": None, "_ws_cache": {}, "_noise_fifo": [], "_rng_state": None, "_bc": None}
        for k, v in self.__dict__.items():
            new.__dict__[k] = skip[k] if k in skip else _copy.deepcopy(v, memo)
        trunk = new._modules.get("trans")
        if trunk is not None:
            trunk._set_owner(new)          # the copied trunk runs on the copy's arena
        return new

    # nn.Module plumbing: .to()/.cuda()/.float() replace p.data -> arenas are re-bound lazily
    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._arena = None
        return r

    def _grads_for_autograd(self, needs: List[bool]) -> List[Optional[torch.Tensor]]:
        ps = dict(self.named_parameters())
        out = []
        for (name, off), need in zip(self._named_offsets(), needs):
            if not need or self._unused(name):
                out.append(None)
            else:
                p = ps[name]
                out.append(self._garena[off:off + p.numel()].view(p.shape).clone())
        return out


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("dgvit_b200 runs on CUDA (sm_100a) only; move the module with .to('cuda'). "
                           "There is no CPU fallback.")


# --------------------------------------------------------------------------- trunk on its own
class _TrunkFn(torch.autograd.Function):
    """GoT.forward through ``dgvit_trunk_forward`` / ``dgvit_trunk_backward``."""

    @staticmethod
    def forward(ctx, be, got, need_grad, img, goal, *params):
        B, dev = img.shape[0], img.device
        keep: List = []
        drop, _ = be._drop_struct(B, keep, training=got.training)
        goal_c = goal.detach().contiguous()
        z = torch.empty(B, got.dim, device=dev)
        nbytes = C.c_size_t()
        L.check(L.lib().dgvit_trunk_workspace_bytes(C.byref(be._cfg), B, be._precision_code(), int(need_grad),
                                                    C.byref(nbytes)), "trunk_workspace_bytes")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        io = L.TrunkIO(img=img.data_ptr(), goal=goal_c.data_ptr(), drop=drop, sample_offset=0, z=z.data_ptr())
        net = be.net_struct()
        if be.precision == "bf16":
            be.refresh_shadow()
        L.check(L.lib().dgvit_trunk_forward(C.byref(net), C.byref(io), B, be._precision_code(), int(need_grad),
                                            ws.data_ptr(), ws.numel(), _stream(dev)), "trunk_forward")
        ctx.be, ctx.got, ctx.io, ctx.ws, ctx.B = be, got, io, ws, B
        ctx.keep = (img, goal_c, z, keep)
        ctx.needs = [p.requires_grad for p in params]
        ctx.goal_grad = goal.requires_grad
        return z

    @staticmethod
    def backward(ctx, d_z):
        be, got, B = ctx.be, ctx.got, ctx.B
        dev = ctx.keep[0].device
        d_z = d_z.contiguous().float()
        d_goal = torch.empty(B, got.dim, device=dev) if ctx.goal_grad else None
        net = be.net_struct()
        L.check(L.lib().dgvit_trunk_backward(C.byref(net), C.byref(ctx.io), d_z.data_ptr(), L.ptr(d_goal), B,
                                             be._precision_code(), ctx.ws.data_ptr(), ctx.ws.numel(), _stream(dev)),
                "trunk_backward")
        offs = dict(_trunk_offsets(be.layout(), be._cfg.depth, pre=""))
        grads = []
        for (name, p), need in zip(got.named_parameters(), ctx.needs):
            if not need or be._unused(name):
                grads.append(None)
            else:
                grads.append(be._garena[offs[name]:offs[name] + p.numel()].view(p.shape).clone())
        return (None, None, None, None, d_goal) + tuple(grads)


class _StandaloneTrunk(_ArenaModule):
    """Arena + C structs for a ``GoT`` used outside GoTPolicy / GoTQNetwork (an actor-kind layout whose head ranges
    stay unused)."""

    KIND = L.ACTOR

    def __init__(self, got: "GoT"):
        super().__init__()
        self.trans = got
        self._init_backend(2, 2, got.depth, got.heads, got.dim, image_size=got.image_size, patch_size=got.patch_size,
                           mlp_dim=got.mlp_dim, dim_head=got.dim_head)

    def _named_offsets(self):
        return _trunk_offsets(self.layout(), self._cfg.depth)


# --------------------------------------------------------------------------- actor
class _ActorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, need_grad, img, pstate, eps, *params):
        B = img.shape[0]
        dev = img.device
        na = mod._cfg.n_act
        keep: List = []
        drop, inj_eps = mod._drop_struct(B, keep)
        if eps is None:
            eps = inj_eps if inj_eps is not None else torch.randn(B, na, device=dev, dtype=torch.float32)
        eps = eps.contiguous()
        mean = torch.empty(B, na, device=dev)
        log_std = torch.empty(B, na, device=dev)
        action = torch.empty(B, na, device=dev)
        log_prob = torch.empty(B, 1, device=dev)
        mean_t = torch.empty(B, na, device=dev)
        scale = mod.action_scale.to(dev, torch.float32).expand(na).contiguous()
        bias = mod.action_bias.to(dev, torch.float32).expand(na).contiguous()
        ws = mod._workspace(B, need_grad)
        io = L.ActorIO(img=img.data_ptr(), pstate=pstate.data_ptr(), eps=eps.data_ptr(),
                       action_scale=scale.data_ptr(), action_bias=bias.data_ptr(), drop=drop, sample_offset=0,
                       mean=mean.data_ptr(), log_std=log_std.data_ptr(), action=action.data_ptr(),
                       log_prob=log_prob.data_ptr(), mean_t=mean_t.data_ptr(), eps_out=None)
        net = mod.net_struct()
        if mod.precision == "bf16":
            mod.refresh_shadow()
        L.check(L.lib().dgvit_actor_forward(C.byref(net), C.byref(io), B, mod._precision_code(), int(need_grad),
                                            ws.data_ptr(), ws.numel(), _stream(dev)), "actor_forward")
        ctx.mod, ctx.io, ctx.ws, ctx.B = mod, io, ws, B
        ctx.keep = (img, pstate, eps, scale, bias, keep)
        ctx.needs = [p.requires_grad for p in params]
        return mean, log_std, action, log_prob, mean_t

    @staticmethod
    def backward(ctx, d_mean, d_log_std, d_action, d_log_prob, d_mean_t):
        mod, B = ctx.mod, ctx.B
        dev = ctx.keep[0].device
        gs = [None if g is None else g.contiguous().float() for g in (d_mean, d_log_std, d_action, d_log_prob, d_mean_t)]
        g = L.ActorGrad(d_mean=L.ptr(gs[0]), d_log_std=L.ptr(gs[1]), d_action=L.ptr(gs[2]), d_log_prob=L.ptr(gs[3]),
                        d_mean_t=L.ptr(gs[4]), d_log_prob_const=0.0)
        net = mod.net_struct()
        L.check(L.lib().dgvit_actor_backward(C.byref(net), C.byref(ctx.io), C.byref(g), B, mod._precision_code(),
                                             ctx.ws.data_ptr(), ctx.ws.numel(), _stream(dev)), "actor_backward")
        return (None, None, None, None, None) + tuple(mod._grads_for_autograd(ctx.needs))


class GoTPolicy(_ArenaModule):
    """Drop-in for the reference ``GoTPolicy`` (vn/got_sac_network.py:172-256)."""

    KIND = L.ACTOR

    def __init__(self, nb_actions, nb_pstate, block, head, l_f_size, action_space=None, *,
                 image_size=(128, 160), mlp_dim=2048):
        super().__init__()
        self.trans = GoT(image_size=image_size, patch_size=(16, 20), num_classes=2, dim=l_f_size, depth=block,
                         heads=head, mlp_dim=mlp_dim, channels=4)
        self.fc_embed = nn.Linear(nb_pstate, l_f_size)
        self.fc1 = nn.Linear(l_f_size, 128)
        self.fc2 = nn.Linear(128, 128)
        self.mean_linear = nn.Linear(128, nb_actions)
        self.log_std_linear = nn.Linear(128, nb_actions)
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.apply(weights_init_)
        if action_space is None:
            self.action_scale = torch.tensor(1.0)
            self.action_bias = torch.tensor(0.0)
        else:
            self.action_scale = torch.FloatTensor((action_space.high - action_space.low) / 2.0)
            self.action_bias = torch.FloatTensor((action_space.high + action_space.low) / 2.0)
        self._init_backend(nb_actions, nb_pstate, block, head, l_f_size, image_size=image_size, mlp_dim=mlp_dim)
        self.trans._set_owner(self)

    def _named_offsets(self):
        lay = self.layout()
        return _trunk_offsets(lay, self._cfg.depth) + [
            ("fc_embed.weight", lay.embed_w), ("fc_embed.bias", lay.embed_b),
            ("fc1.weight", lay.fc1_w), ("fc1.bias", lay.fc1_b), ("fc2.weight", lay.fc2_w), ("fc2.bias", lay.fc2_b),
            ("mean_linear.weight", lay.mean_w), ("mean_linear.bias", lay.mean_b),
            ("log_std_linear.weight", lay.lstd_w), ("log_std_linear.bias", lay.lstd_b)]

    def _run(self, inp, eps=None):
        istate, pstate = inp
        self.bind()
        _require_cuda(self._arena)
        img = self._check_img(istate)
        ps = pstate.to(img.device, torch.float32).contiguous()
        ps_ = list(self.parameters())
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in ps_)
        return _ActorFn.apply(self, need_grad, img, ps, eps, *ps_)

    def forward(self, inp):
        mean, log_std, _, _, _ = self._run(inp)
        return mean, log_std

    def sample(self, inp):
        _, _, action, log_prob, mean_t = self._run(inp)
        return action, log_prob, mean_t

    def bc_step(self, istate, pstate, action, lr=1e-3, max_action=1.0, max_norm=10.0) -> torch.Tensor:
        """One behaviour-cloning step of the reference's imitation script (vn/attention_imitating.py:48-67) in ONE library
        call: ``policy.sample`` -> ``sqrt(mean((mean.clip(-max_action, max_action) - action) ** 2))`` -> backward ->
        ``clip_grad_norm_(parameters, max_norm)`` -> ``Adam(lr).step()``.  The Adam moments live with the module (created on
        the first call, ``bc_reset()`` drops them).  Returns the loss as a device tensor (no host sync); the total gradient
        norm before clipping is kept in ``self._bc["grad_norm"]``."""
        self.bind()
        _require_cuda(self._arena)
        dev = self._arena.device
        img = self._check_img(istate)
        ps = pstate.to(dev, torch.float32).contiguous()
        tgt = action.to(dev, torch.float32).contiguous()
        B = img.shape[0]
        st = getattr(self, "_bc", None)
        if st is None or st["arena"] != self._arena.data_ptr():
            n = self.layout().total
            f32 = dict(dtype=torch.float32, device=dev)
            st = self._bc = dict(arena=self._arena.data_ptr(), m=torch.zeros(n, **f32), v=torch.zeros(n, **f32),
                                 step=torch.zeros(1, dtype=torch.int64, device=dev), loss=torch.zeros(1, **f32),
                                 grad_norm=torch.zeros(1, **f32), ws=None,
                                 scale=self.action_scale.to(dev, torch.float32).expand(self._cfg.n_act).contiguous(),
                                 bias=self.action_bias.to(dev, torch.float32).expand(self._cfg.n_act).contiguous())
        nb = C.c_size_t()
        L.check(L.lib().dgvit_bc_workspace_bytes(C.byref(self._cfg), B, self._precision_code(), C.byref(nb)), "bc_workspace_bytes")
        if st["ws"] is None or st["ws"].numel() < nb.value:
            st["ws"] = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        keep: List = []
        drop, eps = self._drop_struct(B, keep, training=True)
        if self.precision == "bf16":
            self._sync_shadow(st)
        net = self.net_struct()
        opt = L.Adam(m=st["m"].data_ptr(), v=st["v"].data_ptr(), step=st["step"].data_ptr(), lr=float(lr), beta1=0.9, beta2=0.999,
                     eps=1e-8)
        io = L.BcIO(img=img.data_ptr(), pstate=ps.data_ptr(), target=tgt.data_ptr(), eps=L.ptr(eps),
                    action_scale=st["scale"].data_ptr(), action_bias=st["bias"].data_ptr(), drop=drop, sample_offset=0,
                    advance_rng=0, max_action=float(max_action), max_norm=float(max_norm), loss=st["loss"].data_ptr(),
                    grad_norm=st["grad_norm"].data_ptr())
        L.check(L.lib().dgvit_bc_step(C.byref(net), C.byref(opt), C.byref(io), B, self._precision_code(), st["ws"].data_ptr(),
                                      st["ws"].numel(), _stream(dev)), "bc_step")
        st["keep"] = (keep, img, ps, tgt, eps)       # inputs stay alive until the next step has been queued behind this one
        if self.precision == "bf16":
            st["shadow_version"] = self._arena._version      # (the step's Adam pass refreshed the 16-bit copies itself)
        return st["loss"]

    def bc_reset(self):
        self._bc = None

    def choose_action(self, istate, pstate, evaluate=False):
        """numpy (H,W,1) frame + (2,) goal -> numpy action (vn/got_sac_network.py:205-220).

        Batch-1 control-loop path: pinned host staging -> one CUDA graph (H2D, the whole actor
        forward, D2H) -> one stream synchronise.  ``evaluate=True`` returns tanh(mean)."""
        istate = np.asarray(istate)
        if istate.ndim >= 4:
            raise ValueError("4-D (frame-stacked) observations are a legacy path the DGViT trunk rejects")
        self.bind()
        _require_cuda(self._arena)
        c = self._cfg
        st = self._act_state
        key = (self._arena.data_ptr(), self.training, self.precision)
        if st is None or st["key"] != key:
            st = self._build_act_graph(key)
        st["img_np"][...] = istate.reshape(1, c.img_h, c.img_w)
        st["ps_np"][...] = np.asarray(pstate, dtype=np.float32).reshape(1, c.n_pstate)
        self._sync_shadow(st)
        st["graph"].replay()
        st["done"].record()
        st["done"].synchronize()
        return (st["mean_t_np"] if evaluate else st["action_np"])[0].copy()

    def _sync_shadow(self, st):
        """The 16-bit operand copies follow the fp32 arena: the library's own optimizer kernels keep them current; a torch-side
        in-place write (torch.optim step, load_state_dict, ...) bumps the arena's version counter and is caught here — outside
        the replayed graph, so the control loop does not pay a 1.4 M-element pass per action."""
        if self.precision != "bf16":
            return
        v = self._arena._version
        if st.get("shadow_version") != v:
            self.refresh_shadow()
            st["shadow_version"] = v

    def _build_act_graph(self, key):
        c = self._cfg
        dev = self._arena.device
        na = c.n_act
        f32 = dict(dtype=torch.float32)
        st = dict(key=key,
                  img_host=torch.zeros(1, c.img_h, c.img_w, **f32).pin_memory(),
                  ps_host=torch.zeros(1, c.n_pstate, **f32).pin_memory(),
                  action_host=torch.zeros(1, na, **f32).pin_memory(), mean_t_host=torch.zeros(1, na, **f32).pin_memory(),
                  img=torch.zeros(1, c.img_h, c.img_w, device=dev, **f32), ps=torch.zeros(1, c.n_pstate, device=dev, **f32),
                  mean=torch.zeros(1, na, device=dev, **f32), log_std=torch.zeros(1, na, device=dev, **f32),
                  action=torch.zeros(1, na, device=dev, **f32), log_prob=torch.zeros(1, 1, device=dev, **f32),
                  mean_t=torch.zeros(1, na, device=dev, **f32),
                  scale=self.action_scale.to(dev, torch.float32).expand(na).contiguous(),
                  bias=self.action_bias.to(dev, torch.float32).expand(na).contiguous(),
                  ws=self._workspace(1, False), done=torch.cuda.Event())
        if self._rng_state is None:
            self._rng_state = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
        drop = L.Drop(mode=L.DROP_RNG if (self.training and self.trans.emb_dropout > 0) else L.DROP_NONE,
                      p=float(self.trans.emb_dropout), keep_mask=None, rng_state=self._rng_state.data_ptr(), stream_id=7)
        # Zero-copy staging: pinned host memory is device-addressable (UVA), so the patchify / goal-token kernels read the
        # 80 KB frame and the goal straight from the host buffers and the sampling kernel writes the action back into host
        # memory -- four memcpy nodes (H2D x2, D2H x2, ~6 us each in a graph) fewer on a 130 us critical path.
        zc = os.environ.get("DGVIT_ACT_ZERO_COPY", "1") == "1"
        src = (lambda k: st[k + "_host"].data_ptr()) if zc else (lambda k: st[k].data_ptr())
        io = L.ActorIO(img=src("img"), pstate=src("ps"), eps=None, action_scale=st["scale"].data_ptr(),
                       action_bias=st["bias"].data_ptr(), drop=drop, sample_offset=0, mean=st["mean"].data_ptr(),
                       log_std=st["log_std"].data_ptr(), action=src("action"), log_prob=st["log_prob"].data_ptr(),
                       mean_t=src("mean_t"), eps_out=None, advance_rng=1)     # fresh rsample / dropout stream per call
        net = self.net_struct()
        st["shadow_version"] = -1

        def run():
            if not zc:
                st["img"].copy_(st["img_host"], non_blocking=True)
                st["ps"].copy_(st["ps_host"], non_blocking=True)
            L.check(L.lib().dgvit_actor_forward(C.byref(net), C.byref(io), 1, self._precision_code(), 0,
                                                st["ws"].data_ptr(), st["ws"].numel(), _stream(dev)), "actor_forward")
            if not zc:
                st["action_host"].copy_(st["action"], non_blocking=True)
                st["mean_t_host"].copy_(st["mean_t"], non_blocking=True)

        self._sync_shadow(st)
        run()                                   # eager warm-up (lazy kernel attributes, tensor-map entry point)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        st["graph"] = g
        st["io"] = io
        for k in ("img", "ps", "action", "mean_t"):      # numpy views of the pinned staging buffers
            st[k + "_np"] = st[k + "_host"].numpy()
        self._act_state = st
        return st

    def to(self, device):
        self.action_scale = self.action_scale.to(device)
        self.action_bias = self.action_bias.to(device)
        return super().to(device)


# --------------------------------------------------------------------------- critic
class _CriticFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, need_grad, img, pstate, action, *params):
        B = img.shape[0]
        dev = img.device
        na = mod._cfg.n_act
        keep: List = []
        drop, _ = mod._drop_struct(B, keep)
        q1 = torch.empty(B, na, device=dev)
        q2 = torch.empty(B, na, device=dev)
        act = action.detach().contiguous()
        ws = mod._workspace(B, need_grad)
        io = L.CriticIO(img=img.data_ptr(), pstate=pstate.data_ptr(), action=act.data_ptr(), drop=drop,
                        q1=q1.data_ptr(), q2=q2.data_ptr())
        net = mod.net_struct()
        if mod.precision == "bf16":
            mod.refresh_shadow()
        L.check(L.lib().dgvit_critic_forward(C.byref(net), C.byref(io), B, mod._precision_code(), int(need_grad),
                                             ws.data_ptr(), ws.numel(), _stream(dev)), "critic_forward")
        ctx.mod, ctx.io, ctx.ws, ctx.B = mod, io, ws, B
        ctx.keep = (img, pstate, act, keep)
        ctx.needs = [p.requires_grad for p in params]
        ctx.act_grad = action.requires_grad
        return q1, q2

    @staticmethod
    def backward(ctx, d_q1, d_q2):
        mod, B = ctx.mod, ctx.B
        dev = ctx.keep[0].device
        na = mod._cfg.n_act
        d_q1 = torch.zeros(B, na, device=dev) if d_q1 is None else d_q1.contiguous().float()
        d_q2 = torch.zeros(B, na, device=dev) if d_q2 is None else d_q2.contiguous().float()
        d_act = torch.empty(B, na, device=dev) if ctx.act_grad else None
        pg = any(ctx.needs)
        net = mod.net_struct()
        L.check(L.lib().dgvit_critic_backward(C.byref(net), C.byref(ctx.io), d_q1.data_ptr(), d_q2.data_ptr(),
                                              L.ptr(d_act), int(pg), B, mod._precision_code(), ctx.ws.data_ptr(),
                                              ctx.ws.numel(), _stream(dev)), "critic_backward")
        grads = mod._grads_for_autograd(ctx.needs) if pg else [None] * len(ctx.needs)
        return (None, None, None, None, d_act) + tuple(grads)


class GoTQNetwork(_ArenaModule):
    """Drop-in for the reference ``GoTQNetwork`` (vn/got_sac_network.py:75-123)."""

    KIND = L.CRITIC

    def __init__(self, nb_actions, nb_pstate, block, head, l_f_size, *, image_size=(128, 160), mlp_dim=2048):
        super().__init__()
        self.trans = GoT(image_size=image_size, patch_size=(16, 20), num_classes=2, dim=l_f_size, depth=block,
                         heads=head, mlp_dim=mlp_dim, channels=1)
        # constructed (and kept in state_dict) but never used by forward, as in the reference (:90-94)
        self.conv1 = nn.Conv2d(4, 16, 5, stride=2)
        self.conv2 = nn.Conv2d(16, 64, 5, stride=2)
        self.conv3 = nn.Conv2d(64, 256, 5, stride=2)
        self.avg = nn.AdaptiveAvgPool2d(output_size=(1, 1))
        self.fc1 = nn.Linear(l_f_size + nb_actions, 128)
        self.fc2 = nn.Linear(128, 32)
        self.fc3 = nn.Linear(32, nb_actions)
        self.fc_embed = nn.Linear(nb_pstate, l_f_size)
        self.fc11 = nn.Linear(l_f_size + nb_actions, 128)
        self.fc21 = nn.Linear(128, 32)
        self.fc31 = nn.Linear(32, nb_actions)
        self.apply(weights_init_)
        self._init_backend(nb_actions, nb_pstate, block, head, l_f_size, image_size=image_size, mlp_dim=mlp_dim)
        self.trans._set_owner(self)

    def _named_offsets(self):
        lay = self.layout()
        return _trunk_offsets(lay, self._cfg.depth) + [
            ("conv1.weight", lay.conv1_w), ("conv1.bias", lay.conv1_b), ("conv2.weight", lay.conv2_w),
            ("conv2.bias", lay.conv2_b), ("conv3.weight", lay.conv3_w), ("conv3.bias", lay.conv3_b),
            ("fc1.weight", lay.fc1_w), ("fc1.bias", lay.fc1_b), ("fc2.weight", lay.fc2_w), ("fc2.bias", lay.fc2_b),
            ("fc3.weight", lay.fc3_w), ("fc3.bias", lay.fc3_b),
            ("fc_embed.weight", lay.embed_w), ("fc_embed.bias", lay.embed_b),
            ("fc11.weight", lay.fc11_w), ("fc11.bias", lay.fc11_b), ("fc21.weight", lay.fc21_w),
            ("fc21.bias", lay.fc21_b), ("fc31.weight", lay.fc31_w), ("fc31.bias", lay.fc31_b)]

    def forward(self, inp):
        istate, pstate, a = inp
        self.bind()
        _require_cuda(self._arena)
        img = self._check_img(istate)
        ps = pstate.to(img.device, torch.float32).contiguous()
        act = a.to(img.device, torch.float32)
        ps_ = list(self.parameters())
        need_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in ps_) or act.requires_grad)
        return _CriticFn.apply(self, need_grad, img, ps, act, *ps_)


# --------------------------------------------------------------------------- CNN critic
class _QnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, need_grad, img, pstate, action, *params):
        B, dev, na = img.shape[0], img.device, mod.nb_actions
        q1 = torch.empty(B, na, device=dev)
        q2 = torch.empty(B, na, device=dev)
        act = action.detach().contiguous()
        ws = mod._workspace(B, need_grad)
        L.check(L.lib().dgvit_qnet_forward(mod._arena.data_ptr(), img.data_ptr(), pstate.data_ptr(), act.data_ptr(),
                                           q1.data_ptr(), q2.data_ptr(), mod.image_size[0], mod.image_size[1], na,
                                           mod.nb_pstate, B, mod._precision_code(), ws.data_ptr(), ws.numel(), _stream(dev)),
                "qnet_forward")
        ctx.mod, ctx.ws, ctx.B = mod, ws, B
        ctx.keep = (img, pstate, act)
        ctx.needs = [p.requires_grad for p in params]
        ctx.act_grad = action.requires_grad
        return q1, q2

    @staticmethod
    def backward(ctx, d_q1, d_q2):
        mod, B = ctx.mod, ctx.B
        img, pstate, _ = ctx.keep
        dev, na = img.device, mod.nb_actions
        d_q1 = torch.zeros(B, na, device=dev) if d_q1 is None else d_q1.contiguous().float()
        d_q2 = torch.zeros(B, na, device=dev) if d_q2 is None else d_q2.contiguous().float()
        d_act = torch.empty(B, na, device=dev) if ctx.act_grad else None
        pg = any(ctx.needs)
        L.check(L.lib().dgvit_qnet_backward(mod._arena.data_ptr(), mod._garena.data_ptr(), img.data_ptr(), pstate.data_ptr(),
                                            d_q1.data_ptr(), d_q2.data_ptr(), L.ptr(d_act), int(pg), mod.image_size[0],
                                            mod.image_size[1], na, mod.nb_pstate, B, mod._precision_code(),
                                            ctx.ws.data_ptr(), ctx.ws.numel(), _stream(dev)), "qnet_backward")
        grads = [None] * len(ctx.needs)
        if pg:
            ps = dict(mod.named_parameters())
            grads = [mod._garena[off:off + ps[n].numel()].view(ps[n].shape).clone() if need else None
                     for (n, off), need in zip(mod._named_offsets(), ctx.needs)]
        return (None, None, None, None, d_act) + tuple(grads)


class QNetwork(nn.Module):
    """Drop-in for the reference CNN twin-Q critic ``QNetwork`` (vn/got_sac_network.py:125-170), the shipped
    default ``critic_type`` (vn/config.yaml:61): same constructor, ``forward([istate, pstate, a]) -> (q1, q2)``,
    attribute names and ``state_dict`` keys.  The parameters are views into one flat arena
    (``dgvit_qnet_param_layout``); forward / backward run ``dgvit_qnet_forward`` / ``dgvit_qnet_backward``."""

    def __init__(self, nb_actions, nb_pstate, *, image_size=(128, 160)):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 16, 5, stride=2)
        self.conv2 = nn.Conv2d(16, 64, 5, stride=2)
        self.conv3 = nn.Conv2d(64, 256, 5, stride=2)
        self.avg = nn.AdaptiveAvgPool2d(output_size=(1, 1))
        self.fc1 = nn.Linear(256 + 32 + nb_actions, 128)
        self.fc2 = nn.Linear(128, 32)
        self.fc3 = nn.Linear(32, nb_actions)
        self.fc_embed = nn.Linear(nb_pstate, 32)
        self.fc11 = nn.Linear(256 + 32 + nb_actions, 128)
        self.fc21 = nn.Linear(128, 32)
        self.fc31 = nn.Linear(32, nb_actions)
        self.apply(weights_init_)
        self.nb_actions, self.nb_pstate, self.image_size = nb_actions, nb_pstate, tuple(image_size)
        self.precision = "fp32"
        self._arena = self._garena = None
        self._ws_cache: Dict = {}
        self._layout = None

    def layout(self) -> L.QnetLayout:
        if self._layout is None:
            out = L.QnetLayout()
            L.check(L.lib().dgvit_qnet_param_layout(self.nb_actions, self.nb_pstate, C.byref(out)), "qnet_param_layout")
            self._layout = out
        return self._layout

    def _named_offsets(self) -> List[Tuple[str, int]]:
        lay = self.layout()
        out = []
        for i in range(3):
            out += [(f"conv{i + 1}.weight", lay.conv_w[i]), (f"conv{i + 1}.bias", lay.conv_b[i])]
        for n, f in (("fc1", "fc1"), ("fc2", "fc2"), ("fc3", "fc3"), ("fc_embed", "embed"), ("fc11", "fc11"),
                     ("fc21", "fc21"), ("fc31", "fc31")):
            out += [(n + ".weight", getattr(lay, f + "_w")), (n + ".bias", getattr(lay, f + "_b"))]
        return out

    def _bound(self) -> bool:
        if self._arena is None:
            return False
        ps = dict(self.named_parameters())
        base = self._arena.data_ptr()
        offs = self._named_offsets()
        return all(ps[n].device == self._arena.device and ps[n].data_ptr() == base + 4 * off for n, off in (offs[0], offs[-1]))

    def bind(self, force: bool = False):
        if not force and self._bound():
            return self
        lay = self.layout()
        ps = dict(self.named_parameters())
        offs = self._named_offsets()
        assert [n for n, _ in offs] == list(ps.keys()), "parameter registration order differs from the C layout"
        dev = next(iter(ps.values())).device
        arena = torch.zeros(lay.total, dtype=torch.float32, device=dev)
        for name, off in offs:
            p = ps[name]
            arena[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = arena[off:off + p.numel()].view(p.shape)
        self._arena, self._garena, self._ws_cache = arena, torch.zeros_like(arena), {}
        return self

    def net_struct(self) -> L.Net:
        """The critic / critic_target slot of ``dgvit_sac`` (cfg.kind == DGVIT_QNET: only the input shapes are read)."""
        self.bind()
        cfg = L.Cfg(kind=L.QNET, img_h=self.image_size[0], img_w=self.image_size[1], n_act=self.nb_actions,
                    n_pstate=self.nb_pstate)
        return L.Net(cfg=cfg, params=self._arena.data_ptr(), grads=self._garena.data_ptr(), shadow=None)

    def _precision_code(self) -> int:
        return {"fp32": L.FP32, "bf16": L.BF16}[self.precision]

    def _workspace(self, B: int, save: bool) -> torch.Tensor:
        nbytes = C.c_size_t()
        L.check(L.lib().dgvit_qnet_workspace_bytes(self.image_size[0], self.image_size[1], self.nb_actions, self.nb_pstate, B,
                                                   self._precision_code(), C.byref(nbytes)), "qnet_workspace_bytes")
        dev = self._arena.device
        if save:
            return torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        key = (B, self.precision)
        ws = self._ws_cache.get(key)
        if ws is None or ws.numel() < nbytes.value:
            ws = self._ws_cache[key] = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        return ws

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._arena = None
        return r

    def __deepcopy__(self, memo):
        import copy as _copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_ws_cache" else _copy.deepcopy(v, memo)
        return new

    def forward(self, inp):
        istate, pstate, a = inp
        self.bind()
        _require_cuda(self._arena)
        h, w = self.image_size
        if istate.dim() != 3 or istate.shape[1] != h or istate.shape[2] != w:
            raise ValueError(f"istate must be [B,{h},{w}], got {tuple(istate.shape)}")
        img = istate.to(self._arena.device, torch.float32).contiguous()
        ps = pstate.to(img.device, torch.float32).contiguous()
        act = a.to(img.device, torch.float32)
        ps_ = list(self.parameters())
        need_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in ps_) or act.requires_grad)
        return _QnetFn.apply(self, need_grad, img, ps, act, *ps_)
