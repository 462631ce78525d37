"""CPU oracle for the DGViT actor-critic hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32) *restatement* of the reference
algorithm.  It is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path
(``dgvit_b200``) never routes through it and has no CPU fallback.

Pinning: the reference ships no golden vectors / KATs for this path
(SURVEY.md §4, §8c), so the oracle is pinned against the *imported reference
modules themselves* in this container by ``oracle/make_golden.py``
(module forward/backward parity, and the unmodified ``SAC.learn`` run with a
stub ``cpprb``).  The outputs of that run are committed as fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the oracle
against them on any box (the reference tree does not travel).

Every function cites the reference lines it follows; paths are relative to
``/root/reference/src/vis_nav/vis_nav/`` (abbreviated ``vn/``).

Stochastic inputs are explicit: ``drop_mask`` is the {0,1} keep-mask of the
embedding dropout (vn/GoalFormer.py:163, p=0.1, live because nothing in the
reference ever calls ``.eval()``) and ``eps`` the N(0,1) draw of
``Normal.rsample`` (vn/got_sac_network.py:242).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

LOG_SIG_MAX = 2.0     # vn/got_sac_network.py:17
LOG_SIG_MIN = -20.0   # vn/got_sac_network.py:18
EPSILON = 1e-6        # vn/got_sac_network.py:19
EMB_DROPOUT = 0.1     # vn/GoalFormer.py:124 (emb_dropout default)
LN_EPS = 1e-5         # nn.LayerNorm default, vn/GoalFormer.py:34
RMS_EPS = 1e-12       # F.normalize default eps, vn/GoalFormer.py:122


@dataclass(frozen=True)
class Cfg:
    """Shapes of one DGViT network (vn/got_sac_network.py:76-88,173-185)."""
    dim: int = 64            # l_f_size
    depth: int = 4           # block
    heads: int = 4           # head
    dim_head: int = 64       # vn/GoalFormer.py:124 default
    mlp_dim: int = 2048      # vn/got_sac_network.py:86,183
    img_h: int = 128
    img_w: int = 160
    patch_h: int = 16        # vn/GoalFormer.py:138 (hard-coded p1)
    patch_w: int = 20        # vn/GoalFormer.py:138 (hard-coded p2)
    nb_actions: int = 2
    nb_pstate: int = 2

    @property
    def n_patches(self) -> int:
        return (self.img_h // self.patch_h) * (self.img_w // self.patch_w)

    @property
    def n_tokens(self) -> int:
        return self.n_patches + 1

    @property
    def patch_dim(self) -> int:
        return self.patch_h * self.patch_w

    @property
    def inner(self) -> int:
        return self.heads * self.dim_head


# --------------------------------------------------------------------------
# Trunk (vn/GoalFormer.py)
# --------------------------------------------------------------------------

def patchify(img: torch.Tensor, cfg: Cfg) -> torch.Tensor:
    """``Rearrange('b (h p1) (w p2) -> b (h w) (p1 p2)')`` — vn/GoalFormer.py:138."""
    b = img.shape[0]
    gh, gw = cfg.img_h // cfg.patch_h, cfg.img_w // cfg.patch_w
    x = img.reshape(b, gh, cfg.patch_h, gw, cfg.patch_w)
    return x.permute(0, 1, 3, 2, 4).reshape(b, gh * gw, cfg.patch_h * cfg.patch_w)


def layer_norm(x, w, b):
    """``nn.LayerNorm(dim)`` — vn/GoalFormer.py:31-37."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def attention(x, p, pre, cfg: Cfg):
    """``Attention.forward`` — vn/GoalFormer.py:71-82."""
    b, n, _ = x.shape
    h, dh = cfg.heads, cfg.dim_head
    qkv = x @ p[pre + "fn.to_qkv.weight"].t()                      # :72 (bias=False, :64)
    q, k, v = qkv.chunk(3, dim=-1)                                 # :72
    q, k, v = (t.reshape(b, n, h, dh).permute(0, 2, 1, 3) for t in (q, k, v))   # :73
    dots = torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5)     # :75, scale :59
    attn = torch.softmax(dots, dim=-1)                             # :77 (dropout p=0, :78)
    out = torch.matmul(attn, v)                                    # :80
    out = out.permute(0, 2, 1, 3).reshape(b, n, h * dh)            # :81
    return out @ p[pre + "fn.to_out.0.weight"].t() + p[pre + "fn.to_out.0.bias"]  # :82


def feed_forward(x, p, pre):
    """``FeedForward.forward`` — vn/GoalFormer.py:39-50 (exact-erf GELU, dropouts p=0)."""
    hdn = x @ p[pre + "fn.net.0.weight"].t() + p[pre + "fn.net.0.bias"]
    hdn = 0.5 * hdn * (1.0 + torch.erf(hdn * (1.0 / math.sqrt(2.0))))
    return hdn @ p[pre + "fn.net.3.weight"].t() + p[pre + "fn.net.3.bias"]


def trunk_forward(p: Dict[str, torch.Tensor], img, goal_tok, cfg: Cfg,
                  drop_mask: Optional[torch.Tensor] = None, pre: str = "trans."):
    """``GoT.forward(img, goal)`` — vn/GoalFormer.py:156-171.

    ``drop_mask``: keep-mask [B, N, D] in {0,1} or None (= eval mode).
    """
    x = patchify(img, cfg) @ p[pre + "to_patch_embedding.1.weight"].t() \
        + p[pre + "to_patch_embedding.1.bias"]                    # :157
    x = torch.cat((goal_tok.unsqueeze(1), x), dim=1)               # :160-161
    x = x + p[pre + "pos_embedding"][:, : x.shape[1]]              # :162
    if drop_mask is not None:                                      # :163
        # aten dropout: noise = bernoulli(1-p) / (1-p); out = x * noise
        x = x * (drop_mask / (1.0 - EMB_DROPOUT))
    for l in range(cfg.depth):                                     # :101-105
        a = f"{pre}transformer.layers.{l}.0."
        f = f"{pre}transformer.layers.{l}.1."
        x = attention(layer_norm(x, p[a + "norm.weight"], p[a + "norm.bias"]), p, a, cfg) + x
        x = feed_forward(layer_norm(x, p[f + "norm.weight"], p[f + "norm.bias"]), p, f) + x
    x = x[:, 0]                                                    # :167 (pool='cls')
    # RMSNorm, :120-122 : F.normalize(x, dim=-1) * sqrt(dim) * g
    nrm = x.norm(dim=-1, keepdim=True).clamp_min(RMS_EPS)
    return x / nrm * (cfg.dim ** 0.5) * p[pre + "layer_norm.g"]


# --------------------------------------------------------------------------
# Actor / critic (vn/got_sac_network.py)
# --------------------------------------------------------------------------

def actor_forward(p, img, pstate, cfg: Cfg, drop_mask=None):
    """``GoTPolicy.forward`` — vn/got_sac_network.py:221-236."""
    tok = pstate @ p["fc_embed.weight"].t() + p["fc_embed.bias"]   # :226, no activation
    z = trunk_forward(p, img, tok, cfg, drop_mask)                 # :228
    x = F.relu(z @ p["fc1.weight"].t() + p["fc1.bias"])            # :230
    x = F.relu(x @ p["fc2.weight"].t() + p["fc2.bias"])            # :231
    mean = x @ p["mean_linear.weight"].t() + p["mean_linear.bias"]          # :233
    log_std = x @ p["log_std_linear.weight"].t() + p["log_std_linear.bias"]  # :234
    return mean, torch.clamp(log_std, min=LOG_SIG_MIN, max=LOG_SIG_MAX)     # :235


def actor_sample(p, img, pstate, eps, cfg: Cfg, drop_mask=None,
                 action_scale=1.0, action_bias=0.0):
    """``GoTPolicy.sample`` — vn/got_sac_network.py:238-251."""
    mean, log_std = actor_forward(p, img, pstate, cfg, drop_mask)
    std = log_std.exp()                                            # :240
    x_t = mean + std * eps                                         # :242 rsample
    y_t = torch.tanh(x_t)                                          # :243
    action = y_t * action_scale + action_bias                      # :245
    # Normal.log_prob: -((x-mu)^2)/(2 var) - log(std) - log(sqrt(2 pi))
    log_prob = -((x_t - mean) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))  # :246
    log_prob = log_prob - torch.log(action_scale * (1 - y_t.pow(2)) + EPSILON)   # :248
    log_prob = log_prob.sum(1, keepdim=True)                       # :249
    mean_t = torch.tanh(mean) * action_scale + action_bias         # :250
    return action, log_prob, mean_t


def critic_forward(p, img, pstate, a, cfg: Cfg, drop_mask=None):
    """``GoTQNetwork.forward`` — vn/got_sac_network.py:107-123."""
    tok = F.relu(pstate @ p["fc_embed.weight"].t() + p["fc_embed.bias"])   # :111 (ReLU here)
    z = trunk_forward(p, img, tok, cfg, drop_mask)                 # :112
    x = torch.cat([z, a], dim=1)                                   # :114
    q1 = F.relu(x @ p["fc1.weight"].t() + p["fc1.bias"])
    q1 = F.relu(q1 @ p["fc2.weight"].t() + p["fc2.bias"])
    q1 = q1 @ p["fc3.weight"].t() + p["fc3.bias"]                  # :115-117
    q2 = F.relu(x @ p["fc11.weight"].t() + p["fc11.bias"])
    q2 = F.relu(q2 @ p["fc21.weight"].t() + p["fc21.bias"])
    q2 = q2 @ p["fc31.weight"].t() + p["fc31.bias"]                # :119-121
    return q1, q2


def qnet_forward(p, img, pstate, a):
    """``QNetwork.forward`` (CNN twin-Q critic, the shipped default ``critic_type``) —
    vn/got_sac_network.py:149-170; layers :129-144."""
    x1 = img.unsqueeze(1)                                                       # :153
    x1 = F.relu(F.conv2d(x1, p["conv1.weight"], p["conv1.bias"], stride=2))     # :154
    x1 = F.relu(F.conv2d(x1, p["conv2.weight"], p["conv2.bias"], stride=2))     # :155
    x1 = F.relu(F.conv2d(x1, p["conv3.weight"], p["conv3.bias"], stride=2))     # :156
    x1 = x1.mean(dim=(2, 3))                                                    # :157-158 AdaptiveAvgPool2d((1,1)) + view
    x2 = F.relu(pstate @ p["fc_embed.weight"].t() + p["fc_embed.bias"])         # :160-161
    x = torch.cat([x1, x2, a], dim=1)                                           # :163
    q1 = F.relu(x @ p["fc1.weight"].t() + p["fc1.bias"])
    q1 = F.relu(q1 @ p["fc2.weight"].t() + p["fc2.bias"])
    q1 = q1 @ p["fc3.weight"].t() + p["fc3.bias"]                               # :165-167
    q2 = F.relu(x @ p["fc11.weight"].t() + p["fc11.bias"])
    q2 = F.relu(q2 @ p["fc21.weight"].t() + p["fc21.bias"])
    q2 = q2 @ p["fc31.weight"].t() + p["fc31.bias"]                             # :169-171
    return q1, q2


# --------------------------------------------------------------------------
# Optimizer / target update
# --------------------------------------------------------------------------

class Adam:
    """torch.optim.Adam defaults (betas .9/.999, eps 1e-8, no wd, no amsgrad),
    single-tensor formula; params whose grad is None are skipped, exactly as
    ``Adam.step`` does (vn/DRL.py:113,139,150 ctor; :403,414,422 step)."""

    def __init__(self, names, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.names, self.lr, self.b1, self.b2, self.eps = list(names), lr, b1, b2, eps
        self.m: Dict[str, torch.Tensor] = {}
        self.v: Dict[str, torch.Tensor] = {}
        self.t: Dict[str, int] = {}

    def step(self, params: Dict[str, torch.Tensor], grads: Dict[str, Optional[torch.Tensor]]):
        for k in self.names:
            g = grads.get(k)
            if g is None:
                continue
            if k not in self.m:
                self.m[k] = torch.zeros_like(params[k])
                self.v[k] = torch.zeros_like(params[k])
                self.t[k] = 0
            self.t[k] += 1
            t = self.t[k]
            self.m[k].lerp_(g, 1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1 = 1 - self.b1 ** t
            bc2 = 1 - self.b2 ** t
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            params[k].addcdiv_(self.m[k], denom, value=-(self.lr / bc1))


def soft_update(target: Dict[str, torch.Tensor], source: Dict[str, torch.Tensor], names, tau):
    """vn/utils.py:31-33 over ``parameters()`` in registration order (all params)."""
    for k in names:
        target[k].copy_(target[k] * (1.0 - tau) + source[k] * tau)


def hard_update(target, source, names):
    """vn/utils.py:35-37."""
    for k in names:
        target[k].copy_(source[k])


# --------------------------------------------------------------------------
# SAC.learn (vn/DRL.py:373-437) on tensors
# --------------------------------------------------------------------------

class SACOracle:
    """State + one ``learn`` step of the reference agent (``GaussianTransformer``
    actor, ``Transformer`` critic, AUTO_TUNE) — vn/DRL.py:35-169,373-437.

    ``actor`` / ``critic`` are dicts of fp32 tensors keyed like the reference
    ``state_dict`` *restricted to parameters, in registration order*.
    """

    def __init__(self, actor, critic, cfg: Cfg, lr_a=1e-3, lr_c=1e-3, lr_alpha=1e-4,
                 gamma=0.999, tau=5e-4, alpha=1.0, policy_freq=1,
                 automatic_entropy_tuning=True):
        self.cfg = cfg
        self.actor = {k: v.detach().clone() for k, v in actor.items()}
        self.critic = {k: v.detach().clone() for k, v in critic.items()}
        self.critic_target = {k: v.detach().clone() for k, v in critic.items()}   # hard_update, :123
        self.gamma, self.tau, self.alpha = gamma, tau, alpha
        self.policy_freq = policy_freq
        self.auto = automatic_entropy_tuning
        self.target_entropy = -float(cfg.nb_actions)                # :137
        self.log_alpha = torch.zeros(1)                             # :138
        self.actor_opt = Adam(self.actor.keys(), lr_a)              # :150
        self.critic_opt = Adam(self.critic.keys(), lr_c)            # :113
        self.alpha_opt = Adam(["log_alpha"], lr_alpha)              # :139
        self.itera = 0

    def learn(self, batch, noise):
        """``batch``: obs,pobs,act,rew,next_obs,next_pobs (done is read but unused, :394).
        ``noise``: eps_next, eps_pi [B,2]; keep-masks mask_a_next, mask_ct, mask_c,
        mask_a, mask_c_pi [B,N,D] (None = no dropout), in the order the reference
        consumes its generator."""
        cfg = self.cfg
        s, ps, a = batch["obs"], batch["pobs"], batch["act"]
        r, s2, ps2 = batch["rew"], batch["next_obs"], batch["next_pobs"]
        alpha = self.alpha

        with torch.no_grad():                                       # :389-393
            a2, logp2, _ = actor_sample(self.actor, s2, ps2, noise["eps_next"], cfg, noise.get("mask_a_next"))
            q1t, q2t = critic_forward(self.critic_target, s2, ps2, a2, cfg, noise.get("mask_ct"))
            min_q = torch.min(q1t, q2t) - alpha * logp2
            nq = r + self.gamma * min_q                             # :393, no (1-done)

        cp = {k: v.clone().requires_grad_(True) for k, v in self.critic.items()}
        q1, q2 = critic_forward(cp, s, ps, a, cfg, noise.get("mask_c"))      # :395
        qf1_loss = F.mse_loss(q1, nq)                               # :396
        qf2_loss = F.mse_loss(q2, nq)
        (qf1_loss + qf2_loss).backward()                            # :398-401
        cgrads = {k: v.grad for k, v in cp.items()}
        self.critic_opt.step(self.critic, cgrads)                   # :402
        self.last_critic_grads = cgrads

        ap = {k: v.clone().requires_grad_(True) for k, v in self.actor.items()}
        pi, log_pi, _ = actor_sample(ap, s, ps, noise["eps_pi"], cfg, noise.get("mask_a"))   # :404
        q1p, q2p = critic_forward(self.critic, s, ps, pi, cfg, noise.get("mask_c_pi"))       # :406
        min_qp = torch.min(q1p, q2p)                                # :407
        policy_loss = ((alpha * log_pi) - min_qp).mean()            # :409
        policy_loss.backward()                                      # :411-413
        agrads = {k: v.grad for k, v in ap.items()}
        self.actor_opt.step(self.actor, agrads)
        self.last_actor_grads = agrads

        if self.auto:                                               # :416-424
            la = self.log_alpha.clone().requires_grad_(True)
            alpha_loss = -(la * (log_pi + self.target_entropy).detach()).mean()
            alpha_loss.backward()
            pd = {"log_alpha": self.log_alpha}
            self.alpha_opt.step(pd, {"log_alpha": la.grad})
            self.alpha = float(self.log_alpha.exp())
            self.last_alpha_loss = float(alpha_loss.detach())

        if self.itera % self.policy_freq == 0:                      # :430-431
            soft_update(self.critic_target, self.critic, self.critic.keys(), self.tau)
        self.itera += 1
        self.last = dict(nq=nq, q1=q1.detach(), q2=q2.detach(), pi=pi.detach(),
                         log_pi=log_pi.detach(), q1p=q1p.detach(), q2p=q2p.detach())
        return float(qf1_loss.detach()), float(policy_loss.detach())                 # :437


    def learn_guidence(self, batch, noise, expert=None, engage_rows=None,
                       guidence_weight=1.0, engage_weight=1.0):
        """``SAC.learn_guidence`` on tensors — vn/DRL.py:187-301.

        ``batch``: the minibatch the reference builds at :199-226 (agent rows, then expert rows when
        ``expert`` is given).  ``expert``: dict obs,pobs,act of the expert rows (:259-263), ``engage_rows``:
        index tensor of the rows with engage == 1 (:266-273).  Extra ``noise`` keys: mask_g/eps_g (the
        policy.sample on the expert rows) and mask_e/eps_e (the one on the engaged rows), drawn by the
        reference right after mask_c_pi."""
        cfg = self.cfg
        s, ps, a = batch["obs"], batch["pobs"], batch["act"]
        r, s2, ps2 = batch["rew"], batch["next_obs"], batch["next_pobs"]
        alpha = self.alpha
        with torch.no_grad():                                       # :238-242
            a2, logp2, _ = actor_sample(self.actor, s2, ps2, noise["eps_next"], cfg, noise.get("mask_a_next"))
            q1t, q2t = critic_forward(self.critic_target, s2, ps2, a2, cfg, noise.get("mask_ct"))
            nq = r + self.gamma * (torch.min(q1t, q2t) - alpha * logp2)
        cp = {k: v.clone().requires_grad_(True) for k, v in self.critic.items()}
        q1, q2 = critic_forward(cp, s, ps, a, cfg, noise.get("mask_c"))      # :244
        qf1_loss = F.mse_loss(q1, nq)
        qf2_loss = F.mse_loss(q2, nq)
        (qf1_loss + qf2_loss).backward()                            # :249-251
        cgrads = {k: v.grad for k, v in cp.items()}
        self.critic_opt.step(self.critic, cgrads)
        self.last_critic_grads = cgrads
        ap = {k: v.clone().requires_grad_(True) for k, v in self.actor.items()}
        pi, log_pi, _ = actor_sample(ap, s, ps, noise["eps_pi"], cfg, noise.get("mask_a"))   # :253
        q1p, q2p = critic_forward(self.critic, s, ps, pi, cfg, noise.get("mask_c_pi"))       # :255
        min_qp = torch.min(q1p, q2p)
        guidence_loss = 0.0
        if expert is not None:                                      # :259-265
            _, _, pred = actor_sample(ap, expert["obs"], expert["pobs"], noise["eps_g"], cfg, noise.get("mask_g"))
            guidence_loss = guidence_weight * F.mse_loss(pred, expert["act"]).mean()
        engage_loss = 0.0
        if engage_rows is not None and engage_rows.numel() > 0:     # :268-276
            _, _, pred = actor_sample(ap, s[engage_rows], ps[engage_rows], noise["eps_e"], cfg, noise.get("mask_e"))
            engage_loss = engage_weight * F.mse_loss(pred, a[engage_rows]).mean()
        policy_loss = ((alpha * log_pi) - min_qp).mean() + guidence_loss + engage_loss       # :278
        policy_loss.backward()
        agrads = {k: v.grad for k, v in ap.items()}
        self.actor_opt.step(self.actor, agrads)
        self.last_actor_grads = agrads
        if self.auto:                                               # :285-293
            la = self.log_alpha.clone().requires_grad_(True)
            alpha_loss = -(la * (log_pi + self.target_entropy).detach()).mean()
            alpha_loss.backward()
            self.alpha_opt.step({"log_alpha": self.log_alpha}, {"log_alpha": la.grad})
            self.alpha = float(self.log_alpha.exp())
        if self.itera % self.policy_freq == 0:                      # :294-295
            soft_update(self.critic_target, self.critic, self.critic.keys(), self.tau)
        self.itera += 1
        return float(qf1_loss.detach()), float(policy_loss.detach())            # :299


# --------------------------------------------------------------------------
# Replay gather (cpprb semantics; vn/DRL.py:80-89,375-386)
# --------------------------------------------------------------------------

def replay_gather(store: Dict[str, np.ndarray], idx: np.ndarray, size: int):
    """Row gather from the ring store.  cpprb with ``next_of="obs"`` shares
    storage: ``next_obs[i]`` is ``obs[(i+1) % size]`` (cpprb is third-party and
    absent; semantics restated from its documented behaviour)."""
    out = {}
    for k, v in store.items():
        out[k] = v[idx]
    out["next_obs"] = store["obs"][(idx + 1) % size]
    return out


# --------------------------------------------------------------------------
# Depth normalise + noise augmentation + resize (vn/env_lab.py)
# --------------------------------------------------------------------------

def _reflect101(i, n):
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * (n - 1) - i
    return i


def _sep_blur(img: np.ndarray, k: np.ndarray) -> np.ndarray:
    """Separable filter, BORDER_REFLECT_101 (cv2.GaussianBlur default), float64."""
    r = len(k) // 2
    h, w = img.shape
    cols = np.array([[_reflect101(x + d, w) for d in range(-r, r + 1)] for x in range(w)])
    tmp = (img[:, cols] * k[None, None, :]).sum(-1)
    rows = np.array([[_reflect101(y + d, h) for d in range(-r, r + 1)] for y in range(h)])
    return (tmp[rows, :] * k[None, :, None]).sum(1)


def gaussian_kernel(ksize: int) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma<=0): fixed tables for ksize<=7, else
    sigma = 0.3*((ksize-1)*0.5-1)+0.8 and exp(-x^2/(2 sigma^2)) normalised."""
    if ksize == 5:
        return np.array([0.0625, 0.25, 0.375, 0.25, 0.0625])
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    x = np.arange(ksize) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2 * sigma * sigma))
    return k / k.sum()


def depth_augment(raw: np.ndarray, noise: np.ndarray, out_hw=(128, 160)) -> np.ndarray:
    """vn/env_lab.py:420-434 callback → :78-90 add_nose(50) → :69-76 blurring →
    :295/:348 resize → /255.  ``raw`` f32 [H,W]; ``noise`` f64 [H,W] ~ N(0,50).
    Returns float64 [h,w] in [0,1] (the reference keeps float64 until cpprb
    stores it as float32)."""
    raw = raw.astype(np.float32)
    mn, mx = float(raw.min()), float(raw.max())
    # cv2.normalize NORM_MINMAX to [0,255]: scale = 255/(max-min), shift = -min*scale (f32 dst)
    scale = (255.0 / (mx - mn)) if mx - mn > np.finfo(np.float64).eps else 0.0
    shift = 0.0 - mn * scale
    norm = (raw.astype(np.float64) * scale + shift).astype(np.float32)
    u8 = norm.astype(np.uint8)                                     # :425 truncation
    img = np.clip(u8.astype(np.float32) + noise, 0, 255)           # :86-88 (float64)
    img = _sep_blur(img, gaussian_kernel(5))                       # :89
    h = img.shape[0]
    bh = h // 5                                                    # :33-39
    y1 = h // 2 - bh // 2
    y2 = y1 + bh
    band = _sep_blur(img[y1:y2].copy(), gaussian_kernel(11))       # :74 (border reflects inside the band)
    img = img.copy()
    img[y1:y2] = band
    oh, ow = out_hw
    fy, fx = img.shape[0] // oh, img.shape[1] // ow
    assert fy * oh == img.shape[0] and fx * ow == img.shape[1] and fy == fx and fy % 2 == 0
    # cv2.resize INTER_LINEAR with an even integer factor f: sample centre falls between
    # pixels f*i + f/2 - 1 and f*i + f/2 with weight 0.5 each, in both axes.
    o = fy // 2 - 1
    a = img[o::fy, o::fx][:oh, :ow]
    b = img[o::fy, o + 1::fx][:oh, :ow]
    c = img[o + 1::fy, o::fx][:oh, :ow]
    d = img[o + 1::fy, o + 1::fx][:oh, :ow]
    res = 0.25 * a + 0.25 * b + 0.25 * c + 0.25 * d
    return res / 255.0                                             # :299
