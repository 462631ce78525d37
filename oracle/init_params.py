"""Seeded random init + seeded synthetic inputs for the oracle.  TEST INFRASTRUCTURE ONLY.

``reference_init`` rebuilds the parameter dict of the reference actor / critic
by creating ``torch.nn`` layers in the *same order* as the reference
constructors, so that under the same ``torch.manual_seed`` the same generator
draws land in the same tensors:

* ``GoT.__init__``        vn/GoalFormer.py:124-154  (patch Linear, pos_embedding,
  cls_token, per block: to_qkv, to_out, ff.0, ff.3; mlp_head)
* ``GoTPolicy.__init__``  vn/got_sac_network.py:173-194 (trans, fc_embed, fc1, fc2,
  mean_linear, log_std_linear, then ``apply(weights_init_)``)
* ``GoTQNetwork.__init__`` vn/got_sac_network.py:76-105 (trans, conv1-3, fc1, fc2,
  fc3, fc_embed, fc11, fc21, fc31, then ``apply(weights_init_)``)
* ``weights_init_``       vn/got_sac_network.py:30-33 (Xavier-uniform gain 1 on
  every ``nn.Linear.weight``, module post-order == registration order here).

The returned dict is keyed and ordered like ``module.named_parameters()`` of the
reference (own parameters first, then children in registration order).
``oracle/make_golden.py`` checks it bit-for-bit against the imported reference.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from .dgvit_oracle import Cfg


def _trunk_layers(cfg: Cfg):
    """Creation order == RNG order (vn/GoalFormer.py:137-154)."""
    made = OrderedDict()
    made["to_patch_embedding.1"] = nn.Linear(cfg.patch_dim, cfg.dim)        # :139 (hard-coded 320)
    made["pos_embedding"] = nn.Parameter(torch.randn(1, cfg.n_patches + 1, cfg.dim))   # :142
    made["cls_token"] = nn.Parameter(torch.randn(1, 1, cfg.dim))            # :143
    for l in range(cfg.depth):                                              # :97-100
        pre = f"transformer.layers.{l}."
        made[pre + "0.fn.to_qkv"] = nn.Linear(cfg.dim, cfg.inner * 3, bias=False)   # :64
        made[pre + "0.fn.to_out.0"] = nn.Linear(cfg.inner, cfg.dim)         # :67
        made[pre + "0.norm"] = nn.LayerNorm(cfg.dim)                        # :34
        made[pre + "1.fn.net.0"] = nn.Linear(cfg.dim, cfg.mlp_dim)          # :43
        made[pre + "1.fn.net.3"] = nn.Linear(cfg.mlp_dim, cfg.dim)          # :46
        made[pre + "1.norm"] = nn.LayerNorm(cfg.dim)
    made["mlp_head.0"] = nn.LayerNorm(cfg.dim)                              # :152
    made["mlp_head.1"] = nn.Linear(cfg.dim, 2)                              # :153 (num_classes=2)
    return made


def _trunk_named(made, cfg: Cfg):
    """named_parameters() order of GoT: own params, then children as registered
    (layer_norm, to_patch_embedding, transformer, mlp_head)."""
    out = OrderedDict()
    out["trans.pos_embedding"] = made["pos_embedding"]
    out["trans.cls_token"] = made["cls_token"]
    out["trans.layer_norm.g"] = nn.Parameter(torch.ones(cfg.dim))            # :117-118
    out["trans.to_patch_embedding.1.weight"] = made["to_patch_embedding.1"].weight
    out["trans.to_patch_embedding.1.bias"] = made["to_patch_embedding.1"].bias
    for l in range(cfg.depth):
        pre = f"transformer.layers.{l}."
        out[f"trans.{pre}0.norm.weight"] = made[pre + "0.norm"].weight
        out[f"trans.{pre}0.norm.bias"] = made[pre + "0.norm"].bias
        out[f"trans.{pre}0.fn.to_qkv.weight"] = made[pre + "0.fn.to_qkv"].weight
        out[f"trans.{pre}0.fn.to_out.0.weight"] = made[pre + "0.fn.to_out.0"].weight
        out[f"trans.{pre}0.fn.to_out.0.bias"] = made[pre + "0.fn.to_out.0"].bias
        out[f"trans.{pre}1.norm.weight"] = made[pre + "1.norm"].weight
        out[f"trans.{pre}1.norm.bias"] = made[pre + "1.norm"].bias
        out[f"trans.{pre}1.fn.net.0.weight"] = made[pre + "1.fn.net.0"].weight
        out[f"trans.{pre}1.fn.net.0.bias"] = made[pre + "1.fn.net.0"].bias
        out[f"trans.{pre}1.fn.net.3.weight"] = made[pre + "1.fn.net.3"].weight
        out[f"trans.{pre}1.fn.net.3.bias"] = made[pre + "1.fn.net.3"].bias
    out["trans.mlp_head.0.weight"] = made["mlp_head.0"].weight
    out["trans.mlp_head.0.bias"] = made["mlp_head.0"].bias
    out["trans.mlp_head.1.weight"] = made["mlp_head.1"].weight
    out["trans.mlp_head.1.bias"] = made["mlp_head.1"].bias
    return out


def _xavier_linears(named: "OrderedDict[str, nn.Module]"):
    for m in named.values():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight, gain=1)


def reference_init(kind: str, cfg: Cfg, seed) -> "OrderedDict[str, torch.Tensor]":
    """kind in {"actor", "critic"}; returns detached fp32 tensors.  ``seed=None``
    continues the current global generator stream."""
    if seed is not None:
        torch.manual_seed(seed)
    made = _trunk_layers(cfg)
    out = _trunk_named(made, cfg)
    heads = OrderedDict()
    if kind == "actor":
        heads["fc_embed"] = nn.Linear(cfg.nb_pstate, cfg.dim)
        heads["fc1"] = nn.Linear(cfg.dim, 128)
        heads["fc2"] = nn.Linear(128, 128)
        heads["mean_linear"] = nn.Linear(128, cfg.nb_actions)
        heads["log_std_linear"] = nn.Linear(128, cfg.nb_actions)
    elif kind == "critic":
        heads["conv1"] = nn.Conv2d(4, 16, 5, stride=2)
        heads["conv2"] = nn.Conv2d(16, 64, 5, stride=2)
        heads["conv3"] = nn.Conv2d(64, 256, 5, stride=2)
        heads["fc1"] = nn.Linear(cfg.dim + cfg.nb_actions, 128)
        heads["fc2"] = nn.Linear(128, 32)
        heads["fc3"] = nn.Linear(32, cfg.nb_actions)
        heads["fc_embed"] = nn.Linear(cfg.nb_pstate, cfg.dim)
        heads["fc11"] = nn.Linear(cfg.dim + cfg.nb_actions, 128)
        heads["fc21"] = nn.Linear(128, 32)
        heads["fc31"] = nn.Linear(32, cfg.nb_actions)
    else:
        raise ValueError(kind)
    # apply(weights_init_): children first in registration order -> trunk linears
    # (registration order: to_patch_embedding, transformer.*, mlp_head), then heads.
    reg = OrderedDict()
    reg["to_patch_embedding.1"] = made["to_patch_embedding.1"]
    for l in range(cfg.depth):
        pre = f"transformer.layers.{l}."
        for k in ("0.fn.to_qkv", "0.fn.to_out.0", "1.fn.net.0", "1.fn.net.3"):
            reg[pre + k] = made[pre + k]
    reg["mlp_head.1"] = made["mlp_head.1"]
    _xavier_linears(reg)
    _xavier_linears(heads)
    for name, m in heads.items():
        out[name + ".weight"] = m.weight
        out[name + ".bias"] = m.bias
    return OrderedDict((k, v.detach().clone()) for k, v in out.items())


def reference_qnet_init(seed, nb_actions: int = 2, nb_pstate: int = 2) -> "OrderedDict[str, torch.Tensor]":
    """Parameters of the CNN critic as ``QNetwork.__init__`` creates them (vn/got_sac_network.py:126-147):
    conv1-3 (default Conv2d init), fc1, fc2, fc3, fc_embed, fc11, fc21, fc31, then ``apply(weights_init_)``."""
    if seed is not None:
        torch.manual_seed(seed)
    m = OrderedDict()
    m["conv1"] = nn.Conv2d(1, 16, 5, stride=2)
    m["conv2"] = nn.Conv2d(16, 64, 5, stride=2)
    m["conv3"] = nn.Conv2d(64, 256, 5, stride=2)
    m["fc1"] = nn.Linear(256 + 32 + nb_actions, 128)
    m["fc2"] = nn.Linear(128, 32)
    m["fc3"] = nn.Linear(32, nb_actions)
    m["fc_embed"] = nn.Linear(nb_pstate, 32)
    m["fc11"] = nn.Linear(256 + 32 + nb_actions, 128)
    m["fc21"] = nn.Linear(128, 32)
    m["fc31"] = nn.Linear(32, nb_actions)
    _xavier_linears(m)
    out = OrderedDict()
    for name, mod in m.items():
        out[name + ".weight"] = mod.weight
        out[name + ".bias"] = mod.bias
    return OrderedDict((k, v.detach().clone()) for k, v in out.items())


def reference_sac_init(cfg: Cfg, seed: int):
    """Weights as ``SAC.__init__`` creates them (vn/DRL.py:71-78 seeding, :105-106 critic,
    :115-116 critic_target (consumes the generator, then overwritten by hard_update :123),
    :142-143 policy).  Returns (actor, critic)."""
    torch.manual_seed(seed)
    critic = reference_init("critic", cfg, None)
    reference_init("critic", cfg, None)          # critic_target draws
    actor = reference_init("actor", cfg, None)
    return actor, critic


def synthetic_batch(cfg: Cfg, B: int, seed: int):
    """Seeded synthetic minibatch: U(0,1) depth frames, goals
    (d~U(0,1), heading~U(-1,1)), actions U(-1,1)^2, rewards clip(N(0,20),-200,500),
    done~Bernoulli(0.01)."""
    g = torch.Generator().manual_seed(seed)
    obs = torch.rand(B, cfg.img_h, cfg.img_w, generator=g)
    next_obs = torch.rand(B, cfg.img_h, cfg.img_w, generator=g)

    def goal():
        d = torch.rand(B, 1, generator=g)
        h = torch.rand(B, 1, generator=g) * 2 - 1
        return torch.cat([d, h], dim=1)

    pobs, next_pobs = goal(), goal()
    act = torch.rand(B, cfg.nb_actions, generator=g) * 2 - 1
    rew = (torch.randn(B, 1, generator=g) * 20).clamp(-200, 500)
    done = (torch.rand(B, 1, generator=g) < 0.01).float()
    return dict(obs=obs, next_obs=next_obs, pobs=pobs, next_pobs=next_pobs, act=act, rew=rew, done=done)


def synthetic_noise(cfg: Cfg, B: int, seed: int, dropout: bool = True):
    """Seeded stochastic inputs of one ``learn`` (masks are {0,1} keep-masks)."""
    g = torch.Generator().manual_seed(seed)
    out = dict(eps_next=torch.randn(B, cfg.nb_actions, generator=g),
               eps_pi=torch.randn(B, cfg.nb_actions, generator=g))
    for k in ("mask_a_next", "mask_ct", "mask_c", "mask_a", "mask_c_pi"):
        out[k] = (torch.rand(B, cfg.n_tokens, cfg.dim, generator=g) >= 0.1).float() if dropout else None
    return out
