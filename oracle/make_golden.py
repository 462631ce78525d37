"""Pin the oracle against the live reference and write tests/golden/*.npz.

Run HERE (the container that mounts /root/reference); the reference tree does
not travel to the GPU box, the fixtures do.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden

What is checked (hard asserts) and what is stored:

1. ``reference_init`` == the reference constructors under the same seed
   (bit-for-bit) -> weight checksums stored.
2. Actor ``forward``/``sample`` and critic ``forward`` of the *imported*
   reference modules (vn/got_sac_network.py) vs the oracle: eval mode, and
   train mode with the dropout mask / rsample noise replayed from the same
   generator state -> reference outputs stored.
3. The UNMODIFIED reference ``SAC.learn`` (vn/DRL.py:373-437), imported with a
   stub ``cpprb`` that returns a fixed minibatch, vs ``SACOracle.learn`` for 3
   steps: losses, gradients, parameters after each step -> stored summaries.
4. The reference depth pipeline (source slices of vn/env_lab.py exec'd with
   this image's cv2) vs ``depth_augment`` -> stored small frames.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src/vis_nav/vis_nav"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import dgvit_oracle as O          # noqa: E402
from oracle.init_params import reference_init, reference_sac_init, synthetic_batch   # noqa: E402

SEED = 3407   # vn/config.yaml:7


# ---------------------------------------------------------------- reference import
class _StubPER:
    """Minimal stand-in for cpprb.PrioritizedReplayBuffer (absent here): ``sample``
    returns the minibatch planted by the test.  Index selection is an *input* of
    the hot path, so nothing numeric is stubbed."""
    planted = None

    def __init__(self, size, env_dict=None, next_of=None, **kw):
        self.size = size

    def sample(self, batch_size):
        return {k: v.copy() for k, v in _StubPER.planted.items()}

    def add(self, **kw):
        pass

    def get_stored_size(self):
        return self.size


def import_reference():
    sys.path.insert(0, REF)
    stub = types.ModuleType("cpprb")
    stub.PrioritizedReplayBuffer = _StubPER
    sys.modules["cpprb"] = stub
    import got_sac_network as G          # noqa
    import DRL as D                      # noqa
    return G, D


def params_of(module):
    return {k: v.detach().clone() for k, v in module.named_parameters()}


def checksum(d):
    names = list(d.keys())
    s = np.array([float(d[k].double().sum()) for k in names])
    a = np.array([float(d[k].double().abs().sum()) for k in names])
    return names, s, a


def maxrel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ---------------------------------------------------------------- cases
def case_modules(G, tag, block, head, lfs, B):
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    torch.manual_seed(SEED)
    ref_actor = G.GoTPolicy(2, 2, block, head, lfs)
    torch.manual_seed(SEED + 1)
    ref_critic = G.GoTQNetwork(2, 2, block, head, lfs)
    pa, pc = params_of(ref_actor), params_of(ref_critic)
    ia = reference_init("actor", cfg, SEED)
    ic = reference_init("critic", cfg, SEED + 1)
    assert list(ia.keys()) == list(pa.keys()), "actor param order"
    assert list(ic.keys()) == list(pc.keys()), "critic param order"
    for k in pa:
        assert torch.equal(pa[k], ia[k]), f"actor init {k}"
    for k in pc:
        assert torch.equal(pc[k], ic[k]), f"critic init {k}"

    batch = synthetic_batch(cfg, B, SEED + 2)
    img, goal, act = batch["obs"], batch["pobs"], batch["act"]
    out = {}
    # eval mode
    ref_actor.eval(); ref_critic.eval()
    with torch.no_grad():
        m, ls = ref_actor.forward([img, goal])
        q1, q2 = ref_critic.forward([img, goal, act])
        om, ols = O.actor_forward(pa, img, goal, cfg)
        oq1, oq2 = O.critic_forward(pc, img, goal, act, cfg)
    for n, (r, o) in dict(mean=(m, om), log_std=(ls, ols), q1=(q1, oq1), q2=(q2, oq2)).items():
        e = maxrel(o, r)
        assert e < 1e-5, (tag, "eval", n, e)
        out["eval_" + n] = r.numpy()
    # train mode, replayed generator
    ref_actor.train(); ref_critic.train()
    torch.manual_seed(SEED + 3)
    with torch.no_grad():
        a, lp, mt = ref_actor.sample([img, goal])
        q1, q2 = ref_critic.forward([img, goal, a])
    torch.manual_seed(SEED + 3)
    mask_a = torch.empty(B, cfg.n_tokens, cfg.dim).bernoulli_(1 - O.EMB_DROPOUT)
    eps = torch.empty(B, 2).normal_()
    mask_c = torch.empty(B, cfg.n_tokens, cfg.dim).bernoulli_(1 - O.EMB_DROPOUT)
    with torch.no_grad():
        oa, olp, omt = O.actor_sample(pa, img, goal, eps, cfg, mask_a)
        oq1, oq2 = O.critic_forward(pc, img, goal, oa, cfg, mask_c)
    for n, (r, o) in dict(action=(a, oa), log_prob=(lp, olp), mean_t=(mt, omt), q1=(q1, oq1), q2=(q2, oq2)).items():
        e = maxrel(o, r)
        assert e < 2e-5, (tag, "train", n, e)
        out["train_" + n] = r.numpy()
    out["train_mask_a"] = np.packbits(mask_a.numpy().astype(np.uint8))
    out["train_mask_c"] = np.packbits(mask_c.numpy().astype(np.uint8))
    out["train_eps"] = eps.numpy()
    # gradient parity (train-mode graph, injected noise == replayed generator)
    ref_actor.zero_grad(); ref_critic.zero_grad()
    torch.manual_seed(SEED + 3)
    a, lp, mt = ref_actor.sample([img, goal])
    q1, q2 = ref_critic.forward([img, goal, a])
    loss = (lp.mean() * 0.3 - torch.min(q1, q2).mean()) + (q1 ** 2).mean() * 0.1
    loss.backward()
    pa_g = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    pc_g = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    oa, olp, omt = O.actor_sample(pa_g, img, goal, eps, cfg, mask_a)
    oq1, oq2 = O.critic_forward(pc_g, img, goal, oa, cfg, mask_c)
    oloss = (olp.mean() * 0.3 - torch.min(oq1, oq2).mean()) + (oq1 ** 2).mean() * 0.1
    oloss.backward()
    for mod, og_d in ((ref_actor, pa_g), (ref_critic, pc_g)):
        for k, rp in mod.named_parameters():
            og = og_d[k].grad
            if rp.grad is None:
                assert og is None, k
                continue
            e = maxrel(og, rp.grad)
            assert e < 2e-4, (tag, "grad", k, e)
    out["grad_loss"] = np.array(float(loss.detach()))
    out["grad_actor_norms"] = np.array([0.0 if p.grad is None else float(p.grad.double().norm())
                                        for _, p in ref_actor.named_parameters()])
    out["grad_critic_norms"] = np.array([0.0 if p.grad is None else float(p.grad.double().norm())
                                         for _, p in ref_critic.named_parameters()])
    for nm, d in (("actor", pa), ("critic", pc)):
        names, s, ab = checksum(d)
        out[f"{nm}_names"] = np.array(names)
        out[f"{nm}_sum"], out[f"{nm}_abssum"] = s, ab
    out["cfg"] = np.array([lfs, block, head, B])
    np.savez_compressed(os.path.join(GOLD, f"modules_{tag}.npz"), **out)
    print(f"[golden] modules_{tag}: ok")


def case_qnet(G, B):
    """CNN twin-Q critic ``QNetwork`` (vn/got_sac_network.py:125-170): init, forward and gradients vs the oracle."""
    from oracle.init_params import reference_qnet_init
    cfg = O.Cfg()
    torch.manual_seed(SEED + 7)
    ref = G.QNetwork(2, 2)
    pr = params_of(ref)
    io = reference_qnet_init(SEED + 7)
    assert list(io.keys()) == list(pr.keys()), "qnet param order"
    for k in pr:
        assert torch.equal(pr[k], io[k]), f"qnet init {k}"
    batch = synthetic_batch(cfg, B, SEED + 8)
    img, goal, act = batch["obs"], batch["pobs"], batch["act"]
    q1, q2 = ref.forward([img, goal, act])
    w = torch.linspace(0.5, 1.5, B).unsqueeze(1)
    loss = ((q1 - 0.3) ** 2 * w).mean() + (torch.min(q1, q2) * w).mean()
    ref.zero_grad()
    loss.backward()
    pg = {k: v.clone().requires_grad_(True) for k, v in pr.items()}
    oq1, oq2 = O.qnet_forward(pg, img, goal, act)
    oloss = ((oq1 - 0.3) ** 2 * w).mean() + (torch.min(oq1, oq2) * w).mean()
    oloss.backward()
    out = {}
    for n, (r, o) in dict(q1=(q1, oq1), q2=(q2, oq2)).items():
        e = maxrel(o, r)
        assert e < 1e-5, ("qnet", n, e)
        out[n] = r.detach().numpy()
    for k, rp in ref.named_parameters():
        e = maxrel(pg[k].grad, rp.grad)
        assert e < 2e-4, ("qnet grad", k, e)
    out["loss"] = np.array(float(loss.detach()))
    out["grad_norms"] = np.array([float(p.grad.double().norm()) for _, p in ref.named_parameters()])
    out["grad_fc1_bias"] = dict(ref.named_parameters())["fc1.bias"].grad.numpy()
    out["grad_conv1_weight"] = dict(ref.named_parameters())["conv1.weight"].grad.numpy()
    names, sm, ab = checksum(pr)
    out["names"], out["sum"], out["abssum"] = np.array(names), sm, ab
    out["cfg"] = np.array([B])
    np.savez_compressed(os.path.join(GOLD, "qnet.npz"), **out)
    print("[golden] qnet: ok")


def case_learn(G, D, tag, block, head, lfs, B, steps=3):
    """Unmodified reference SAC.learn vs SACOracle.learn."""
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    agent = D.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, SEED,
                  LR_C=1e-3, LR_A=1e-3, LR_ALPHA=1e-4, BUFFER_SIZE=64, TAU=5e-4, POLICY_FREQ=1,
                  GAMMA=0.999, ALPHA=1.0, block=block, head=head, l_f_size=lfs,
                  automatic_entropy_tuning=True)
    assert str(agent.device) == "cpu"
    orc = O.SACOracle(params_of(agent.policy), params_of(agent.critic), cfg)
    ia, ic = reference_sac_init(cfg, SEED)
    for k, v in agent.policy.named_parameters():
        assert torch.equal(v.detach(), ia[k]), k
    for k, v in agent.critic.named_parameters():
        assert torch.equal(v.detach(), ic[k]), k
    # SAC.__init__ seeds with SEED then builds critic, critic_target, policy: record how
    # to rebuild the same weights from reference_init-style construction order.
    out = {}
    for s in range(steps):
        batch = synthetic_batch(cfg, B, SEED + 10 + s)
        _StubPER.planted = {k: v.numpy() for k, v in batch.items()}
        torch.manual_seed(SEED + 100 + s)
        l_ref = agent.learn(B)
        torch.manual_seed(SEED + 100 + s)
        shp = (B, cfg.n_tokens, cfg.dim)
        noise = {}
        noise["mask_a_next"] = torch.empty(shp).bernoulli_(0.9)
        noise["eps_next"] = torch.empty(B, 2).normal_()
        noise["mask_ct"] = torch.empty(shp).bernoulli_(0.9)
        noise["mask_c"] = torch.empty(shp).bernoulli_(0.9)
        noise["mask_a"] = torch.empty(shp).bernoulli_(0.9)
        noise["eps_pi"] = torch.empty(B, 2).normal_()
        noise["mask_c_pi"] = torch.empty(shp).bernoulli_(0.9)
        l_orc = orc.learn(batch, noise)
        for a_, b_ in zip(l_ref, l_orc):
            assert abs(a_ - b_) <= 1e-5 * max(1.0, abs(a_)), (tag, s, l_ref, l_orc)
        # parameters after the step
        for nm, mod, od in (("actor", agent.policy, orc.actor), ("critic", agent.critic, orc.critic),
                            ("target", agent.critic_target, orc.critic_target)):
            worst = 0.0
            for k, p in mod.named_parameters():
                worst = max(worst, float((p.detach() - od[k]).abs().max()))
            # Adam's g/(|g|+eps) amplifies rounding where |g| ~ 1e-8; allow a loose abs bound
            assert worst < 2e-3 * (s + 1), (tag, s, nm, worst)
            names, sm, ab = checksum(params_of(mod))
            out[f"step{s}_{nm}_sum"], out[f"step{s}_{nm}_abssum"] = sm, ab
        la_ref = float(agent.log_alpha.detach())
        assert abs(la_ref - float(orc.log_alpha)) < 1e-7
        out[f"step{s}_losses"] = np.array(l_ref)
        out[f"step{s}_log_alpha"] = np.array(la_ref)
        out[f"step{s}_noise_bits"] = np.concatenate([np.packbits(noise[k].numpy().astype(np.uint8))
                                                     for k in ("mask_a_next", "mask_ct", "mask_c", "mask_a", "mask_c_pi")])
        out[f"step{s}_eps"] = np.stack([noise["eps_next"].numpy(), noise["eps_pi"].numpy()])
        if s == 0:
            out["step0_critic_grad_norms"] = np.array(
                [0.0 if p.grad is None else 1.0 for _, p in agent.critic.named_parameters()])
            out["step0_nq"] = orc.last["nq"].numpy()
            out["step0_q1"] = orc.last["q1"].numpy()
            out["step0_pi"] = orc.last["pi"].numpy()
            out["step0_log_pi"] = orc.last["log_pi"].numpy()
            out["step0_actor_gnorm"] = np.array([0.0 if g is None else float(g.double().norm())
                                                 for g in orc.last_actor_grads.values()])
            out["step0_critic_gnorm"] = np.array([0.0 if g is None else float(g.double().norm())
                                                  for g in orc.last_critic_grads.values()])
            # grads that the reference leaves as None (Adam skips them): verified identical sets
            ref_none = [k for k, p in agent.policy.named_parameters() if p.grad is None]
            orc_none = [k for k, g in orc.last_actor_grads.items() if g is None]
            assert ref_none == orc_none, (ref_none, orc_none)
            out["actor_unused"] = np.array(ref_none)
    out["cfg"] = np.array([lfs, block, head, B, steps])
    np.savez_compressed(os.path.join(GOLD, f"learn_{tag}.npz"), **out)
    print(f"[golden] learn_{tag}: ok")


def case_guidence(G, D, tag, block, head, lfs, B, Be, steps=2):
    """Unmodified reference SAC.learn_guidence (expert buffer + engage rows) vs SACOracle.learn_guidence."""
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    agent = D.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, True, SEED,
                  LR_C=1e-3, LR_A=1e-3, LR_ALPHA=1e-4, BUFFER_SIZE=64, TAU=5e-4, POLICY_FREQ=1,
                  GAMMA=0.999, ALPHA=1.0, block=block, head=head, l_f_size=lfs, buffer_size_expert=63,
                  automatic_entropy_tuning=True)
    orc = O.SACOracle(params_of(agent.policy), params_of(agent.critic), cfg)
    out = {}
    # get_stored_size(): agent 64, expert 64 -> batch_expert = min(floor(64/64*B), B) = B  (vn/DRL.py:193-196);
    # the stub returns Be expert rows whatever is asked: plant Be = B
    assert Be == B
    for s in range(steps):
        ba = synthetic_batch(cfg, B, SEED + 30 + s)
        be = synthetic_batch(cfg, Be, SEED + 40 + s)
        engage = np.zeros((B, 1), np.float32)
        engage[[1, B - 1], 0] = 1.0
        agent_dict = {k: v.numpy() for k, v in ba.items()}
        agent_dict["engage"] = engage
        expert_dict = {k: v.numpy() for k, v in be.items()}
        expert_dict["act_exp"] = expert_dict.pop("act")
        planted = [agent_dict, expert_dict]

        class _Seq:
            i = 0
        def sample_seq(self_, n, _p=planted, _c=_Seq):
            d = _p[_c.i % 2]
            _c.i += 1
            return {k: v.copy() for k, v in d.items()}
        _StubPER.sample = sample_seq
        torch.manual_seed(SEED + 300 + s)
        l_ref = agent.learn_guidence(False, B)
        torch.manual_seed(SEED + 300 + s)
        Bc = B + Be
        shp = lambda n: (n, cfg.n_tokens, cfg.dim)
        noise = {}
        noise["mask_a_next"] = torch.empty(shp(Bc)).bernoulli_(0.9)
        noise["eps_next"] = torch.empty(Bc, 2).normal_()
        noise["mask_ct"] = torch.empty(shp(Bc)).bernoulli_(0.9)
        noise["mask_c"] = torch.empty(shp(Bc)).bernoulli_(0.9)
        noise["mask_a"] = torch.empty(shp(Bc)).bernoulli_(0.9)
        noise["eps_pi"] = torch.empty(Bc, 2).normal_()
        noise["mask_c_pi"] = torch.empty(shp(Bc)).bernoulli_(0.9)
        noise["mask_g"] = torch.empty(shp(Be)).bernoulli_(0.9)
        noise["eps_g"] = torch.empty(Be, 2).normal_()
        noise["mask_e"] = torch.empty(shp(2)).bernoulli_(0.9)
        noise["eps_e"] = torch.empty(2, 2).normal_()
        cat = {k: torch.cat([ba[k], be[k]], 0) for k in ba}
        l_orc = orc.learn_guidence(cat, noise, expert=dict(obs=be["obs"], pobs=be["pobs"], act=be["act"]),
                                   engage_rows=torch.tensor([1, B - 1]))
        for a_, b_ in zip(l_ref, l_orc):
            assert abs(a_ - b_) <= 1e-5 * max(1.0, abs(a_)), (tag, s, l_ref, l_orc)
        for nm, mod, od in (("actor", agent.policy, orc.actor), ("critic", agent.critic, orc.critic)):
            worst = max(float((p.detach() - od[k]).abs().max()) for k, p in mod.named_parameters())
            assert worst < 2e-3 * (s + 1), (tag, s, nm, worst)
            names, sm, ab = checksum(params_of(mod))
            out[f"step{s}_{nm}_abssum"] = ab
        out[f"step{s}_losses"] = np.array(l_ref)
        out[f"step{s}_log_alpha"] = np.array(float(agent.log_alpha.detach()))
        out[f"step{s}_noise_bits"] = np.concatenate([np.packbits(noise[k].numpy().astype(np.uint8)) for k in
                                                     ("mask_a_next", "mask_ct", "mask_c", "mask_a", "mask_c_pi", "mask_g", "mask_e")])
        out[f"step{s}_eps"] = np.concatenate([noise[k].numpy().reshape(-1) for k in ("eps_next", "eps_pi", "eps_g", "eps_e")])
    del _StubPER.sample
    _StubPER.sample = lambda self, n: {k: v.copy() for k, v in _StubPER.planted.items()}
    out["cfg"] = np.array([lfs, block, head, B, Be, steps])
    np.savez_compressed(os.path.join(GOLD, f"guidence_{tag}.npz"), **out)
    print(f"[golden] guidence_{tag}: ok")


def case_depth():
    """vn/env_lab.py source slices executed with this image's cv2."""
    import cv2
    src = open(os.path.join(REF, "env_lab.py")).read().split("\n")
    code = "\n".join(src[32:39] + src[68:90])      # get_center_band, blurring, add_nose (lines 33-39, 69-90)
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    ns = {"np": np, "cv2": cv2}
    exec(code, ns)
    rng = np.random.RandomState(SEED)
    out = {}
    for i, (H, W) in enumerate([(512, 640), (256, 320)]):
        yy, xx = np.mgrid[0:H, 0:W]
        raw = (0.03 + 7.97 * (0.5 + 0.5 * np.sin(xx / 37.0 + i) * np.cos(yy / 53.0))).astype(np.float32)
        raw[H // 3: H // 3 + 40, W // 4: W // 4 + 90] = 1.25
        # --- reference pipeline, vn/env_lab.py:420-434 then :295-299
        dn = cv2.normalize(raw, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
        np.random.seed(SEED + i)
        noise = np.random.normal(0, 50, dn.shape)
        np.random.seed(SEED + i)
        im = ns["add_nose"](dn, noise_level=50)
        im = ns["blurring"](im)
        state = cv2.resize(im, (W // 4, H // 4)) / 255
        got = O.depth_augment(raw, noise, out_hw=(H // 4, W // 4))
        err = float(np.abs(got - state).max())
        assert err < 1e-9, ("depth", i, err)
        out[f"raw_{i}_params"] = np.array([H, W, i])
        out[f"noise_seed_{i}"] = np.array(SEED + i)
        out[f"state_{i}"] = state.astype(np.float64)
    np.savez_compressed(os.path.join(GOLD, "depth_aug.npz"), **out)
    print("[golden] depth_aug: ok")


def main():
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "qnet":       # regenerate one fixture only
        G, _ = import_reference()
        return case_qnet(G, B=4)
    torch.set_num_threads(8)
    G, D = import_reference()
    case_modules(G, "small", block=2, head=2, lfs=32, B=3)
    case_modules(G, "shipped", block=4, head=4, lfs=64, B=4)
    case_learn(G, D, "small", block=2, head=2, lfs=32, B=4)
    case_learn(G, D, "shipped", block=4, head=4, lfs=64, B=4)
    case_guidence(G, D, "small", block=2, head=2, lfs=32, B=4, Be=4)
    case_depth()
    case_qnet(G, B=4)


if __name__ == "__main__":
    main()
