"""CPU: the oracle against the committed golden vectors (made from the live reference by
oracle/make_golden.py).  No reference tree, no GPU needed."""
import numpy as np
import pytest
import torch

from helpers import O, SEED, golden, reference_init, reference_sac_init, relerr, synthetic_batch, unpack_mask


@pytest.mark.parametrize("tag,block,head,lfs,B", [("small", 2, 2, 32, 3), ("shipped", 4, 4, 64, 4)])
def test_modules_against_golden(tag, block, head, lfs, B):
    g = golden(f"modules_{tag}.npz")
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    pa = reference_init("actor", cfg, SEED)
    pc = reference_init("critic", cfg, SEED + 1)
    # weights are the reference's (checksums recorded from the reference constructors)
    for nm, d in (("actor", pa), ("critic", pc)):
        assert list(d.keys()) == [str(s) for s in g[f"{nm}_names"]]
        s = np.array([float(v.double().sum()) for v in d.values()])
        a = np.array([float(v.double().abs().sum()) for v in d.values()])
        np.testing.assert_allclose(s, g[f"{nm}_sum"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(a, g[f"{nm}_abssum"], rtol=0, atol=1e-9)
    batch = synthetic_batch(cfg, B, SEED + 2)
    img, goal, act = batch["obs"], batch["pobs"], batch["act"]
    with torch.no_grad():
        m, ls = O.actor_forward(pa, img, goal, cfg)
        q1, q2 = O.critic_forward(pc, img, goal, act, cfg)
    for n, v in dict(mean=m, log_std=ls, q1=q1, q2=q2).items():
        assert relerr(v, g["eval_" + n]) < 1e-5, n
    shp = (B, cfg.n_tokens, cfg.dim)
    mask_a, mask_c = unpack_mask(g["train_mask_a"], shp), unpack_mask(g["train_mask_c"], shp)
    eps = torch.from_numpy(g["train_eps"])
    pa_g = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    pc_g = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    a, lp, mt = O.actor_sample(pa_g, img, goal, eps, cfg, mask_a)
    q1, q2 = O.critic_forward(pc_g, img, goal, a, cfg, mask_c)
    for n, v in dict(action=a, log_prob=lp, mean_t=mt, q1=q1, q2=q2).items():
        assert relerr(v, g["train_" + n]) < 2e-5, n
    loss = (lp.mean() * 0.3 - torch.min(q1, q2).mean()) + (q1 ** 2).mean() * 0.1
    loss.backward()
    assert abs(float(loss.detach()) - float(g["grad_loss"])) < 1e-5 * max(1.0, abs(float(g["grad_loss"])))
    an = np.array([0.0 if v.grad is None else float(v.grad.double().norm()) for v in pa_g.values()])
    cn = np.array([0.0 if v.grad is None else float(v.grad.double().norm()) for v in pc_g.values()])
    np.testing.assert_allclose(an, g["grad_actor_norms"], rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(cn, g["grad_critic_norms"], rtol=2e-4, atol=1e-9)


@pytest.mark.parametrize("tag,block,head,lfs,B", [("small", 2, 2, 32, 4), ("shipped", 4, 4, 64, 4)])
def test_learn_against_golden(tag, block, head, lfs, B):
    """SACOracle.learn vs the recorded run of the UNMODIFIED reference SAC.learn."""
    g = golden(f"learn_{tag}.npz")
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    actor, critic = reference_sac_init(cfg, SEED)
    orc = O.SACOracle(actor, critic, cfg)
    shp = (B, cfg.n_tokens, cfg.dim)
    n = int(np.prod(shp))
    steps = int(g["cfg"][4])
    for s in range(steps):
        batch = synthetic_batch(cfg, B, SEED + 10 + s)
        bits = np.unpackbits(g[f"step{s}_noise_bits"])
        noise = {k: torch.from_numpy(bits[i * ((n + 7) // 8 * 8):][:n].reshape(shp).astype(np.float32))
                 for i, k in enumerate(("mask_a_next", "mask_ct", "mask_c", "mask_a", "mask_c_pi"))}
        noise["eps_next"], noise["eps_pi"] = (torch.from_numpy(x) for x in g[f"step{s}_eps"])
        l = orc.learn(batch, noise)
        np.testing.assert_allclose(np.array(l), g[f"step{s}_losses"], rtol=1e-5, atol=1e-6)
        assert abs(float(orc.log_alpha) - float(g[f"step{s}_log_alpha"])) < 1e-7
        for nm, d in (("actor", orc.actor), ("critic", orc.critic), ("target", orc.critic_target)):
            a = np.array([float(v.double().abs().sum()) for v in d.values()])
            np.testing.assert_allclose(a, g[f"step{s}_{nm}_abssum"], rtol=2e-4, atol=1e-3)
        if s == 0:
            assert [k for k, v in orc.last_actor_grads.items() if v is None] == [str(x) for x in g["actor_unused"]]
            gn = np.array([0.0 if v is None else float(v.double().norm()) for v in orc.last_actor_grads.values()])
            np.testing.assert_allclose(gn, g["step0_actor_gnorm"], rtol=1e-4, atol=1e-9)


def test_depth_augment_against_golden():
    g = golden("depth_aug.npz")
    for i in range(2):
        H, W, k = (int(x) for x in g[f"raw_{i}_params"])
        yy, xx = np.mgrid[0:H, 0:W]
        raw = (0.03 + 7.97 * (0.5 + 0.5 * np.sin(xx / 37.0 + k) * np.cos(yy / 53.0))).astype(np.float32)
        raw[H // 3: H // 3 + 40, W // 4: W // 4 + 90] = 1.25
        np.random.seed(int(g[f"noise_seed_{i}"]))
        noise = np.random.normal(0, 50, raw.shape)
        got = O.depth_augment(raw, noise, out_hw=(H // 4, W // 4))
        assert np.abs(got - g[f"state_{i}"]).max() < 1e-9


def test_replay_gather_semantics():
    rng = np.random.RandomState(0)
    size = 17
    store = dict(obs=rng.rand(size, 8).astype(np.float32), act=rng.rand(size, 2).astype(np.float32))
    idx = np.array([0, 16, 5, 5])
    out = O.replay_gather(store, idx, size)
    assert np.array_equal(out["obs"], store["obs"][idx])
    assert np.array_equal(out["next_obs"], store["obs"][[1, 0, 6, 6]])


def _guidence_inputs(g, cfg, B, Be, s):
    """Rebuild the planted minibatches / replayed noise of oracle/make_golden.py:case_guidence."""
    ba, be = synthetic_batch(cfg, B, SEED + 30 + s), synthetic_batch(cfg, Be, SEED + 40 + s)
    Bc = B + Be
    sizes = dict(mask_a_next=Bc, mask_ct=Bc, mask_c=Bc, mask_a=Bc, mask_c_pi=Bc, mask_g=Be, mask_e=2)
    bits = np.unpackbits(g[f"step{s}_noise_bits"])
    noise, off = {}, 0
    for k, n in sizes.items():
        cnt = n * cfg.n_tokens * cfg.dim
        noise[k] = torch.from_numpy(bits[off:off + cnt].reshape(n, cfg.n_tokens, cfg.dim).astype(np.float32))
        off += (cnt + 7) // 8 * 8
    eps, off = g[f"step{s}_eps"], 0
    for k, n in dict(eps_next=Bc, eps_pi=Bc, eps_g=Be, eps_e=2).items():
        noise[k] = torch.from_numpy(eps[off:off + 2 * n].reshape(n, 2).copy())
        off += 2 * n
    cat = {k: torch.cat([ba[k], be[k]], 0) for k in ba}
    return ba, be, cat, noise


def test_learn_guidence_against_golden():
    """SACOracle.learn_guidence vs the recorded run of the UNMODIFIED reference SAC.learn_guidence
    (expert minibatch + two engaged rows)."""
    g = golden("guidence_small.npz")
    lfs, block, head, B, Be, steps = (int(x) for x in g["cfg"])
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    actor, critic = reference_sac_init(cfg, SEED)
    orc = O.SACOracle(actor, critic, cfg)
    for s in range(steps):
        ba, be, cat, noise = _guidence_inputs(g, cfg, B, Be, s)
        l = orc.learn_guidence(cat, noise, expert=dict(obs=be["obs"], pobs=be["pobs"], act=be["act"]),
                               engage_rows=torch.tensor([1, B - 1]))
        np.testing.assert_allclose(np.array(l), g[f"step{s}_losses"], rtol=1e-5, atol=1e-6)
        assert abs(float(orc.log_alpha) - float(g[f"step{s}_log_alpha"])) < 1e-7
        for nm, d in (("actor", orc.actor), ("critic", orc.critic)):
            a = np.array([float(v.double().abs().sum()) for v in d.values()])
            np.testing.assert_allclose(a, g[f"step{s}_{nm}_abssum"], rtol=2e-4, atol=1e-3)


def test_qnet_oracle_against_golden():
    """CNN twin-Q critic: the oracle restatement vs the outputs recorded from the unmodified reference ``QNetwork``."""
    from oracle.init_params import reference_qnet_init, synthetic_batch
    G = golden("qnet.npz")
    B = int(G["cfg"][0])
    p = reference_qnet_init(SEED + 7)
    assert [str(n) for n in G["names"]] == list(p.keys())
    assert np.allclose([float(v.double().sum()) for v in p.values()], G["sum"], rtol=0, atol=1e-9)
    batch = synthetic_batch(O.Cfg(), B, SEED + 8)
    pg = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    q1, q2 = O.qnet_forward(pg, batch["obs"], batch["pobs"], batch["act"])
    assert relerr(q1.detach(), torch.tensor(G["q1"])) < 1e-5 and relerr(q2.detach(), torch.tensor(G["q2"])) < 1e-5
    w = torch.linspace(0.5, 1.5, B).unsqueeze(1)
    loss = ((q1 - 0.3) ** 2 * w).mean() + (torch.min(q1, q2) * w).mean()
    loss.backward()
    norms = np.array([float(v.grad.double().norm()) for v in pg.values()])
    assert np.allclose(norms, G["grad_norms"], rtol=2e-4, atol=1e-9)
