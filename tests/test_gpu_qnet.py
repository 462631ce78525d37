"""GPU: CNN twin-Q critic ``QNetwork`` (vn/got_sac_network.py:125-170, the shipped default critic) through the C ABI
against the oracle restatement and the outputs recorded from the unmodified reference (tests/golden/qnet.npz)."""
import numpy as np
import pytest
import torch

import dgvit_b200 as dg
from helpers import O, SEED, golden, relerr
from oracle.init_params import reference_qnet_init, synthetic_batch

pytestmark = pytest.mark.gpu


def _module(params, precision):
    m = dg.QNetwork(2, 2)
    m.load_state_dict(params)
    m = m.to("cuda")
    m.precision = precision
    return m


def _loss(q1, q2, w):
    return ((q1 - 0.3) ** 2 * w).mean() + (torch.min(q1, q2) * w).mean()


@pytest.mark.parametrize("precision,tol_out,tol_grad", [("fp32", 1e-4, 2e-4), ("bf16", 1e-2, 4e-2)])
def test_qnet_matches_reference_recording(precision, tol_out, tol_grad):
    G = golden("qnet.npz")
    B = int(G["cfg"][0])
    params = reference_qnet_init(SEED + 7)
    assert list(params.keys()) == [str(n) for n in G["names"]]
    batch = synthetic_batch(O.Cfg(), B, SEED + 8)
    m = _module(params, precision)
    assert list(m.state_dict().keys()) == list(params.keys())
    img, goal, act = batch["obs"].cuda(), batch["pobs"].cuda(), batch["act"].cuda().requires_grad_(True)
    q1, q2 = m([img, goal, act])
    assert relerr(q1.detach().cpu(), torch.tensor(G["q1"])) < tol_out
    assert relerr(q2.detach().cpu(), torch.tensor(G["q2"])) < tol_out
    w = torch.linspace(0.5, 1.5, B).unsqueeze(1).cuda()
    loss = _loss(q1, q2, w)
    assert abs(float(loss) - float(G["loss"])) < tol_out * max(1.0, abs(float(G["loss"])))
    loss.backward()
    norms = np.array([float(p.grad.double().norm()) for p in m.parameters()])
    assert np.all(np.abs(norms - G["grad_norms"]) <= tol_grad * np.maximum(G["grad_norms"], 1e-6)), (norms, G["grad_norms"])
    gp = dict(m.named_parameters())
    assert relerr(gp["fc1.bias"].grad.cpu(), torch.tensor(G["grad_fc1_bias"])) < tol_grad
    assert relerr(gp["conv1.weight"].grad.cpu(), torch.tensor(G["grad_conv1_weight"])) < tol_grad
    assert act.grad is not None and torch.isfinite(act.grad).all()


@pytest.mark.parametrize("B,hw", [(1, (128, 160)), (5, (61, 77)), (37, (128, 160))])
def test_qnet_against_oracle_all_gradients(B, hw):
    """fp32 path, every parameter gradient and d/d(action), odd batch sizes and a ragged image size."""
    g = torch.Generator().manual_seed(B * 7 + hw[0])
    params = reference_qnet_init(B + 100)
    img = torch.rand(B, *hw, generator=g)
    goal = torch.rand(B, 2, generator=g) * 2 - 1
    act = torch.rand(B, 2, generator=g) * 2 - 1
    w = torch.linspace(0.5, 1.5, B).unsqueeze(1)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ar = act.clone().requires_grad_(True)
    oq1, oq2 = O.qnet_forward(pr, img, goal, ar)
    _loss(oq1, oq2, w).backward()
    m = dg.QNetwork(2, 2, image_size=hw)
    m.load_state_dict(params)
    m = m.to("cuda")
    ag = act.cuda().requires_grad_(True)
    q1, q2 = m([img.cuda(), goal.cuda(), ag])
    assert relerr(q1.detach().cpu(), oq1.detach()) < 1e-4 and relerr(q2.detach().cpu(), oq2.detach()) < 1e-4
    _loss(q1, q2, w.cuda()).backward()
    for k, p in m.named_parameters():
        assert relerr(p.grad.cpu(), pr[k].grad) < 2e-4, k
    assert relerr(ag.grad.cpu(), ar.grad) < 2e-4
    # policy-loss style call: only d/d(action) is wanted (critic parameters frozen)
    for p in m.parameters():
        p.requires_grad_(False)
    ag2 = act.cuda().requires_grad_(True)
    q1, q2 = m([img.cuda(), goal.cuda(), ag2])
    torch.min(q1, q2).mean().backward()
    pr2 = {k: v.clone() for k, v in params.items()}
    ar2 = act.clone().requires_grad_(True)
    torch.min(*O.qnet_forward(pr2, img, goal, ar2)).mean().backward()
    assert relerr(ag2.grad.cpu(), ar2.grad) < 2e-4


def test_qnet_module_protocol():
    import copy
    m = dg.QNetwork(2, 2).to("cuda")
    m.bind()
    t = copy.deepcopy(m)
    dg.hard_update(t, m)
    for (k, a), (_, b) in zip(m.named_parameters(), t.named_parameters()):
        assert torch.equal(a, b), k
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    img, goal, act = torch.rand(3, 128, 160).cuda(), torch.rand(3, 2).cuda(), torch.rand(3, 2).cuda()
    q1, q2 = m([img, goal, act])
    (q1.pow(2).mean() + q2.pow(2).mean()).backward()
    before = m.fc1.weight.detach().clone()
    opt.step()
    assert not torch.equal(before, m.fc1.weight) and m._bound()
    dg.soft_update(t, m, 0.5)
    assert torch.allclose(t.fc1.weight, 0.5 * before + 0.5 * m.fc1.weight)


def test_sac_learn_with_cnn_critic_matches_reference_style_update():
    """critic_type != "Transformer" (vn/DRL.py:118-121): two ``learn`` updates of the agent (QNetwork critic + DGViT actor,
    fp32) against the same statements (vn/DRL.py:388-434) run with autograd on the oracle's functions."""
    import torch.nn.functional as F
    from helpers import reference_init
    from oracle.init_params import synthetic_noise
    cfg = O.Cfg(dim=32, depth=2, heads=2)
    B = 5
    ag = dg.SAC(2, 2, "GaussianTransformer", "CNN", False, False, False, 11, LR_C=1e-3, LR_A=1e-3, LR_ALPHA=1e-4,
                BUFFER_SIZE=64, TAU=5e-3, POLICY_FREQ=1, GAMMA=0.99, ALPHA=0.2, block=2, head=2, l_f_size=32,
                precision="fp32")
    pa, pc = reference_init("actor", cfg, 21), reference_qnet_init(22)
    ag.policy.load_state_dict(pa); ag.critic.load_state_dict(pc); ag.critic_target.load_state_dict(pc)
    ag._after_load()
    ra = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    rc = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    rt = {k: v.clone() for k, v in pc.items()}
    log_alpha = torch.zeros(1, requires_grad=True)
    opt_c, opt_a = torch.optim.Adam(list(rc.values()), lr=1e-3), torch.optim.Adam(list(ra.values()), lr=1e-3)
    opt_al = torch.optim.Adam([log_alpha], lr=1e-4)
    alpha = 0.2
    for step in range(2):
        batch, nz = synthetic_batch(cfg, B, 300 + step), synthetic_noise(cfg, B, 400 + step)
        s, s2, ps, ps2, a, r = (batch[k] for k in ("obs", "next_obs", "pobs", "next_pobs", "act", "rew"))
        with torch.no_grad():
            a2, lp2, _ = O.actor_sample(ra, s2, ps2, nz["eps_next"], cfg, nz["mask_a_next"])
            q1t, q2t = O.qnet_forward(rt, s2, ps2, a2)
            nq = r + 0.99 * (torch.min(q1t, q2t) - alpha * lp2)
        q1, q2 = O.qnet_forward(rc, s, ps, a)
        l1 = F.mse_loss(q1, nq)
        lq = l1 + F.mse_loss(q2, nq)
        opt_c.zero_grad(); lq.backward(); opt_c.step()
        pi, lp, _ = O.actor_sample(ra, s, ps, nz["eps_pi"], cfg, nz["mask_a"])
        q1p, q2p = O.qnet_forward(rc, s, ps, pi)
        lpol = ((alpha * lp) - torch.min(q1p, q2p)).mean()
        opt_a.zero_grad(); lpol.backward()
        opt_a.step()
        lal = -(log_alpha * (lp + (-2.0)).detach()).mean()
        opt_al.zero_grad(); lal.backward(); opt_al.step()
        alpha = float(log_alpha.exp())
        with torch.no_grad():
            for k in rt:
                rt[k].mul_(1 - 5e-3).add_(rc[k].detach(), alpha=5e-3)
        gb = {k: v.cuda() for k, v in batch.items()}
        qg, pg = ag._learn_cnn(gb, {k: (v.cuda() if v is not None else None) for k, v in nz.items()})
        assert abs(float(qg) - float(l1)) < 1e-4 * max(1.0, abs(float(l1))), (step, float(qg), float(l1))
        assert abs(float(pg) - float(lpol)) < 1e-4 * max(1.0, abs(float(lpol))), (step, float(pg), float(lpol))
        assert abs(float(ag.alpha) - alpha) < 1e-6
        for k, p in ag.critic.named_parameters():
            assert float((p.detach().cpu() - rc[k].detach()).abs().max()) < 3e-5 * (step + 1), ("critic", step, k)
        for k, p in ag.critic_target.named_parameters():
            assert float((p.detach().cpu() - rt[k]).abs().max()) < 1e-5, ("target", step, k)
        for k, p in ag.policy.named_parameters():
            d = (p.detach().cpu() - ra[k].detach()).abs()
            assert float((d > 3e-5 * (step + 1)).float().mean()) < 5e-3, ("actor", step, k)
    # the public entry point: replay store -> gather -> update
    ag.replay_buffer.fill_synthetic(32)
    q, p = ag.learn(4)
    assert np.isfinite(q) and np.isfinite(p)


@pytest.mark.gpu
def test_sac_learn_guidence_with_cnn_critic():
    """The reference's shipped configuration (vn/config.yaml:61 CNN critic + PRE_BUFFER): ``learn_guidence``
    (vn/DRL.py:237-299) with the QNetwork critic -- the statements run with autograd on the oracle's functions, expert
    rows and engaged rows through two separate ``policy.sample`` calls as in the reference; the drop-in runs them as one
    pass over [expert rows | engaged rows] with per-row weights."""
    import torch.nn.functional as F
    from helpers import reference_init
    from oracle.init_params import synthetic_noise
    cfg = O.Cfg(dim=32, depth=2, heads=2)
    B, Be = 4, 2                                   # agent rows, expert rows
    ag = dg.SAC(2, 2, "GaussianTransformer", "CNN", False, False, True, 11, LR_C=1e-3, LR_A=1e-3, LR_ALPHA=1e-4,
                BUFFER_SIZE=64, TAU=5e-3, POLICY_FREQ=1, GAMMA=0.99, ALPHA=0.2, block=2, head=2, l_f_size=32,
                precision="fp32", buffer_size_expert=16)
    pa, pc = reference_init("actor", cfg, 31), reference_qnet_init(32)
    ag.policy.load_state_dict(pa); ag.critic.load_state_dict(pc); ag.critic_target.load_state_dict(pc)
    ag._after_load()
    ra = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    rc = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    rt = {k: v.clone() for k, v in pc.items()}
    opt_c, opt_a = torch.optim.Adam(list(rc.values()), lr=1e-3), torch.optim.Adam(list(ra.values()), lr=1e-3)
    alpha, gw, ew = 0.2, ag.guidence_weight, ag.engage_weight
    Bc = B + Be
    batch, nz = synthetic_batch(cfg, Bc, 500), synthetic_noise(cfg, Bc, 600)
    eng = torch.tensor([1, 3])                     # engaged agent rows
    g = torch.Generator().manual_seed(7)
    mask_g = (torch.rand(Be, cfg.n_tokens, cfg.dim, generator=g) >= 0.1).float()
    mask_e = (torch.rand(len(eng), cfg.n_tokens, cfg.dim, generator=g) >= 0.1).float()
    eps0 = torch.zeros(Be, 2), torch.zeros(len(eng), 2)
    s, s2, ps, ps2, a, r = (batch[k] for k in ("obs", "next_obs", "pobs", "next_pobs", "act", "rew"))
    with torch.no_grad():
        a2, lp2, _ = O.actor_sample(ra, s2, ps2, nz["eps_next"], cfg, nz["mask_a_next"])
        q1t, q2t = O.qnet_forward(rt, s2, ps2, a2)
        nq = r + 0.99 * (torch.min(q1t, q2t) - alpha * lp2)
    q1, q2 = O.qnet_forward(rc, s, ps, a)
    l1 = F.mse_loss(q1, nq)
    lq = l1 + F.mse_loss(q2, nq)
    opt_c.zero_grad(); lq.backward(); opt_c.step()
    pi, lp, _ = O.actor_sample(ra, s, ps, nz["eps_pi"], cfg, nz["mask_a"])
    q1p, q2p = O.qnet_forward(rc, s, ps, pi)
    _, _, pred_g = O.actor_sample(ra, s[B:Bc], ps[B:Bc], eps0[0], cfg, mask_g)                   # :259-263
    _, _, pred_e = O.actor_sample(ra, s[eng], ps[eng], eps0[1], cfg, mask_e)                     # :267-273
    lpol = ((alpha * lp) - torch.min(q1p, q2p)).mean() + gw * F.mse_loss(pred_g, a[B:Bc]).mean() \
        + ew * F.mse_loss(pred_e, a[eng]).mean()
    opt_a.zero_grad(); lpol.backward(); opt_a.step()

    gb = {k: v.cuda() for k, v in batch.items()}
    n_g, n_e = Be * 2, len(eng) * 2
    extra = dict(obs=torch.cat([s[B:Bc], s[eng]]).cuda(), pobs=torch.cat([ps[B:Bc], ps[eng]]).cuda(),
                 target=torch.cat([a[B:Bc], a[eng]]).cuda(),
                 weight=torch.cat([torch.full((Be,), gw / n_g), torch.full((len(eng),), ew / n_e)]).cuda())
    nzg = {k: (v.cuda() if v is not None else None) for k, v in nz.items()}
    nzg["mask_x"] = torch.cat([mask_g, mask_e]).cuda()
    nzg["eps_x"] = torch.zeros(Be + len(eng), 2).cuda()
    qg, pg = ag._learn_cnn(gb, nzg, extra=extra)
    assert abs(float(qg) - float(l1)) < 1e-4 * max(1.0, abs(float(l1))), (float(qg), float(l1))
    assert abs(float(pg) - float(lpol)) < 1e-4 * max(1.0, abs(float(lpol))), (float(pg), float(lpol))
    for k, p in ag.critic.named_parameters():
        assert float((p.detach().cpu() - rc[k].detach()).abs().max()) < 3e-5, ("critic", k)
    for k, p in ag.policy.named_parameters():
        d = (p.detach().cpu() - ra[k].detach()).abs()
        assert float((d > 3e-5).float().mean()) < 5e-3, ("actor", k)

    # the public entry point: agent + expert replay stores, engaged transitions
    rs = np.random.RandomState(0)
    for i in range(10):
        f1, f2 = rs.rand(128, 160).astype(np.float32), rs.rand(128, 160).astype(np.float32)
        ag.store_transition(f1, rs.rand(2) * 2 - 1, rs.rand(2), rs.rand(2), float(rs.randn()), f2, float(i % 3 == 0), None, 0)
        ag.initialize_expert_buffer(f1, rs.rand(2) * 2 - 1, rs.rand(2), rs.rand(2), 1.0, f2, 0)
    for _ in range(2):
        q, p = ag.learn_guidence(False, 6)
        assert np.isfinite(q) and np.isfinite(p)


def _cnn_agent(seed, precision, **kw):
    return dg.SAC(2, 2, "GaussianTransformer", "CNN", False, False, False, seed, LR_C=1e-3, LR_A=1e-3, LR_ALPHA=1e-4,
                  BUFFER_SIZE=256, TAU=5e-3, POLICY_FREQ=1, GAMMA=0.99, ALPHA=0.2, block=4, head=4, l_f_size=64,
                  precision=precision, **kw)


@pytest.mark.gpu
def test_cnn_critic_update_bf16_tracks_fp32_at_the_shipped_preset():
    """The single-call update with the CNN critic at the shipped preset (D=64, 4 blocks, 4 heads), B=64: the bf16 path
    (tcgen05 convolution GEMMs, fused MLP) against the fp32 path of the same library on the same batch and injected noise
    (the fp32 path is the one checked against the oracle above): TD target, Q, actions, losses, critic gradients."""
    from oracle.init_params import synthetic_noise
    cfg = O.Cfg()
    B = 64
    batch, nz = synthetic_batch(cfg, B, 700), synthetic_noise(cfg, B, 701)
    gb = {k: v.reshape(B, -1).cuda().contiguous() for k, v in batch.items()}
    noise = {k: nz[k].cuda() for k in ("eps_next", "eps_pi", "mask_a_next", "mask_a")}
    noise = {k: (v.to(torch.uint8) if k.startswith("mask") else v).contiguous() for k, v in noise.items()}
    noise["mask_c"] = noise["mask_a"]
    res = {}
    for prec in ("fp32", "bf16"):
        ag = _cnn_agent(5, prec)
        dbg = torch.zeros(B * 11, device="cuda")
        losses = ag.update_from_batch(gb, noise, debug=dbg).tolist()
        res[prec] = (losses, dbg.cpu(), ag.critic._garena.detach().cpu().clone(), ag.policy._garena.detach().cpu().clone())
    (lf, df, gcf, gaf), (lb, db, gcb, gab) = res["fp32"], res["bf16"]
    n2 = B * 2
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))
    assert rel(db[:n2], df[:n2]) < 2e-2            # TD target
    assert rel(db[n2:2 * n2], df[n2:2 * n2]) < 2e-2    # Q1(s, a)
    assert rel(db[3 * n2:4 * n2], df[3 * n2:4 * n2]) < 2e-2    # pi
    assert abs(lb[0] - lf[0]) <= 3e-2 * abs(lf[0]) and abs(lb[1] - lf[1]) <= 3e-2 * max(1.0, abs(lf[1])), (lb, lf)
    cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))
    assert cos(gcb, gcf) > 0.99 and 0.95 < float(gcb.norm() / gcf.norm()) < 1.05
    assert cos(gab, gaf) > 0.97


@pytest.mark.gpu
def test_cnn_critic_graph_replay_equals_eager():
    """``learn_async`` with the CNN critic: eager, captured and replayed updates of one agent follow the same trajectory as an
    agent that never uses a graph (same seed: same sampled indexes, same in-kernel Philox streams)."""
    out = {}
    for graph in (False, True):
        ag = _cnn_agent(9, "fp32", use_cuda_graph=graph)
        ag.replay_buffer.fill_synthetic(200, seed=3)
        ls = [ag.learn_async(8).clone() for _ in range(5)]
        torch.cuda.synchronize()
        out[graph] = (torch.stack(ls).cpu(), ag.critic._arena.detach().cpu().clone(), ag.policy._arena.detach().cpu().clone(),
                      ag.critic_target._arena.detach().cpu().clone())
        ag.close()
    for a, b in zip(out[False], out[True]):
        assert torch.isfinite(a).all()
        assert float((a - b).abs().max()) <= 1e-5 * max(1.0, float(a.abs().max()))
