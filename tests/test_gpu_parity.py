"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs, and against the committed golden vectors recorded from the live reference.

Tolerances (BASELINE.json north_star): fp32 path <= 1e-4 relative; bf16 path <= 1e-2 relative on
actions / Q; replay gather bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

import dgvit_b200 as dg
from dgvit_b200 import _lib as L
from helpers import (O, SEED, golden, load_params, reference_init, reference_sac_init, relerr, synthetic_batch,
                     synthetic_noise, unpack_mask)

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4
BF16_TOL = 1e-2


def _mk(kind, cfg, params, precision="fp32"):
    cls = dg.GoTPolicy if kind == "actor" else dg.GoTQNetwork
    m = cls(2, 2, cfg.depth, cfg.heads, cfg.dim, image_size=(cfg.img_h, cfg.img_w))
    load_params(m, params)
    m = m.to("cuda")
    m.precision = precision
    return m


@pytest.mark.parametrize("tag,block,head,lfs,B", [("small", 2, 2, 32, 3), ("shipped", 4, 4, 64, 4)])
def test_forward_matches_golden_reference_outputs(tag, block, head, lfs, B):
    """Outputs recorded from the imported reference modules (eval mode and train mode with the
    replayed dropout mask / rsample noise)."""
    g = golden(f"modules_{tag}.npz")
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    a = _mk("actor", cfg, reference_init("actor", cfg, SEED))
    c = _mk("critic", cfg, reference_init("critic", cfg, SEED + 1))
    batch = synthetic_batch(cfg, B, SEED + 2)
    img, goal, act = (batch[k].cuda() for k in ("obs", "pobs", "act"))
    a.eval(); c.eval()
    with torch.no_grad():
        mean, log_std = a([img, goal])
        q1, q2 = c([img, goal, act])
    for n, v in dict(mean=mean, log_std=log_std, q1=q1, q2=q2).items():
        assert relerr(v, g["eval_" + n]) < FP32_TOL, n
    a.train(); c.train()
    shp = (B, cfg.n_tokens, cfg.dim)
    a.inject_noise(mask=unpack_mask(g["train_mask_a"], shp), eps=torch.from_numpy(g["train_eps"]))
    c.inject_noise(mask=unpack_mask(g["train_mask_c"], shp))
    with torch.no_grad():
        action, logp, mean_t = a.sample([img, goal])
        q1, q2 = c([img, goal, action])
    for n, v in dict(action=action, log_prob=logp, mean_t=mean_t, q1=q1, q2=q2).items():
        assert relerr(v, g["train_" + n]) < FP32_TOL, n


@pytest.mark.parametrize("block,head,lfs,B", [(2, 2, 32, 5), (4, 4, 64, 3), (1, 3, 96, 2)])
def test_gradients_match_oracle_fp32(block, head, lfs, B):
    """autograd through the nn.Module surface == oracle autograd, per parameter."""
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    pa, pc = reference_init("actor", cfg, 11), reference_init("critic", cfg, 12)
    a, c = _mk("actor", cfg, pa), _mk("critic", cfg, pc)
    batch = synthetic_batch(cfg, B, 13)
    nz = synthetic_noise(cfg, B, 14)
    img, goal = batch["obs"], batch["pobs"]

    def loss_fn(lp, q1, q2, mt):
        return (lp.mean() * 0.3 - torch.min(q1, q2).mean()) + (q1 ** 2).mean() * 0.1 + (mt ** 2).sum() * 0.01

    pa_g = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    pc_g = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    oa, olp, omt = O.actor_sample(pa_g, img, goal, nz["eps_pi"], cfg, nz["mask_a"])
    oq1, oq2 = O.critic_forward(pc_g, img, goal, oa, cfg, nz["mask_c"])
    loss_fn(olp, oq1, oq2, omt).backward()

    a.inject_noise(mask=nz["mask_a"], eps=nz["eps_pi"])
    c.inject_noise(mask=nz["mask_c"])
    act, lp, mt = a.sample([img.cuda(), goal.cuda()])
    q1, q2 = c([img.cuda(), goal.cuda(), act])
    for n, (x, y) in dict(action=(act, oa), logp=(lp, olp), q1=(q1, oq1), q2=(q2, oq2)).items():
        assert relerr(x, y) < FP32_TOL, n
    loss_fn(lp, q1, q2, mt).backward()
    for mod, og in ((a, pa_g), (c, pc_g)):
        for k, p in mod.named_parameters():
            if og[k].grad is None:
                assert p.grad is None, k            # Adam must skip exactly these (cls_token, mlp_head, conv)
                continue
            assert p.grad is not None, k
            assert relerr(p.grad, og[k].grad) < 2e-4, (k, relerr(p.grad, og[k].grad))


def _agent(cfg, precision, seed=SEED):
    ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, seed, LR_C=1e-3, LR_A=1e-3,
                LR_ALPHA=1e-4, BUFFER_SIZE=64, TAU=5e-4, POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0, block=cfg.depth,
                head=cfg.heads, l_f_size=cfg.dim, automatic_entropy_tuning=True, precision=precision)
    return ag


def _noise_cuda(noise):
    out = {}
    for k, v in noise.items():
        if v is None:
            continue
        out[k] = v.cuda().to(torch.uint8).contiguous() if k.startswith("mask") else v.cuda().contiguous()
    return out


@pytest.mark.parametrize("tag,block,head,lfs,B", [("small", 2, 2, 32, 4), ("shipped", 4, 4, 64, 4)])
def test_learn_matches_golden_reference_run(tag, block, head, lfs, B):
    """3 fused updates vs the recorded run of the UNMODIFIED reference SAC.learn (same seeds,
    replayed noise) and vs the oracle stepping beside it."""
    g = golden(f"learn_{tag}.npz")
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    ag = _agent(cfg, "fp32")
    actor0, critic0 = reference_sac_init(cfg, SEED)
    for k, v in ag.policy.state_dict().items():      # same seed + same construction order => same weights
        assert torch.equal(v.cpu(), actor0[k]), k
    for k, v in ag.critic.state_dict().items():
        assert torch.equal(v.cpu(), critic0[k]), k
    orc = O.SACOracle(actor0, critic0, cfg)
    shp = (B, cfg.n_tokens, cfg.dim)
    n = int(np.prod(shp))
    for s in range(int(g["cfg"][4])):
        batch = synthetic_batch(cfg, B, SEED + 10 + s)
        bits = np.unpackbits(g[f"step{s}_noise_bits"])
        noise = {k: torch.from_numpy(bits[i * n:][:n].reshape(shp).astype(np.float32))
                 for i, k in enumerate(("mask_a_next", "mask_ct", "mask_c", "mask_a", "mask_c_pi"))}
        noise["eps_next"], noise["eps_pi"] = (torch.from_numpy(x) for x in g[f"step{s}_eps"])
        orc.learn(batch, noise)
        cb = {k: v.reshape(B, -1).cuda().contiguous() for k, v in batch.items()}
        losses = ag.update_from_batch(cb, _noise_cuda(noise)).cpu().numpy()
        ref = g[f"step{s}_losses"]
        assert abs(losses[0] - ref[0]) <= FP32_TOL * max(1.0, abs(ref[0])), (s, losses, ref)
        assert abs(losses[1] - ref[1]) <= FP32_TOL * max(1.0, abs(ref[1])), (s, losses, ref)
        assert abs(float(ag.log_alpha) - float(g[f"step{s}_log_alpha"])) < 1e-6
        for nm, mod, od in (("actor", ag.policy, orc.actor), ("critic", ag.critic, orc.critic),
                            ("target", ag.critic_target, orc.critic_target)):
            worst = 0.0
            for k, p in mod.named_parameters():
                d = (p.detach().cpu() - od[k]).abs()
                worst = max(worst, float(d.max()))
                # Adam's g/(|g|+eps) amplifies rounding where |g| ~ 1e-8: demand the bulk to be tight
                assert float((d > 2e-5 * (s + 1)).float().mean()) < 2e-3, (s, nm, k)
            assert worst < 2.5e-3 * (s + 1), (s, nm, worst)
            ab = np.array([float(p.detach().double().abs().sum()) for p in mod.parameters()])
            np.testing.assert_allclose(ab, g[f"step{s}_{nm}_abssum"], rtol=2e-4, atol=2e-3)


def test_learn_gradients_match_oracle_fp32():
    """Per-parameter gradient parity of one fused update at the shipped preset."""
    cfg = O.Cfg()
    B = 6
    ag = _agent(cfg, "fp32", seed=77)
    actor0 = {k: v.detach().cpu().clone() for k, v in ag.policy.named_parameters()}
    critic0 = {k: v.detach().cpu().clone() for k, v in ag.critic.named_parameters()}
    orc = O.SACOracle(actor0, critic0, cfg)
    batch = synthetic_batch(cfg, B, 5)
    noise = synthetic_noise(cfg, B, 6)
    orc.learn(batch, noise)
    cb = {k: v.reshape(B, -1).cuda().contiguous() for k, v in batch.items()}
    dbg = torch.zeros(B * (5 * 2 + 1), device="cuda")
    ag.update_from_batch(cb, _noise_cuda(noise), debug=dbg)
    d = dbg.cpu()
    for i, k in enumerate(("nq", "q1", "q2", "pi", "q1p")):
        assert relerr(d[i * B * 2:(i + 1) * B * 2].reshape(B, 2), orc.last[k]) < FP32_TOL, k
    assert relerr(d[5 * B * 2:5 * B * 2 + B].reshape(B, 1), orc.last["log_pi"]) < FP32_TOL
    for mod, og in ((ag.critic, orc.last_critic_grads), (ag.policy, orc.last_actor_grads)):
        for (k, off), p in zip(mod._named_offsets(), mod.parameters()):
            gr = mod._garena[off:off + p.numel()].view(p.shape).cpu()
            if og[k] is None:
                assert float(gr.abs().max()) == 0.0, k
                continue
            assert relerr(gr, og[k]) < 2e-4, (k, relerr(gr, og[k]))


def test_replay_gather_bit_exact():
    torch.manual_seed(0)
    st = dg.ReplayStore(50, (128, 160), 2, 2, "cuda", seed=1)
    st.fill_synthetic(50, seed=9)
    rng = np.random.RandomState(3)
    for B in (1, 7, 64):
        idx = torch.from_numpy(rng.randint(0, 50, size=B)).cuda()
        idx[0] = 49                                        # next_obs wraps into the extra slot
        out = {k: torch.full((B, w), -1.0, device="cuda") for k, w in
               dict(obs=20480, next_obs=20480, pobs=2, next_pobs=2, act=2, rew=1, done=1).items()}
        st.gather(idx, out)
        store = {k: getattr(st, k).cpu().numpy() for k in ("obs", "pobs", "next_pobs", "act", "rew", "done")}
        ref = O.replay_gather(store, idx.cpu().numpy(), st.cap)
        for k in out:
            assert np.array_equal(out[k].cpu().numpy().view(np.uint32), ref[k].view(np.uint32)), k


def test_depth_augment_matches_golden_and_oracle():
    g = golden("depth_aug.npz")
    for i in range(2):
        H, W, k = (int(x) for x in g[f"raw_{i}_params"])
        yy, xx = np.mgrid[0:H, 0:W]
        raw = (0.03 + 7.97 * (0.5 + 0.5 * np.sin(xx / 37.0 + k) * np.cos(yy / 53.0))).astype(np.float32)
        raw[H // 3: H // 3 + 40, W // 4: W // 4 + 90] = 1.25
        np.random.seed(int(g[f"noise_seed_{i}"]))
        noise = np.random.normal(0, 50, raw.shape).astype(np.float32)
        got = dg.depth_augment(torch.from_numpy(raw).cuda(), torch.from_numpy(noise).cuda())[0].cpu().numpy()
        want = O.depth_augment(raw, noise.astype(np.float64), out_hw=(H // 4, W // 4))
        assert np.abs(got - want).max() < 1e-5
        # the golden frame used float64 noise; float32 rounding of N(0,50) moves pixels by < 1e-5*255
        assert np.abs(got - g[f"state_{i}"]).max() < 1e-4


def test_soft_and_hard_update():
    cfg = O.Cfg(dim=32, depth=1, heads=2)
    a = _mk("critic", cfg, reference_init("critic", cfg, 1))
    b = _mk("critic", cfg, reference_init("critic", cfg, 2))
    pa = {k: v.detach().cpu().clone() for k, v in a.named_parameters()}
    pb = {k: v.detach().cpu().clone() for k, v in b.named_parameters()}
    dg.soft_update(a, b, 5e-4)
    O.soft_update(pa, pb, pa.keys(), 5e-4)
    for k, p in a.named_parameters():
        assert relerr(p, pa[k]) < 1e-6, k
    dg.hard_update(a, b)
    for (k, p), q in zip(a.named_parameters(), b.parameters()):
        assert torch.equal(p, q), k


def test_bf16_path_within_tolerance():
    """bf16 contractions (tensor-core path): actions / Q within 1e-2 relative of the fp32 oracle."""
    cfg = O.Cfg()
    B = 8
    pa, pc = reference_init("actor", cfg, 21), reference_init("critic", cfg, 22)
    a, c = _mk("actor", cfg, pa, "bf16"), _mk("critic", cfg, pc, "bf16")
    batch = synthetic_batch(cfg, B, 23)
    nz = synthetic_noise(cfg, B, 24)
    img, goal = batch["obs"], batch["pobs"]
    with torch.no_grad():
        oa, olp, omt = O.actor_sample(pa, img, goal, nz["eps_pi"], cfg, nz["mask_a"])
        oq1, oq2 = O.critic_forward(pc, img, goal, oa, cfg, nz["mask_c"])
        a.inject_noise(mask=nz["mask_a"], eps=nz["eps_pi"])
        c.inject_noise(mask=nz["mask_c"])
        act, lp, mt = a.sample([img.cuda(), goal.cuda()])
        q1, q2 = c([img.cuda(), goal.cuda(), oa.cuda()])
    assert relerr(act, oa) < BF16_TOL and relerr(mt, omt) < BF16_TOL
    assert relerr(q1, oq1) < BF16_TOL and relerr(q2, oq2) < BF16_TOL


def test_bf16_update_gradients_close_to_fp32_oracle():
    """bf16 path (tcgen05 GEMMs + tcgen05 attention): every parameter gradient of one fused update
    points the same way as the fp32 oracle's (cosine > 0.99, norm within 5 %)."""
    cfg = O.Cfg()
    B = 16
    ag = _agent(cfg, "bf16", seed=5)
    actor0 = {k: v.detach().cpu().clone() for k, v in ag.policy.named_parameters()}
    critic0 = {k: v.detach().cpu().clone() for k, v in ag.critic.named_parameters()}
    orc = O.SACOracle(actor0, critic0, cfg)
    batch, noise = synthetic_batch(cfg, B, 31), synthetic_noise(cfg, B, 32)
    want = orc.learn(batch, noise)
    cb = {k: v.reshape(B, -1).cuda().contiguous() for k, v in batch.items()}
    dbg = torch.zeros(B * 11, device="cuda")
    got = ag.update_from_batch(cb, _noise_cuda(noise), debug=dbg).tolist()
    assert abs(got[0] - want[0]) <= BF16_TOL * abs(want[0])
    # the policy loss mixes alpha*log_pi with -min Q; the stated bf16 bound is on actions / Q (below)
    assert abs(got[1] - want[1]) <= 5e-2 * max(1.0, abs(want[1]))
    d = dbg.cpu()
    assert relerr(d[B * 2:2 * B * 2].reshape(B, 2), orc.last["q1"]) < BF16_TOL      # Q
    assert relerr(d[3 * B * 2:4 * B * 2].reshape(B, 2), orc.last["pi"]) < BF16_TOL  # actions
    for mod, og in ((ag.critic, orc.last_critic_grads), (ag.policy, orc.last_actor_grads)):
        for (k, off), p in zip(mod._named_offsets(), mod.parameters()):
            if og[k] is None:
                continue
            gr = mod._garena[off:off + p.numel()].view(p.shape).cpu().double().flatten()
            rf = og[k].double().flatten()
            if float(rf.norm()) < 1e-12:
                continue
            cos = float((gr @ rf) / (gr.norm() * rf.norm() + 1e-300))
            ratio = float(gr.norm() / rf.norm())
            assert cos > 0.99 and 0.95 < ratio < 1.05, (k, cos, ratio)


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_choose_action_batch1(precision, tol):
    """Batch-1 act (vn/DRL.py:170-185): numpy frame (H,W,1) + goal (2,) -> numpy action, graph-replayed."""
    cfg = O.Cfg()
    pa = reference_init("actor", cfg, 41)
    a = _mk("actor", cfg, pa, precision)
    a.eval()                                   # deterministic: no dropout, evaluate=True -> tanh(mean)
    rs = np.random.RandomState(1)
    for i in range(3):                         # repeated calls replay the captured graph with new inputs
        frame = rs.rand(128, 160, 1).astype(np.float32)
        goal = rs.rand(2).astype(np.float32)
        got = a.choose_action(frame, goal, evaluate=True)
        with torch.no_grad():
            _, _, want = O.actor_sample(pa, torch.from_numpy(frame).permute(2, 0, 1), torch.from_numpy(goal)[None],
                                        torch.zeros(1, 2), cfg, None)
        assert got.shape == (2,) and got.dtype == np.float32
        assert relerr(got, want[0]) < tol, (i, got, want)
    a.train()                                  # stochastic path: bounded actions, changes call to call
    x = a.choose_action(frame, goal)
    y = a.choose_action(frame, goal)
    assert np.all(np.abs(x) <= 1) and not np.array_equal(x, y)


def test_depth_augment_in_kernel_noise_statistics():
    """Generated noise (Philox + Box-Muller, sigma 50 on the 0..255 scale): same statistics as the
    injected-noise path, deterministic for a given rng state, different across frames / counters."""
    H, W, n = 512, 640, 4
    raw = (torch.rand(n, H, W, generator=torch.Generator().manual_seed(3)) * 6 + 1).cuda()
    rng = torch.tensor([1234, 0], dtype=torch.int64, device="cuda")
    a = dg.depth_augment(raw, None, rng)
    b = dg.depth_augment(raw, None, rng)
    assert torch.equal(a, b)
    rng2 = torch.tensor([1234, 1], dtype=torch.int64, device="cuda")
    c = dg.depth_augment(raw, None, rng2)
    assert not torch.equal(a, c) and not torch.equal(a[0], a[1])
    noise = torch.randn(n, H, W, generator=torch.Generator().manual_seed(4)).cuda() * 50
    ref = dg.depth_augment(raw, noise)
    assert a.shape == ref.shape == (n, H // 4, W // 4)
    assert float(a.min()) >= 0 and float(a.max()) <= 1
    assert abs(float(a.mean()) - float(ref.mean())) < 5e-3
    assert abs(float(a.std()) - float(ref.std())) < 5e-3


def test_wide_deep_variant_at_2x_resolution():
    """BASELINE config 5 shape family: 256x320 frames (257 tokens), D=128, 6 heads, fp32
    forward + every parameter gradient vs the oracle, bf16 forward within tolerance."""
    cfg = O.Cfg(dim=128, depth=2, heads=6, img_h=256, img_w=320)
    B = 2
    pa, pc = reference_init("actor", cfg, 51), reference_init("critic", cfg, 52)
    batch, nz = synthetic_batch(cfg, B, 53), synthetic_noise(cfg, B, 54)
    img, goal = batch["obs"], batch["pobs"]
    pa_g = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    pc_g = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    oa, olp, omt = O.actor_sample(pa_g, img, goal, nz["eps_pi"], cfg, nz["mask_a"])
    oq1, oq2 = O.critic_forward(pc_g, img, goal, oa, cfg, nz["mask_c"])
    (olp.mean() * 0.3 - torch.min(oq1, oq2).mean() + (oq1 ** 2).mean() * 0.1).backward()
    a, c = _mk("actor", cfg, pa), _mk("critic", cfg, pc)
    a.inject_noise(mask=nz["mask_a"], eps=nz["eps_pi"])
    c.inject_noise(mask=nz["mask_c"])
    act, lp, mt = a.sample([img.cuda(), goal.cuda()])
    q1, q2 = c([img.cuda(), goal.cuda(), act])
    for n, (x, y) in dict(action=(act, oa), logp=(lp, olp), q1=(q1, oq1), q2=(q2, oq2)).items():
        assert relerr(x, y) < FP32_TOL, n
    (lp.mean() * 0.3 - torch.min(q1, q2).mean() + (q1 ** 2).mean() * 0.1).backward()
    for mod, og in ((a, pa_g), (c, pc_g)):
        for k, p in mod.named_parameters():
            if og[k].grad is None:
                assert p.grad is None, k
                continue
            assert relerr(p.grad, og[k].grad) < 3e-4, (k, relerr(p.grad, og[k].grad))
    ab, cb = _mk("actor", cfg, pa, "bf16"), _mk("critic", cfg, pc, "bf16")
    with torch.no_grad():
        ab.inject_noise(mask=nz["mask_a"], eps=nz["eps_pi"])
        cb.inject_noise(mask=nz["mask_c"])
        act_b, _, _ = ab.sample([img.cuda(), goal.cuda()])
        q1_b, _ = cb([img.cuda(), goal.cuda(), oa.detach().cuda()])
    assert relerr(act_b, oa) < BF16_TOL and relerr(q1_b, oq1) < BF16_TOL


@pytest.mark.parametrize("B,precision", [(1, "fp32"), (3, "fp32"), (130, "bf16"), (1, "bf16")])
def test_update_ragged_batches(B, precision):
    """Batch sizes that do not fill a 128-row tile / a warp: one fused update vs the oracle."""
    cfg = O.Cfg(dim=64, depth=2, heads=4)
    ag = _agent(cfg, precision, seed=9)
    actor0 = {k: v.detach().cpu().clone() for k, v in ag.policy.named_parameters()}
    critic0 = {k: v.detach().cpu().clone() for k, v in ag.critic.named_parameters()}
    orc = O.SACOracle(actor0, critic0, cfg)
    batch, noise = synthetic_batch(cfg, B, 61), synthetic_noise(cfg, B, 62)
    want = orc.learn(batch, noise)
    cb = {k: v.reshape(B, -1).cuda().contiguous() for k, v in batch.items()}
    got = ag.update_from_batch(cb, _noise_cuda(noise)).tolist()
    tol = FP32_TOL if precision == "fp32" else 5e-2
    assert abs(got[0] - want[0]) <= tol * max(1.0, abs(want[0])), (got, want)
    assert abs(got[1] - want[1]) <= tol * max(1.0, abs(want[1])), (got, want)
    for mod, od in ((ag.policy, orc.actor), (ag.critic, orc.critic)):
        for k, p in mod.named_parameters():
            d = (p.detach().cpu() - od[k]).abs()
            assert torch.isfinite(p).all(), k
            if precision == "fp32":
                assert float((d > 2e-5).float().mean()) < 5e-3, k


def test_learn_guidence_matches_golden_reference_run():
    """SAC.learn_guidence semantics (vn/DRL.py:187-301): agent + expert minibatch, guidance rows and engaged
    rows as extra actor rows of ONE fused update, vs the recorded unmodified-reference run and the oracle."""
    from test_oracle_golden import _guidence_inputs
    g = golden("guidence_small.npz")
    lfs, block, head, B, Be, steps = (int(x) for x in g["cfg"])
    cfg = O.Cfg(dim=lfs, depth=block, heads=head)
    ag = _agent(cfg, "fp32")
    actor0, critic0 = reference_sac_init(cfg, SEED)
    orc = O.SACOracle(actor0, critic0, cfg)
    for s in range(steps):
        ba, be, cat, noise = _guidence_inputs(g, cfg, B, Be, s)
        eng = torch.tensor([1, B - 1])
        orc.learn_guidence(cat, noise, expert=dict(obs=be["obs"], pobs=be["pobs"], act=be["act"]), engage_rows=eng)
        Bc, ne = B + Be, 2
        flat = lambda t: t.reshape(t.shape[0], -1)
        batch = dict(obs=torch.cat([flat(cat["obs"]), flat(be["obs"]), flat(cat["obs"][eng])]),
                     pobs=torch.cat([cat["pobs"], be["pobs"], cat["pobs"][eng]]),
                     next_obs=flat(cat["next_obs"]), next_pobs=cat["next_pobs"], act=cat["act"], rew=cat["rew"], done=cat["done"])
        extra = dict(target=torch.cat([be["act"], cat["act"][eng]]),
                     weight=torch.cat([torch.full((Be,), 1.0 / (Be * 2)), torch.full((ne,), 1.0 / (ne * 2))]))
        nz = dict(noise)
        nz["mask_a"] = torch.cat([noise["mask_a"], noise["mask_g"], noise["mask_e"]])
        nz["eps_pi"] = torch.cat([noise["eps_pi"], noise["eps_g"], noise["eps_e"]])
        for k in ("mask_g", "mask_e", "eps_g", "eps_e"):
            nz.pop(k)
        losses = ag.update_from_batch({k: v.cuda().contiguous() for k, v in batch.items()}, _noise_cuda(nz),
                                      extra={k: v.cuda().contiguous() for k, v in extra.items()}).cpu().numpy()
        ref = g[f"step{s}_losses"]
        assert abs(losses[0] - ref[0]) <= FP32_TOL * max(1.0, abs(ref[0])), (s, losses, ref)
        assert abs(losses[1] - ref[1]) <= FP32_TOL * max(1.0, abs(ref[1])), (s, losses, ref)
        assert abs(float(ag.log_alpha) - float(g[f"step{s}_log_alpha"])) < 1e-6
        for nm, mod, od in (("actor", ag.policy, orc.actor), ("critic", ag.critic, orc.critic)):
            for k, p in mod.named_parameters():
                d = (p.detach().cpu() - od[k]).abs()
                assert float((d > 2e-5 * (s + 1)).float().mean()) < 2e-3, (s, nm, k)


def test_learn_guidence_agent_api():
    """The agent-level call: expert buffer fill, engaged transitions, graph-free fused update, finite losses."""
    cfg = O.Cfg(dim=32, depth=1, heads=2)
    ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, True, 5, BUFFER_SIZE=32, TAU=5e-4, POLICY_FREQ=1,
                GAMMA=0.999, ALPHA=1.0, block=1, head=2, l_f_size=32, buffer_size_expert=15, precision="bf16")
    rs = np.random.RandomState(0)
    for i in range(12):
        s, s2 = rs.rand(128, 160).astype(np.float32), rs.rand(128, 160).astype(np.float32)
        ag.store_transition(s, rs.rand(2) * 2 - 1, rs.rand(2), rs.rand(2), float(rs.randn()), s2, float(i % 5 == 0), None, 0)
        ag.initialize_expert_buffer(s, rs.rand(2) * 2 - 1, rs.rand(2), rs.rand(2), 1.0, s2, 0)
    for _ in range(3):
        q, p = ag.learn_guidence(False, 8)
        assert np.isfinite(q) and np.isfinite(p)


def test_behaviour_cloning_step_through_module_surface():
    """vn/attention_imitating.py:45-67 unchanged on top of the drop-in module: policy.sample -> RMSE on the
    clipped tanh-mean -> backward -> clip_grad_norm_(10) -> torch.optim.Adam(policy.parameters()).  Two steps
    against the same code running on the oracle's parameters (autograd on CPU)."""
    cfg = O.Cfg(dim=32, depth=2, heads=2)
    B = 6
    pa = reference_init("actor", cfg, 71)
    a = _mk("actor", cfg, pa)
    opt = torch.optim.Adam(a.parameters(), lr=1e-3)
    ref = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    ropt = torch.optim.Adam(list(ref.values()), lr=1e-3)
    for s in range(2):
        batch, nz = synthetic_batch(cfg, B, 80 + s), synthetic_noise(cfg, B, 90 + s)
        img, goal, act = batch["obs"], batch["pobs"], batch["act"]
        # reference-side (oracle) step
        _, _, mean = O.actor_sample(ref, img, goal, nz["eps_pi"], cfg, nz["mask_a"])
        rloss = torch.sqrt(torch.pow(mean.clip(-1, 1) - act, 2).mean())
        ropt.zero_grad()
        rloss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in ref.values() if p.grad is not None], 10)
        ropt.step()
        # drop-in module
        a.inject_noise(mask=nz["mask_a"], eps=nz["eps_pi"])
        _, _, m = a.sample([img.cuda(), goal.cuda()])
        loss = torch.sqrt(torch.pow(m.clip(-1, 1) - act.cuda(), 2).mean())
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(a.parameters(), 10)
        opt.step()
        assert abs(float(loss) - float(rloss)) < FP32_TOL * max(1.0, abs(float(rloss)))
        for k, p in a.named_parameters():
            d = (p.detach().cpu() - ref[k].detach()).abs()
            assert float((d > 2e-5 * (s + 1)).float().mean()) < 5e-3, (s, k)
    assert a._bound()          # torch.optim updated the flat arena in place through the parameter views


def _one_update(cfg, precision, B, opts):
    """losses + both gradient arenas of one fused update with library options `opts` set during the call"""
    for k, v in opts.items():
        L.check(L.lib().dgvit_set_option(k.encode(), v), "set_option")
    try:
        ag = _agent(cfg, precision, seed=9)
        batch, noise = synthetic_batch(cfg, B, 51), synthetic_noise(cfg, B, 52)
        cb = {k: v.reshape(B, -1).cuda().contiguous() for k, v in batch.items()}
        losses = ag.update_from_batch(cb, _noise_cuda(noise)).cpu()
        torch.cuda.synchronize()
        return losses, ag.critic._garena.cpu().clone(), ag.policy._garena.cpu().clone()
    finally:
        for k, v in dict(mlp_split=1, attention_row0=3, mlp_front=1).items():
            L.check(L.lib().dgvit_set_option(k.encode(), v), "set_option")


@pytest.mark.parametrize("precision,B", [("bf16", 16), ("bf16", 3), ("fp32", 5)])
def test_pruned_block_variants_are_equivalent(precision, B):
    """The last-block shortcuts are exact rewrites, so switching them off must not change the update beyond summation
    order: split-hidden cluster MLP kernels (few token tiles) vs one CTA per tile; single-query-row attention
    (forward and backward) vs the full attention kernels; out-projection + LayerNorm-2 as the MLP kernel's prologue vs
    the separate GEMM."""
    cfg = O.Cfg()
    base = _one_update(cfg, precision, B, dict(mlp_split=0, attention_row0=0, mlp_front=0))
    for opts in (dict(mlp_split=1, attention_row0=0, mlp_front=0), dict(mlp_split=0, attention_row0=2, mlp_front=0),
                 dict(mlp_split=0, attention_row0=0, mlp_front=1), dict(mlp_split=1, attention_row0=3, mlp_front=1)):
        got = _one_update(cfg, precision, B, opts)
        tol = 2e-5 if precision == "fp32" else 2e-2
        assert torch.allclose(got[0], base[0], rtol=tol, atol=tol * 1e-2), (opts, got[0], base[0])
        for g, b in zip(got[1:], base[1:]):
            err = float((g - b).norm() / b.norm().clamp_min(1e-20))
            assert err < tol, (opts, err)


def _async_run(steps, B, opts):
    lib = L.lib()
    for k, v in opts.items():
        L.check(lib.dgvit_set_option(k.encode(), v), "set_option")
    try:
        ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 11, LR_C=1e-3, LR_A=1e-3,
                    LR_ALPHA=1e-4, BUFFER_SIZE=512, TAU=5e-4, POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0, block=4, head=4,
                    l_f_size=64, precision="bf16")
        ag.replay_buffer.fill_synthetic(512, seed=3)
        for _ in range(steps):              # eager, captured, then replayed; the host never waits for the GPU
            ag.learn_async(B)
        torch.cuda.synchronize()
        return ag._losses.clone(), ag.policy._arena.clone(), ag.critic._arena.clone(), ag.critic_target._arena.clone()
    finally:
        for k in opts:
            L.check(lib.dgvit_set_option(k.encode(), 1), "set_option")


@pytest.mark.gpu
def test_async_updates_are_reproducible_and_stream_independent():
    """Back-to-back `learn_async` steps (host running ahead of the GPU, pinned index staging ring, CUDA-graph replay) give
    bit-identical parameters on a second run, and the forked streams inside the library (three forward passes of phase 1,
    weight-gradient lane of the backward) change nothing: every kernel does the same work in the same summation order,
    whichever stream it runs on."""
    a = _async_run(8, 64, {})
    b = _async_run(8, 64, {})
    c = _async_run(8, 64, dict(fork_streams=0, bwd_side=0))
    for x, y, z in zip(a, b, c):
        assert torch.equal(x, y)
        assert torch.equal(x, z)
