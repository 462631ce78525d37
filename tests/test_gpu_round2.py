"""GPU parity tests added in round 2: the timed configuration itself (bf16, B=256, shipped preset), the in-kernel RNG
streams of the graph-replayed update, the stand-alone trunk entry, the replay write path, and the regressions the
round-1 review named (dropout counter bound to the call, per-batch-size graph buffers)."""
import ctypes as C

import numpy as np
import pytest
import torch

import dgvit_b200 as dg
from dgvit_b200 import _lib as L
from helpers import O, SEED, load_params, reference_init, relerr, synthetic_batch, synthetic_noise

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4
BF16_TOL = 1e-2


def _agent(cfg, precision, seed=SEED, **kw):
    return dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, seed, LR_C=1e-3, LR_A=1e-3,
                  LR_ALPHA=1e-4, BUFFER_SIZE=kw.pop("BUFFER_SIZE", 64), TAU=5e-4, POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0,
                  block=cfg.depth, head=cfg.heads, l_f_size=cfg.dim, automatic_entropy_tuning=True, precision=precision,
                  image_size=(cfg.img_h, cfg.img_w), **kw)


def _noise_cuda(noise):
    return {k: (v.cuda().to(torch.uint8).contiguous() if k.startswith("mask") else v.cuda().contiguous())
            for k, v in noise.items() if v is not None}


def _mk(kind, cfg, params, precision="fp32"):
    cls = dg.GoTPolicy if kind == "actor" else dg.GoTQNetwork
    m = cls(2, 2, cfg.depth, cfg.heads, cfg.dim, image_size=(cfg.img_h, cfg.img_w))
    load_params(m, params)
    m = m.to("cuda")
    m.precision = precision
    return m


def _grad_checks(ag, orc, cos_min=0.99, lo=0.95, hi=1.05):
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
            assert cos > cos_min and lo < ratio < hi, (k, cos, ratio)


# ----------------------------------------------------------------------------------------------- the timed configuration
def test_update_bf16_B256_shipped():
    """BASELINE config[1] itself: ONE fused update at B=256, shipped preset (D=64, L=4, H=4), bf16 operands — where all
    130 token tiles, the side-stream backward and the split-K reductions are live — against SACOracle.learn (fp32, CPU)
    on the same minibatch and injected noise: Q / actions <= 1e-2 relative, the losses, and every parameter gradient
    (cosine > 0.99, norm within 5 %)."""
    cfg = O.Cfg()
    B = 256
    ag = _agent(cfg, "bf16", seed=SEED)
    actor0 = {k: v.detach().cpu().clone() for k, v in ag.policy.named_parameters()}
    critic0 = {k: v.detach().cpu().clone() for k, v in ag.critic.named_parameters()}
    orc = O.SACOracle(actor0, critic0, cfg)
    batch, noise = synthetic_batch(cfg, B, 131), synthetic_noise(cfg, B, 132)
    want = orc.learn(batch, noise)
    cb = {k: v.reshape(B, -1).cuda().contiguous() for k, v in batch.items()}
    dbg = torch.zeros(B * 11, device="cuda")
    got = ag.update_from_batch(cb, _noise_cuda(noise), debug=dbg).tolist()
    assert abs(got[0] - want[0]) <= BF16_TOL * abs(want[0]), (got, want)
    # (the policy loss reads Q(s, pi) of the critic AFTER its Adam step, see the q1p bound below)
    assert abs(got[1] - want[1]) <= 0.1 * max(1.0, abs(want[1])), (got, want)
    d = dbg.cpu()
    n2 = B * 2
    assert relerr(d[0:n2].reshape(B, 2), orc.last["nq"]) < BF16_TOL           # TD target
    assert relerr(d[n2:2 * n2].reshape(B, 2), orc.last["q1"]) < BF16_TOL      # Q1(s, a)
    assert relerr(d[2 * n2:3 * n2].reshape(B, 2), orc.last["q2"]) < BF16_TOL  # Q2(s, a)
    assert relerr(d[3 * n2:4 * n2].reshape(B, 2), orc.last["pi"]) < BF16_TOL  # actions
    # Q(s, pi) is evaluated by the critic AFTER its Adam step: the first step moves every weight by ~lr * sign(g), so the
    # gradients whose sign the bf16 rounding flips (|g| ~ 0) move the bf16 and the fp32 critic apart by O(lr): loose bound
    assert relerr(d[4 * n2:5 * n2].reshape(B, 2), orc.last["q1p"]) < 0.3
    _grad_checks(ag, orc)
    for mod in (ag.policy, ag.critic, ag.critic_target):
        assert torch.isfinite(mod._arena).all()


def test_wide_deep_variant_depth6_bf16_gradients():
    """BASELINE config 5 at its real depth: 256x320 frames (257 tokens), D=128, 6 heads, 6 blocks, B=8,
    one fused bf16 update vs the fp32 oracle: Q / actions <= 1e-2, gradients cosine > 0.99 and norm within 5 %."""
    cfg = O.Cfg(dim=128, depth=6, heads=6, img_h=256, img_w=320)
    B = 8
    ag = _agent(cfg, "bf16", seed=21, BUFFER_SIZE=4)
    actor0 = {k: v.detach().cpu().clone() for k, v in ag.policy.named_parameters()}
    critic0 = {k: v.detach().cpu().clone() for k, v in ag.critic.named_parameters()}
    orc = O.SACOracle(actor0, critic0, cfg)
    batch, noise = synthetic_batch(cfg, B, 141), synthetic_noise(cfg, B, 142)
    want = orc.learn(batch, noise)
    cb = {k: v.reshape(B, -1).cuda().contiguous() for k, v in batch.items()}
    dbg = torch.zeros(B * 11, device="cuda")
    got = ag.update_from_batch(cb, _noise_cuda(noise), debug=dbg).tolist()
    assert abs(got[0] - want[0]) <= 2e-2 * abs(want[0]), (got, want)
    d = dbg.cpu()
    n2 = B * 2
    assert relerr(d[n2:2 * n2].reshape(B, 2), orc.last["q1"]) < BF16_TOL
    assert relerr(d[3 * n2:4 * n2].reshape(B, 2), orc.last["pi"]) < BF16_TOL
    _grad_checks(ag, orc, cos_min=0.985, lo=0.93, hi=1.07)


# ----------------------------------------------------------------------------------------------- in-kernel RNG streams
def _drop(mode, rng=None, mask=None, stream_id=0):
    return L.Drop(mode=mode, p=0.1, keep_mask=L.ptr(mask), rng_state=L.ptr(rng), stream_id=stream_id)


def _dump_mask(rng, B, N, D, offset=0, stream_id=0):
    out = torch.zeros(B, N, D, dtype=torch.uint8, device="cuda")
    d = _drop(L.DROP_RNG, rng, stream_id=stream_id)
    L.check(L.lib().dgvit_debug_drop_mask(C.byref(d), B, N, D, offset, out.data_ptr(), None), "debug_drop_mask")
    return out


def _actor_call(m, img, ps, drop, eps=None, offset=0):
    """dgvit_actor_forward through ctypes; returns (action, log_prob, mean_t, eps_used)."""
    B, na = img.shape[0], 2
    z = lambda *s: torch.zeros(*s, device="cuda")
    mean, lstd, act, lp, mt, eo = z(B, na), z(B, na), z(B, na), z(B, 1), z(B, na), z(B, na)
    one, zero = torch.ones(na, device="cuda"), torch.zeros(na, device="cuda")
    net = m.net_struct()                 # (binds the arena)
    ws = m._workspace(B, False)
    io = L.ActorIO(img=img.data_ptr(), pstate=ps.data_ptr(), eps=L.ptr(eps), action_scale=one.data_ptr(),
                   action_bias=zero.data_ptr(), drop=drop, sample_offset=offset, mean=mean.data_ptr(),
                   log_std=lstd.data_ptr(), action=act.data_ptr(), log_prob=lp.data_ptr(), mean_t=mt.data_ptr(),
                   eps_out=eo.data_ptr())
    if m.precision == "bf16":
        m.refresh_shadow()
    L.check(L.lib().dgvit_actor_forward(C.byref(net), C.byref(io), B, m._precision_code(), 0, ws.data_ptr(), ws.numel(),
                                        None), "actor_forward")
    torch.cuda.synchronize()
    return act, lp, mt, eo


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_in_kernel_dropout_and_rsample_streams(precision):
    """The RNG mode the benchmark runs (DGVIT_DROP_RNG, eps == NULL), pinned to the injected-noise mode the parity tests
    use: (1) the keep decisions the kernels evaluate in place have rate 1 - p = 0.9 and (2) ARE the mask
    dgvit_debug_drop_mask dumps — a DROP_MASK run with that mask and the generated eps reproduces the outputs bit for bit,
    which also pins the 1 / 0.9 scaling; (3) the generated eps are N(0,1); (4) a data-parallel rank with
    sample_offset = k draws exactly rows [k, k + count) of the single-GPU draw (mask and eps)."""
    cfg = O.Cfg(dim=64, depth=2, heads=4)
    N, D = cfg.n_tokens, cfg.dim
    m = _mk("actor", cfg, reference_init("actor", cfg, 5), precision)
    B = 64
    batch = synthetic_batch(cfg, B, 7)
    img, ps = batch["obs"].cuda().contiguous(), batch["pobs"].cuda().contiguous()
    rng = torch.tensor([SEED, 17], dtype=torch.int64, device="cuda")
    mask = _dump_mask(rng, B, N, D, stream_id=4)
    rate = float(mask.float().mean())
    assert abs(rate - 0.9) < 0.005, rate
    other = _dump_mask(rng, B, N, D, stream_id=1)
    assert float((other != mask).float().mean()) > 0.1          # the passes of one update draw different masks
    a_rng = _actor_call(m, img, ps, _drop(L.DROP_RNG, rng, stream_id=4))
    a_msk = _actor_call(m, img, ps, _drop(L.DROP_MASK, mask=mask), eps=a_rng[3])
    for x, y in zip(a_rng, a_msk):
        assert torch.equal(x, y)
    # the 1/(1-p) scaling itself, against the oracle with the dumped mask
    with torch.no_grad():
        oa, olp, omt = O.actor_sample(reference_init("actor", cfg, 5), batch["obs"], batch["pobs"], a_rng[3].cpu(), cfg,
                                      mask.float().cpu())
    tol = FP32_TOL if precision == "fp32" else BF16_TOL
    assert relerr(a_rng[0], oa) < tol and relerr(a_rng[2], omt) < tol
    # data-parallel slicing: rank with offset k sees rows [k, k+count) of the global draw
    k, cnt = 24, 16
    part = _dump_mask(rng, cnt, N, D, offset=k, stream_id=4)
    assert torch.equal(part, mask[k:k + cnt])
    a_part = _actor_call(m, img[k:k + cnt].contiguous(), ps[k:k + cnt].contiguous(), _drop(L.DROP_RNG, rng, stream_id=4), offset=k)
    assert torch.equal(a_part[3], a_rng[3][k:k + cnt])
    if precision == "fp32":
        assert torch.equal(a_part[0], a_rng[0][k:k + cnt])
    else:       # bf16: the 16-row call takes the split-hidden cluster kernels (other summation order than the 64-row call)
        assert relerr(a_part[0], a_rng[0][k:k + cnt]) < BF16_TOL
    # eps ~ N(0,1): many rows through a shallow network
    cfg1 = O.Cfg(dim=32, depth=1, heads=2)
    m1 = _mk("actor", cfg1, reference_init("actor", cfg1, 6), "fp32")
    Bn = 4096
    g = torch.Generator().manual_seed(3)
    eps = _actor_call(m1, torch.rand(Bn, 128, 160, generator=g).cuda(), torch.rand(Bn, 2, generator=g).cuda(),
                      _drop(L.DROP_RNG, rng, stream_id=1))[3].cpu().double()
    assert abs(float(eps.mean())) < 0.05 and abs(float(eps.var()) - 1.0) < 0.06
    assert abs(float((eps.abs() < 1).double().mean()) - 0.6827) < 0.02
    assert abs(float((eps ** 4).mean()) - 3.0) < 0.35                 # kurtosis of a Gaussian


def test_rng_update_equals_injected_update():
    """One fused update in RNG mode (noise == NULL: what learn / learn_async / bench.py run) equals the same update with
    the five dropout masks and the two rsample draws INJECTED, where the injected inputs are the dumps of the same
    Philox streams (stream ids 1..5, counter of the update).  fp32: losses and both gradient arenas agree to rounding."""
    cfg = O.Cfg(dim=64, depth=2, heads=4)
    B = 12
    N, D = cfg.n_tokens, cfg.dim
    batch = synthetic_batch(cfg, B, 77)
    cb = {k: v.reshape(B, -1).cuda().contiguous() for k, v in batch.items()}
    a = _agent(cfg, "fp32", seed=9)
    b = _agent(cfg, "fp32", seed=9)
    rng = a._rng.clone()
    # the generated streams of this update: masks by stream id, eps through the actor entry with the same ids
    masks = {name: _dump_mask(rng, B, N, D, stream_id=sid) for name, sid in
             (("mask_a_next", 1), ("mask_ct", 2), ("mask_c", 3), ("mask_a", 4), ("mask_c_pi", 5))}
    img2, ps2 = cb["next_obs"].view(B, 128, 160), cb["next_pobs"]
    img1, ps1 = cb["obs"].view(B, 128, 160), cb["pobs"]
    eps_next = _actor_call(a.policy, img2, ps2, _drop(L.DROP_RNG, rng, stream_id=1))[3]
    eps_pi = _actor_call(a.policy, img1, ps1, _drop(L.DROP_RNG, rng, stream_id=4))[3]
    la = a.update_from_batch(cb).clone()
    lb = b.update_from_batch(cb, dict(eps_next=eps_next, eps_pi=eps_pi, **masks)).clone()
    torch.cuda.synchronize()
    assert torch.allclose(la, lb, rtol=1e-6, atol=1e-7), (la, lb)
    for ma, mb in ((a.critic, b.critic), (a.policy, b.policy)):
        err = float((ma._garena - mb._garena).norm() / mb._garena.norm())
        assert err < 1e-6, err
    assert int(a._rng[1]) == int(rng[1]) + 1                  # the counter advances once per update


def test_dropout_counter_is_bound_to_the_call():
    """Two forwards on one module, then one backward (learn_guidence with the CNN critic: policy.sample(s), then
    policy.sample(imitation rows), then a single policy_loss.backward()): the backward of the FIRST pass must regenerate
    the mask that pass used, although the module's counter has moved on.  Gradients equal the run where each forward is
    followed immediately by its own backward."""
    cfg = O.Cfg(dim=32, depth=2, heads=2)
    pa = reference_init("actor", cfg, 3)
    B = 6
    b1, b2 = synthetic_batch(cfg, B, 1), synthetic_batch(cfg, B, 2)
    eps1, eps2 = torch.randn(B, 2, generator=torch.Generator().manual_seed(1)), torch.randn(B, 2, generator=torch.Generator().manual_seed(2))

    def run(interleaved):
        torch.manual_seed(1234)                       # seeds the module's dropout stream ({initial_seed, counter})
        m = _mk("actor", cfg, pa)
        m.train()
        grads = []
        x1 = [b1["obs"].cuda(), b1["pobs"].cuda()]
        x2 = [b2["obs"].cuda(), b2["pobs"].cuda()]
        if interleaved:
            _, lp1, mt1 = m._run(x1, eps=eps1.cuda())[2:]
            (lp1.mean() + (mt1 ** 2).sum()).backward()
            g1 = [p.grad.clone() for p in m.parameters() if p.grad is not None]
            m.zero_grad()
            _, lp2, mt2 = m._run(x2, eps=eps2.cuda())[2:]
            (lp2.mean() + (mt2 ** 2).sum()).backward()
            g2 = [p.grad.clone() for p in m.parameters() if p.grad is not None]
            grads = [a + b for a, b in zip(g1, g2)]
        else:
            _, lp1, mt1 = m._run(x1, eps=eps1.cuda())[2:]
            _, lp2, mt2 = m._run(x2, eps=eps2.cuda())[2:]
            ((lp1.mean() + (mt1 ** 2).sum()) + (lp2.mean() + (mt2 ** 2).sum())).backward()
            grads = [p.grad.clone() for p in m.parameters() if p.grad is not None]
        return grads

    ga, gb = run(True), run(False)
    assert len(ga) == len(gb) > 0
    for x, y in zip(ga, gb):
        assert relerr(x, y) < 1e-5


def test_graph_buffers_follow_the_batch_size():
    """use_cuda_graph=True with alternating batch sizes (learn(64), learn(128), learn(64) ...): every replayed graph must
    find the index / minibatch buffers it was captured with.  Same parameters as the graph-free agent, bit for bit."""
    def run(graph):
        ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 11, BUFFER_SIZE=300, TAU=5e-4,
                    POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0, block=2, head=2, l_f_size=32, precision="bf16",
                    use_cuda_graph=graph)
        ag.replay_buffer.fill_synthetic(300, seed=3)
        for B in (64, 64, 128, 64, 128, 128, 64, 64, 128):
            ag.learn_async(B)
        torch.cuda.synchronize()
        return ag.policy._arena.clone(), ag.critic._arena.clone(), ag._loss_buffer().clone()
    for x, y in zip(run(True), run(False)):
        assert torch.equal(x, y)


# ----------------------------------------------------------------------------------------------- GoT.forward
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_got_forward_trunk_entry(precision, tol):
    """GoT.forward(img, goal) (vn/GoalFormer.py:156-171) through dgvit_trunk_forward / dgvit_trunk_backward: the trunk of
    an owning GoTPolicy (its arena) and a stand-alone GoT, against the oracle's trunk_forward, outputs and gradients."""
    cfg = O.Cfg(dim=64, depth=3, heads=4)
    pa = reference_init("actor", cfg, 31)
    B = 5
    batch = synthetic_batch(cfg, B, 32)
    img = batch["obs"]
    goal = torch.randn(B, cfg.dim, generator=torch.Generator().manual_seed(4))
    mask = (torch.rand(B, cfg.n_tokens, cfg.dim, generator=torch.Generator().manual_seed(5)) > 0.1).float()
    pg = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    goal_o = goal.clone().requires_grad_(True)
    z_o = O.trunk_forward(pg, img, goal_o, cfg, mask)
    wz = torch.randn(B, cfg.dim, generator=torch.Generator().manual_seed(6))     # (sum z^2 is constant after the RMSNorm)
    (z_o * wz).sum().backward()
    m = _mk("actor", cfg, pa, precision)
    m.inject_noise(mask=mask)
    goal_g = goal.cuda().requires_grad_(True)
    z = m.trans(img.cuda(), goal_g)
    assert z.shape == (B, cfg.dim) and relerr(z, z_o) < tol
    (z * wz.cuda()).sum().backward()
    gtol = 2e-4 if precision == "fp32" else 6e-2
    assert relerr(goal_g.grad, goal_o.grad) < gtol
    for k, p in m.named_parameters():
        if not k.startswith("trans.") or pg[k].grad is None:
            continue
        assert p.grad is not None, k
        if precision == "fp32":
            assert relerr(p.grad, pg[k].grad) < gtol, (k, relerr(p.grad, pg[k].grad))
        else:
            a, b = p.grad.double().flatten().cpu(), pg[k].grad.double().flatten()
            assert float((a @ b) / (a.norm() * b.norm() + 1e-300)) > 0.99, k
    # stand-alone trunk (no owner): same weights -> same output
    t = dg.GoT(image_size=(128, 160), patch_size=(16, 20), num_classes=2, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads,
               mlp_dim=2048, channels=1)
    with torch.no_grad():
        sd = t.state_dict()
        for k2 in sd:
            sd[k2].copy_(pa["trans." + k2])
    t = t.to("cuda").eval()
    t._backend().precision = precision
    with torch.no_grad():
        z2 = t(img.cuda(), goal.cuda())
        z_eval = O.trunk_forward(pa, img, goal, cfg, None)
    assert relerr(z2, z_eval) < tol


# ----------------------------------------------------------------------------------------------- replay write path
class _NextOfModel:
    """What cpprb's ``ReplayBuffer(size, next_of="obs")`` returns for ``sample`` indexes after a sequence of ``add``s: a
    ring of ``size`` transitions, each keeping its own obs / next_obs (the model stores them separately; the store under
    test shares frames between consecutive transitions like cpprb does)."""

    def __init__(self, size):
        self.size, self.rows = size, []

    def add(self, **kw):
        self.rows.append(kw)
        if len(self.rows) > self.size:
            self.rows.pop(0)


def test_replay_add_then_gather_follows_next_of_semantics():
    """store_transition (vn/DRL.py:449-467) -> sample-index gather: single adds, a batched add, ring wrap-around.  Every live
    transition returns exactly what was stored (bit-exact), the newest transition's next_obs survives, and after the ring
    has wrapped no live transition is paired with a frame of a newer one (contiguous trajectories: next_obs[k] == obs[k+1],
    the cpprb next_of contract)."""
    size, f = 10, 128 * 160
    st = dg.ReplayStore(size, (128, 160), 2, 2, "cuda", seed=1)
    model = _NextOfModel(size)
    rs = np.random.RandomState(0)
    frames = rs.rand(40, 128, 160).astype(np.float32)

    def tr(i):
        return dict(obs=frames[i], next_obs=frames[i + 1], act=rs.rand(2).astype(np.float32) * 2 - 1,
                    pobs=rs.rand(2).astype(np.float32), next_pobs=rs.rand(2).astype(np.float32), rew=np.float32(rs.randn()),
                    done=np.float32(i % 7 == 0), engage=np.float32(i % 3 == 0))

    def check():
        rows = st.live_rows()
        assert len(rows) == len(model.rows) == st.get_stored_size()
        idx = torch.as_tensor(rows, dtype=torch.int64, device="cuda")
        B = len(rows)
        out = {k: torch.full((B, w), -1.0, device="cuda") for k, w in
               dict(obs=f, next_obs=f, pobs=2, next_pobs=2, act=2, rew=1, done=1).items()}
        st.gather(idx, out)
        for j, want in enumerate(model.rows):
            for k in ("obs", "next_obs", "pobs", "next_pobs", "act", "rew", "done"):
                got = out[k][j].cpu().numpy().reshape(-1)
                assert np.array_equal(got, np.asarray(want[k], np.float32).reshape(-1)), (j, k)
            assert st.engage_host[rows[j]] == want["engage"]
        # sampled indexes are always live rows
        s = set(st.sample_indexes(256).tolist())
        assert s <= set(int(r) for r in rows)

    for i in range(6):                                  # single adds (the control loop), before the wrap
        t = tr(i)
        st.add(**t); model.add(**t)
    check()
    ts = [tr(i) for i in range(6, 17)]                  # one batched add that wraps the ring
    st.add(obs=np.stack([t["obs"] for t in ts]), next_obs=np.stack([t["next_obs"] for t in ts]),
           act=np.stack([t["act"] for t in ts]), pobs=np.stack([t["pobs"] for t in ts]),
           next_pobs=np.stack([t["next_pobs"] for t in ts]), rew=np.array([t["rew"] for t in ts]),
           done=np.array([t["done"] for t in ts]), engage=np.array([t["engage"] for t in ts]))
    for t in ts:
        model.add(**t)
    check()
    for i in range(17, 30):                             # single adds after the wrap
        t = tr(i)
        st.add(**t); model.add(**t)
        check()


def test_replay_save_load_and_demonstration_files(tmp_path):
    """save_transition / load_transition round trip (vn/DRL.py:505-510) and demonstration .npz ingestion
    (vn/demonstration.py:237-245 schema, vn/main.py:232-266): the expert buffer holds the episodes' transitions in order."""
    ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, True, 5, BUFFER_SIZE=12, block=1, head=2,
                l_f_size=32, buffer_size_expert=4, precision="fp32")
    rs = np.random.RandomState(1)
    for i in range(7):
        ag.store_transition(rs.rand(128, 160, 1).astype(np.float32), rs.rand(2), rs.rand(2), rs.rand(2), float(rs.randn()),
                            rs.rand(128, 160, 1).astype(np.float32), float(i % 2), None, 0)
    ag.save_transition(str(tmp_path), 3)
    rows = ag.replay_buffer.live_rows()
    before = {k: getattr(ag.replay_buffer, k)[torch.as_tensor(rows, device="cuda")].cpu() for k in ("obs", "act", "rew", "pobs")}
    ag2 = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 5, BUFFER_SIZE=12, block=1, head=2,
                 l_f_size=32, precision="fp32")
    ag2.load_transition(str(tmp_path / "3"))
    rows2 = ag2.replay_buffer.live_rows()
    assert len(rows2) == 7
    for k, v in before.items():
        assert torch.equal(getattr(ag2.replay_buffer, k)[torch.as_tensor(rows2, device="cuda")].cpu(), v), k
    # demonstrations: two episodes in the reference's schema
    files, total = [], 0
    allobs, allact = [], []
    for e, n in enumerate((3, 5)):
        obs, nobs = rs.rand(n, 128, 160, 1).astype(np.float32), rs.rand(n, 128, 160, 1).astype(np.float32)
        act = (rs.rand(n, 2) * 2 - 1).astype(np.float32)
        fn = str(tmp_path / f"demo_{e}.npz")
        np.savez(fn, obs=obs, act=act, goal=rs.rand(n, 3).astype(np.float32), reward=rs.randn(n).astype(np.float32),
                 next_obs=nobs, next_goal=rs.rand(n, 3).astype(np.float32), done=np.zeros(n, dtype=bool))
        files.append(fn); total += n
        allobs.append(obs); allact.append(act)
    assert ag.load_demonstrations(files) == total
    ex = ag.replay_buffer_expert
    assert ex.get_stored_size() == total
    r = torch.as_tensor(ex.live_rows(), device="cuda")
    assert np.array_equal(ex.obs[r].cpu().numpy(), np.concatenate(allobs).reshape(total, -1))
    assert np.array_equal(ex.act[r].cpu().numpy(), np.concatenate(allact))
    q, p = ag.learn_guidence(False, 4)
    assert np.isfinite(q) and np.isfinite(p)
    with pytest.raises(ValueError):                      # the legacy 4-channel frame-stacked demonstrations are rejected
        ag.replay_buffer.add(rs.rand(128, 160, 4).astype(np.float32), rs.rand(2), rs.rand(2), rs.rand(2), 0.0,
                             rs.rand(128, 160, 4).astype(np.float32))


def test_device_argument_without_set_device():
    """SAC(device="cuda:N") must run on that device whatever torch's current device is (the library switches to the device
    that owns the buffers).  With one GPU this exercises the guard on the current device only."""
    n = torch.cuda.device_count()
    dev = f"cuda:{n - 1}"
    cfg = O.Cfg(dim=32, depth=1, heads=2)
    torch.cuda.set_device(0)
    ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 3, BUFFER_SIZE=32, block=1, head=2, l_f_size=32,
                precision="bf16", device=dev)
    ag.replay_buffer.fill_synthetic(32, seed=1)
    q, p = ag.learn(8)
    assert np.isfinite(q) and np.isfinite(p)
    a = ag.choose_action(np.random.rand(128, 160, 1).astype(np.float32), np.array([0.3, 0.1], np.float32), True)
    assert a.shape == (2,)


@pytest.mark.gpu
def test_depth_streaming_rows_equal_tile_rows():
    """The rows off the centre band leave through the register-streaming role of the depth kernel, the band rows through the
    shared-memory tiles; with `depth_strip` = 0 every row takes the tile path.  Same pixels, same Philox draws: the two agree
    to rounding, for given and for drawn noise, on frame sizes whose width is / is not a multiple of the 30-column segments."""
    from dgvit_b200 import _lib as L
    rng = torch.tensor([77, 3], dtype=torch.int64, device="cuda")
    try:
        for (n, H, W) in ((3, 512, 640), (2, 64, 80), (1, 256, 320), (2, 40, 136), (1, 8, 8), (1, 1024, 1280)):
            raw = (torch.rand(n, H, W, generator=torch.Generator().manual_seed(H + W)) * 6 + 1).cuda()
            noise = torch.randn(n, H, W, generator=torch.Generator().manual_seed(W)).cuda() * 50
            outs = {}
            for strip in (0, -1, 8, 5, 64):
                L.check(L.lib().dgvit_set_option(b"depth_strip", strip), "set_option")
                outs[strip] = (dg.depth_augment(raw, noise).clone(), dg.depth_augment(raw, None, rng).clone())
            for strip in (-1, 8, 5, 64):
                for k in range(2):
                    assert float((outs[strip][k] - outs[0][k]).abs().max()) < 2e-6, (n, H, W, strip, k)
    finally:
        L.check(L.lib().dgvit_set_option(b"depth_strip", -1), "set_option")


@pytest.mark.gpu
def test_behaviour_cloning_step_single_call():
    """SURVEY §8 f3, vn/attention_imitating.py:48-67 as ONE library call (`GoTPolicy.bc_step` -> `dgvit_bc_step`): policy.sample
    -> RMSE on the clipped tanh-mean -> backward -> clip_grad_norm_ -> Adam.  Three fp32 steps against the same statements run
    with autograd on the oracle's parameters; the third step uses a tiny max_norm so the clip is active (scale << 1)."""
    from helpers import reference_init
    cfg = O.Cfg(dim=32, depth=2, heads=2)
    B = 6
    pa = reference_init("actor", cfg, 71)
    a = _mk("actor", cfg, pa)
    ref = {k: v.clone().requires_grad_(True) for k, v in pa.items()}
    ropt = torch.optim.Adam(list(ref.values()), lr=1e-3)
    for s, max_norm in enumerate((10.0, 10.0, 1e-2)):
        batch, nz = synthetic_batch(cfg, B, 80 + s), synthetic_noise(cfg, B, 90 + s)
        img, goal, act = batch["obs"], batch["pobs"], batch["act"]
        _, _, mean = O.actor_sample(ref, img, goal, nz["eps_pi"], cfg, nz["mask_a"])
        rloss = torch.sqrt(torch.pow(mean.clip(-1, 1) - act, 2).mean())
        ropt.zero_grad()
        rloss.backward()
        rnorm = torch.nn.utils.clip_grad_norm_([p for p in ref.values() if p.grad is not None], max_norm)
        ropt.step()
        a.inject_noise(mask=nz["mask_a"], eps=nz["eps_pi"])
        loss = a.bc_step(img.cuda(), goal.cuda(), act.cuda(), lr=1e-3, max_action=1.0, max_norm=max_norm)
        assert abs(float(loss) - float(rloss)) < 1e-4 * max(1.0, abs(float(rloss))), (s, float(loss), float(rloss))
        assert abs(float(a._bc["grad_norm"]) - float(rnorm)) < 2e-4 * max(1.0, float(rnorm)), (s, float(a._bc["grad_norm"]), float(rnorm))
        for k, p in a.named_parameters():
            d = (p.detach().cpu() - ref[k].detach()).abs()
            assert float((d > 2e-5 * (s + 1)).float().mean()) < 5e-3, (s, k, float(d.max()))
    # parameters the loss does not reach did not move (grad is None in the reference)
    for k in ("log_std_linear.weight", "log_std_linear.bias"):
        assert torch.equal(dict(a.named_parameters())[k].detach().cpu(), pa[k])
    # bf16 path: the in-kernel noise, twenty steps on one batch: the imitation loss goes down
    b = _mk("actor", cfg, pa)
    b.precision = "bf16"
    img, goal, act = (batch[k].cuda() for k in ("obs", "pobs", "act"))
    ls = [float(b.bc_step(img, goal, act, lr=1e-3)) for _ in range(20)]
    assert np.isfinite(ls).all() and np.mean(ls[-3:]) < 0.8 * np.mean(ls[:3]), ls


@pytest.mark.gpu
def test_update_is_run_to_run_deterministic_with_unfused_mlp():
    """Regression: with D != 64 the MLP backward runs as separate GEMMs whose split-K partials live at the start of the
    partial-sum workspace; they used to be written from the main stream while the previous block's deferred reduction (side
    stream) could still be reading its queued partial sums from the same place -> LayerNorm / bias gradients of the block
    above changed from run to run whenever another forward pass shifted the timing (`actor_s_when` = 2 made it 8 runs in 12).
    Same seed, same batches: the eager multi-stream update must reproduce bit for bit."""
    def run():
        ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 11, BUFFER_SIZE=300, TAU=5e-4,
                    POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0, block=2, head=2, l_f_size=32, precision="bf16")
        ag.replay_buffer.fill_synthetic(300, seed=3)
        for B in (64, 64, 128, 64, 128, 128, 64):
            ag.learn_async(B)
        torch.cuda.synchronize()
        return torch.cat([ag.policy._arena.flatten(), ag.critic._arena.flatten(), ag._loss_buffer().flatten()]).clone()
    try:
        L.check(L.lib().dgvit_set_option(b"actor_s_when", 2), "set_option")
        ref = run()
        for _ in range(6):
            assert torch.equal(run(), ref)
    finally:
        L.check(L.lib().dgvit_set_option(b"actor_s_when", 1), "set_option")


@pytest.mark.gpu
@pytest.mark.parametrize("critic", ["Transformer", "CNN"])
def test_learn_guidence_graph_replay_matches_eager(critic):
    """``learn_guidence`` (vn/DRL.py:187-301) from a CUDA graph: the engaged rows are padded to a multiple of 32 with rows of
    weight 0; same seeds -> same sampled rows, so the graph-replayed agent must follow the eager agent (fp32: the padding rows
    add exact zeros; only the split points of the reductions move with the row count)."""
    def run(graph):
        ag = dg.SAC(2, 2, "GaussianTransformer", critic, False, False, True, 5, LR_C=1e-3, LR_A=1e-3, LR_ALPHA=1e-4,
                    BUFFER_SIZE=128, TAU=5e-3, POLICY_FREQ=1, GAMMA=0.99, ALPHA=0.2, block=2, head=2, l_f_size=32,
                    precision="fp32", buffer_size_expert=64, use_cuda_graph=graph)
        rs = np.random.RandomState(0)
        for i in range(60):
            f1, f2 = rs.rand(128, 160).astype(np.float32), rs.rand(128, 160).astype(np.float32)
            ag.store_transition(f1, rs.rand(2) * 2 - 1, rs.rand(2), rs.rand(2), float(rs.randn()), f2, float(i % 3 == 0), None, 0)
            if i < 40:
                ag.initialize_expert_buffer(f1, rs.rand(2) * 2 - 1, rs.rand(2), rs.rand(2), 1.0, f2, 0)
        out = [ag.learn_guidence(False, 16) for _ in range(6)]
        torch.cuda.synchronize()
        return np.array(out), ag.policy._arena.detach().cpu().clone(), ag.critic._arena.detach().cpu().clone()
    (lg, pg, cg), (le, pe, ce) = run(True), run(False)
    assert np.isfinite(lg).all() and np.allclose(lg, le, rtol=2e-4, atol=1e-5), (lg, le)
    for a, b in ((pg, pe), (cg, ce)):
        d = (a - b).abs()
        assert float((d > 5e-5).float().mean()) < 5e-3 and float(d.max()) < 5e-3, float(d.max())
