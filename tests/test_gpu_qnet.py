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
