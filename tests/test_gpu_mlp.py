"""GPU: fused tcgen05 GELU-MLP kernels (FeedForward.forward + residual, vn/GoalFormer.py:39-50,104), forward and
backward, through the C ABI, against an fp32 torch reference of the same op (exact-erf GELU) on the same bf16 inputs."""
import pytest
import torch

from dgvit_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _ref(x, W1, b1, W2, b2, resid, d_y=None):
    xf = x.float().requires_grad_(True)
    W1f, W2f = W1.float().requires_grad_(True), W2.float().requires_grad_(True)
    b1f, b2f = b1.clone().requires_grad_(True), b2.clone().requires_grad_(True)
    y = torch.nn.functional.gelu(xf @ W1f.t() + b1f) @ W2f.t() + b2f
    out = (resid + y).detach()
    if d_y is None:
        return out, None
    y.backward(d_y.float())
    return out, (xf.grad, W1f.grad, b1f.grad, W2f.grad, b2f.grad)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.mark.parametrize("rows,hid", [(128, 128), (65, 256), (256, 2048), (1000, 2048), (16640, 2048), (33000, 1024)])
def test_fused_mlp_forward_backward(rows, hid):
    g = torch.Generator(device="cuda").manual_seed(rows + hid)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x = rn(rows, 64).bfloat16()
    W1 = (rn(hid, 64) * 0.125).bfloat16()
    W2 = (rn(64, hid) * (hid ** -0.5)).bfloat16()
    b1, b2 = rn(hid) * 0.3, rn(64) * 0.3
    resid = rn(rows, 64)
    d_y = (rn(rows, 64) * 1e-3).bfloat16()          # gradient-sized values (f16 would flush these)
    out_ref, (dx_r, dW1_r, db1_r, dW2_r, db2_r) = _ref(x, W1, b1, W2, b2, resid, d_y)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.full((rows, 64), float("nan"), device="cuda")
    L.check(L.lib().dgvit_mlp_bf16(x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(),
                                   resid.data_ptr(), out.data_ptr(), None, None, None, None, None, rows, hid, st), "mlp fwd")
    torch.cuda.synchronize()
    y_ref = out_ref - resid
    assert float(((out - resid) - y_ref).abs().max() / y_ref.abs().max()) < 1e-2
    # the variant the update runs: f16 hidden tile x f16 copy of W2 (the f16 copy of a bf16 value is exact here: same W2)
    out16 = torch.full((rows, 64), float("nan"), device="cuda")
    W2h = W2.float().half()
    L.check(L.lib().dgvit_mlp_fwd_f16w2(x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2h.data_ptr(), b2.data_ptr(),
                                        resid.data_ptr(), out16.data_ptr(), rows, hid, st), "mlp fwd f16w2")
    torch.cuda.synchronize()
    assert float(((out16 - resid) - y_ref).abs().max() / y_ref.abs().max()) < 1e-2
    n = L.lib().dgvit_mlp_partial_floats(rows, hid)
    assert n > 0
    partial = torch.empty(n, device="cuda")
    d_x = torch.full((rows, 64), float("nan"), device="cuda")
    d_w = torch.full((hid * 64 * 2 + hid,), float("nan"), device="cuda")
    d_b2 = torch.full((64,), float("nan"), device="cuda")
    L.check(L.lib().dgvit_mlp_bf16(x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), None, None, None,
                                   d_y.data_ptr(), d_x.data_ptr(), d_w.data_ptr(), d_b2.data_ptr(), partial.data_ptr(),
                                   rows, hid, st), "mlp bwd")
    torch.cuda.synchronize()
    dW1, db1, dW2 = d_w[:hid * 64].view(hid, 64), d_w[hid * 64:hid * 65], d_w[hid * 65:].view(64, hid)
    assert _rel(d_x, dx_r) < 1.5e-2, ("dx", _rel(d_x, dx_r))
    assert _rel(dW1, dW1_r) < 1.5e-2, ("dW1", _rel(dW1, dW1_r))
    assert _rel(db1, db1_r) < 1.5e-2, ("db1", _rel(db1, db1_r))
    assert _rel(dW2, dW2_r) < 1.5e-2, ("dW2", _rel(dW2, dW2_r))
    assert _rel(d_b2, db2_r) < 1e-2, ("db2", _rel(d_b2, db2_r))
