"""GPU: tcgen05/TMEM attention kernels (forward + backward) against an fp32 torch reference of the
same op on the same bf16 inputs, and against the CUDA-core kernel."""
import pytest
import torch

from dgvit_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _attn(qkv, o, d_o, d_qkv, B, N, H, tc, stats=None):
    rc = L.lib().dgvit_attention_bf16(qkv.data_ptr(), o.data_ptr(), L.ptr(d_o), L.ptr(d_qkv), B, N, H, 64, int(tc),
                                      L.ptr(stats), torch.cuda.current_stream().cuda_stream)
    L.check(rc, "attention_bf16")
    torch.cuda.synchronize()


def _ref(qkv, B, N, H, d_o=None):
    x = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    q, k, v = x[0], x[1], x[2]
    p = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * N, H * 64)
    if d_o is None:
        return o.detach(), None
    o.backward(d_o.float())
    g = x.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * H * 64)
    return o.detach(), g


@pytest.mark.parametrize("B,N,H", [(1, 65, 1), (3, 65, 4), (40, 65, 4), (2, 33, 2), (2, 128, 3), (300, 65, 4),
                                   # more than one 128-row tile: the key-chunked kernels (257 tokens = BASELINE config 5)
                                   (1, 257, 1), (3, 257, 6), (2, 129, 2), (2, 256, 2), (2, 300, 3), (1, 384, 2), (70, 257, 6)])
def test_attention_forward_backward(B, N, H):
    g = torch.Generator(device="cuda").manual_seed(B * 100 + N + H)
    qkv = (torch.randn(B * N, 3 * H * 64, device="cuda", generator=g) * 1.5).bfloat16()
    d_o = torch.randn(B * N, H * 64, device="cuda", generator=g).bfloat16()
    o_ref, g_ref = _ref(qkv, B, N, H, d_o)
    n_stats = L.lib().dgvit_attention_stats_floats(B, N, H)
    assert (n_stats > 0) == (N > 128)
    stats = torch.full((max(n_stats, 1),), float("nan"), device="cuda") if n_stats else None
    for tc in (False, True):
        o = torch.full((B * N, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        _attn(qkv, o, None, None, B, N, H, tc, stats)
        err = float((o.float() - o_ref).abs().max() / o_ref.abs().max())
        assert err < 1.5e-2, ("fwd", tc, err)
        dq = torch.full((B * N, 3 * H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        _attn(qkv, o, d_o, dq, B, N, H, tc, stats)
        err = float((dq.float() - g_ref).abs().max() / g_ref.abs().max())
        assert err < 2e-2, ("bwd", tc, err)
