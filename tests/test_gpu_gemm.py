"""GPU: the tcgen05/TMA bf16 GEMM against the CUDA-core GEMM on identical bf16 operands, for
every operand layout and shape family the hot path uses (forward NT, dX NN, dW TN split-K)."""
import ctypes as C

import pytest
import torch

from dgvit_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _run(M, N, K, A, a_sm, a_sk, Bm, b_sk, b_sn, splitk, tc):
    Cout = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    part = torch.empty(max(splitk, 1) * M * N, device="cuda", dtype=torch.float32)
    rc = L.lib().dgvit_gemm_bf16(M, N, K, A.data_ptr(), a_sm, a_sk, Bm.data_ptr(), b_sk, b_sn, Cout.data_ptr(), N,
                                 splitk, part.data_ptr(), int(tc), torch.cuda.current_stream().cuda_stream)
    L.check(rc, "gemm_bf16")
    torch.cuda.synchronize()
    return Cout


def _check(M, N, K, layout, splitk=1):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    if layout == "NT":      # y = x W^T : A [M,K], B = W [N,K]
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        Bm = torch.randn(N, K, device="cuda", generator=g).bfloat16()
        st = (K, 1, 1, K)
        ref = A.float() @ Bm.float().t()
    elif layout == "NN":    # dx = dy W : A [M,K], B = W [K,N]
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        Bm = torch.randn(K, N, device="cuda", generator=g).bfloat16()
        st = (K, 1, N, 1)
        ref = A.float() @ Bm.float()
    else:                   # dW = dy^T x : A = dy [K,M], B = x [K,N]
        A = torch.randn(K, M, device="cuda", generator=g).bfloat16()
        Bm = torch.randn(K, N, device="cuda", generator=g).bfloat16()
        st = (1, M, N, 1)
        ref = A.float().t() @ Bm.float()
    simt = _run(M, N, K, A, st[0], st[1], Bm, st[2], st[3], splitk, False)
    tc = _run(M, N, K, A, st[0], st[1], Bm, st[2], st[3], splitk, True)
    scale = float(ref.abs().max())
    assert float((simt - ref).abs().max()) < 2e-3 * scale          # sanity of the cross-check itself
    err = float((tc - simt).abs().max()) / scale
    assert err < 2e-5, (layout, M, N, K, splitk, err)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 64), (1040, 768, 64), (640, 64, 256), (384, 2048, 64),
                                   (512, 64, 2048), (1024, 64, 320), (130, 64, 128), (200, 328, 72)])
def test_gemm_nt(M, N, K):
    _check(M, N, K, "NT")


@pytest.mark.parametrize("M,N,K", [(256, 2048, 64), (512, 64, 2048), (384, 256, 64), (640, 64, 768), (130, 128, 136)])
def test_gemm_nn(M, N, K):
    _check(M, N, K, "NN")


@pytest.mark.parametrize("M,N,K,splitk", [(2048, 64, 1040, 1), (2048, 64, 4160, 4), (64, 2048, 2080, 2),
                                          (768, 64, 1300, 3), (64, 256, 650, 1), (64, 320, 1024, 2),
                                          (128, 128, 70, 1)])
def test_gemm_tn(M, N, K, splitk):
    _check(M, N, K, "TN", splitk)
