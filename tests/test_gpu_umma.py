"""GPU: tcgen05 / TMEM building blocks (csrc/umma.cuh) — operand layouts, both major modes, fp16 hi/lo split."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, N, K, a_mn, b_mn):
    Am = A.T if a_mn else A          # -> [M, K]
    Bm = B.T if b_mn else B          # -> [N, K]
    return Am[:128, :K].double() @ Bm[:N, :K].double().T


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("N,K", [(128, 64), (64, 128), (32, 32), (16, 16)])
def test_umma_fp16_exact_inputs(a_mn, b_mn, N, K):
    """fp16-representable inputs, single product: exact up to FP32 accumulation."""
    from ddrl_b200 import kernels as Kn
    g = torch.Generator().manual_seed(N * 1000 + K + 2 * a_mn + b_mn)
    A = (torch.randint(-8, 9, (K, 128) if a_mn else (128, K), generator=g).float() / 8.0).cuda()
    B = (torch.randint(-8, 9, (K, N) if b_mn else (N, K), generator=g).float() / 8.0).cuda()
    D, status = Kn.umma_selftest(A, B, N, K, a_mn, b_mn, False)
    assert status == 0, "tcgen05 MMA did not complete"
    ref = _ref(A.cpu(), B.cpu(), N, K, a_mn, b_mn)
    assert torch.equal(D.cpu().double(), ref)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, True), (False, True)])
def test_umma_split_reaches_fp32_level_accuracy(a_mn, b_mn):
    from ddrl_b200 import kernels as Kn
    N, K = 128, 128
    g = torch.Generator().manual_seed(7)
    A = torch.randn((K, 128) if a_mn else (128, K), generator=g).cuda()
    B = torch.randn((K, N) if b_mn else (N, K), generator=g).cuda()
    D, status = Kn.umma_selftest(A, B, N, K, a_mn, b_mn, True)
    assert status == 0
    ref = _ref(A.cpu(), B.cpu(), N, K, a_mn, b_mn)
    err = float((D.cpu().double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, err
    D1, _ = Kn.umma_selftest(A, B, N, K, a_mn, b_mn, False)     # single fp16 product for contrast
    err1 = float((D1.cpu().double() - ref).abs().max() / ref.abs().max())
    assert err1 > 20 * err
