"""Known-answer test of the tcgen05/TMEM plumbing (umma.cuh) the tensor-core recurrence kernels are built on."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from rnnwavefunctions_b200 import ops  # noqa: E402


@pytest.mark.parametrize("N,K", [(160, 50), (112, 56), (64, 50), (16, 8), (256, 64)])
def test_umma_matches_fp64_matmul(N, K):
    g = torch.Generator(device="cpu").manual_seed(N * 100 + K)
    a = torch.randn(128, K, generator=g).cuda()
    b = torch.randn(N, K, generator=g).cuda()
    ref = (a.double() @ b.double().T).cpu().numpy()
    scale = np.abs(ref).max()
    d1 = ops.umma_selftest(a, b, passes=1).cpu().numpy()
    d3 = ops.umma_selftest(a, b, passes=3).cpu().numpy()
    e1 = np.abs(d1 - ref).max() / scale
    e3 = np.abs(d3 - ref).max() / scale
    assert e1 < 3e-3, e1          # single TF32 pass: 10-bit mantissa
    assert e3 < 3e-6, e3          # 3xTF32: FP32-grade


@pytest.mark.parametrize("N,K,dcol", [(160, 51, 0), (160, 64, 52), (176, 50, 4), (16, 16, 12)])
def test_umma_f16x3_matches_fp64_matmul(N, K, dcol):
    g = torch.Generator(device="cpu").manual_seed(N * 100 + K)
    a = (torch.rand(128, K, generator=g) * 2 - 1).cuda()            # hidden states live in (-1, 1)
    b = (torch.randn(N, K, generator=g) * 0.3).cuda()
    ref = (a.double() @ b.double().T).cpu().numpy()
    scale = np.abs(ref).max()
    d1 = ops.umma_selftest(a, b, passes=1, f16=True, dcol=dcol).cpu().numpy()
    d3 = ops.umma_selftest(a, b, passes=3, f16=True, dcol=dcol).cpu().numpy()
    e1 = np.abs(d1 - ref).max() / scale
    e3 = np.abs(d3 - ref).max() / scale
    assert e1 < 3e-3, e1          # single FP16 pass: 11-bit significand
    assert e3 < 3e-6, e3          # 3xFP16 (hi*hi + lo*hi + hi*lo, FP32 accumulate): FP32-grade
