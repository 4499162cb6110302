"""GPU (-m gpu): the tcgen05 / TMEM / TMA bf16 path.  First the UMMA descriptor self-test (one tile), then
the fused MLP kernels against the fp32 CUDA path / oracle at the north star's bf16 tolerance."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def kmajor_blob(x):
    """[rows, K] -> chunk-major [K/8][rows][8] (tc_ptx.cuh)"""
    rows, K = x.shape
    return x.reshape(rows, K // 8, 8).permute(1, 0, 2).contiguous()


def mnmajor_blob(x):
    """[MN, K] -> [MN/8][K][8]: the same bytes a K-major [K(samples), MN(features)] activation blob holds"""
    mn, K = x.shape
    return x.reshape(mn // 8, 8, K).permute(0, 2, 1).contiguous()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,K", [(256, 64), (128, 32), (256, 256), (32, 16), (128, 128)])
def test_umma_selftest(mode, N, K):
    from keras_nerf_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    blob = kmajor_blob if mode == 0 else mnmajor_blob
    a_d, b_d = blob(A).to(dev), blob(B).to(dev)
    out = torch.full((128, N), float("nan"), device=dev)
    _lib.call("knerf_selftest_umma", mode, a_d.data_ptr(), b_d.data_ptr(), N, K, _lib.ptr(out), _lib.stream())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = float((out.cpu() - ref).abs().max())
    assert err <= 1e-3 * max(1.0, float(ref.abs().max())), err
