"""The GEMM kernels in isolation (poi_gemm_tn): fp32 FMA path (mode 0), tcgen05 3xTF32 (mode 1,
fp32-faithful) and tcgen05 1xTF32 (mode 2) against a float64 reference."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 32), (128, 64, 64), (256, 384, 256), (1000, 201, 128), (4096, 256, 128), (4096, 128, 204),
          (130, 72, 36), (33, 16, 32), (20000, 384, 256), (777, 130, 100),
          # >= 2 tiles per SM: the persistent kernels (mode 1: A operand in tensor memory), ragged M / N / K tails
          (126976, 384, 256), (40000, 130, 100), (50001, 204, 36), (37889, 128, 512)]


def _ref(A, W, b):
    return A.astype(np.float64) @ W.astype(np.float64).T + (b.astype(np.float64) if b is not None else 0.0)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("mode,tol", [(0, 1e-5), (1, 2e-5), (2, 5e-3)])
def test_gemm_tn(engine, M, N, K, mode, tol):
    rs = np.random.RandomState(M + N + K)
    A = rs.uniform(-0.5, 0.5, (M, K)).astype(np.float32)
    W = rs.uniform(-0.5, 0.5, (N, K)).astype(np.float32)
    b = rs.uniform(-0.5, 0.5, (N + 3,)).astype(np.float32)[:N].copy() if N % 4 == 0 else None
    C = engine.gemm_tn(torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda(),
                       torch.from_numpy(b).cuda() if b is not None else None, mode).cpu().numpy()
    ref = _ref(A, W, b)
    scale = np.sqrt(K) * 0.25 / 3.0 + 0.5          # typical magnitude of an entry
    err = np.max(np.abs(C - ref)) / scale
    assert err < tol, "mode %d max scaled error %.3e" % (mode, err)


@pytest.mark.parametrize("M,N1,N2", [(64, 128, 128), (1000, 384, 256), (126976, 384, 256), (5001, 204, 128), (40, 128, 128), (33333, 256, 128)])
@pytest.mark.parametrize("mode,tol", [(0, 1e-5), (1, 2e-5), (2, 5e-3)])
def test_gemm_atb(engine, M, N1, N2, mode, tol):
    """Weight-gradient contraction C = A^T B over the rows (MN-major tcgen05 operands in modes 1, 2)."""
    rs = np.random.RandomState(M + N1)
    A = rs.uniform(-0.5, 0.5, (M, N1)).astype(np.float32)
    B = rs.uniform(-0.5, 0.5, (M, N2)).astype(np.float32)
    C = engine.gemm_atb(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), mode).cpu().numpy()
    ref = A.astype(np.float64).T @ B.astype(np.float64)
    err = np.max(np.abs(C - ref)) / (np.sqrt(M) * 0.25 / 3.0 + 0.5)
    # tcgen05 adds into its fp32 accumulator with truncation: the error of a chain grows linearly with its length.
    # The engine caps chains at 2048 rows (gemm_tc.cuh:atb_tc_splits); the bound below is for those capped chains.
    if mode == 1 and M > 16384:
        tol = 1e-4
    assert err < tol, "mode %d scaled error %.3e" % (mode, err)
