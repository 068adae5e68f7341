"""Pins the tcgen05 conventions of sceneverse_b200/csrc/tc05.cuh (smem descriptor fields, canonical
K-major layout, instruction descriptor, TMEM lane/column mapping) against torch.matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(A, B, mode=0):
    from sceneverse_b200 import _lib
    lib = _lib.gps()
    N, K = B.shape
    D = torch.full((128, N), float("nan"), device="cuda")
    st = lib.sv_tc05_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, mode,
                              torch.cuda.current_stream().cuda_stream)
    _lib.check(lib, st, "sv_tc05_selftest")
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("N,K", [(64, 16), (64, 64), (128, 64), (128, 144), (256, 128), (16, 32), (48, 256)])
def test_single_tile_gemm(N, K):
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).cuda().to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).cuda().to(torch.bfloat16)
    want = A.float() @ B.float().t()
    got = run(A, B, 0)
    err = (got - want).abs().max().item()
    if not err < 1e-2:
        alt = run(A, B, 1)
        err1 = (alt - want).abs().max().item()
        pytest.fail(f"mode0 err={err} (mode1 err={err1}); got[0,:4]={got[0,:4].tolist()} want[0,:4]={want[0,:4].tolist()}")


def test_structured_operands_identify_layout():
    """A = one-hot rows, B = one-hot columns: any row/column permutation error shows up exactly."""
    K, N = 64, 64
    A = torch.zeros(128, K)
    A[torch.arange(128), torch.arange(128) % K] = 1.0
    B = torch.zeros(N, K)
    B[torch.arange(N), (torch.arange(N) * 7) % K] = torch.arange(1, N + 1).float()
    want = A @ B.t()
    got = run(A.cuda().to(torch.bfloat16), B.cuda().to(torch.bfloat16), 0)
    assert torch.equal(got.cpu(), want), f"nonzeros got={got.nonzero()[:8].tolist()} want={want.nonzero()[:8].tolist()}"
