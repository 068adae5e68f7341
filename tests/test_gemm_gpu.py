"""tcgen05 GEMM (sv_gemm_bf16) vs torch on the same bf16 operands (fp32 accumulate)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 64, 64), (256, 128, 128), (8320, 2304, 768), (5120, 768, 768), (3200, 30522, 768), (130, 72, 768),
          (77, 607, 384), (1, 768, 768), (81920, 256, 272), (640, 768, 2048), (333, 100, 8), (129, 257, 72)]


def ref(a, w, bias, act, residual):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == "relu":
        y = torch.relu(y)
    elif act == "gelu":
        y = torch.nn.functional.gelu(y)
    if residual is not None:
        y = y + residual.float()
    return y


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_plain(M, N, K):
    from sceneverse_b200 import native
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    got = native.gemm(a, w, out_dtype=torch.float32)
    want = ref(a, w, None, None, None)
    err = (got - want).abs().max().item() / (want.abs().max().item() + 1e-9)
    assert err < 2e-3, err


@pytest.mark.parametrize("act", [None, "relu", "gelu"])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_gemm_epilogues(act, out_dtype):
    from sceneverse_b200 import native
    M, N, K = 1000, 600, 136
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(out_dtype)
    got = native.gemm(a, w, bias, act, res, out_dtype=out_dtype).float()
    want = ref(a, w, bias, act, res)
    tol = 2e-2 if out_dtype == torch.bfloat16 else 2e-3
    assert (got - want).abs().max().item() / want.abs().max().item() < tol


def test_gemm_strided_rows_and_rowmax():
    from sceneverse_b200 import native
    g = torch.Generator(device="cuda").manual_seed(3)
    big = torch.randn(512, 400, device="cuda", generator=g).to(torch.bfloat16)
    a = big[:, 8:8 + 264]                      # row stride 400, K = 264
    w = (torch.randn(256, 264, device="cuda", generator=g) / 16).to(torch.bfloat16)
    bias = torch.randn(256, device="cuda", generator=g)
    got = native.gemm(a, w, bias, "relu", out_dtype=torch.float32, rowmax=16)
    want = torch.relu(a.float() @ w.float().t() + bias).view(32, 16, 256).max(dim=1).values
    assert got.shape == (32, 256)
    assert (got - want).abs().max().item() / want.abs().max().item() < 2e-3


@pytest.mark.parametrize("M,N,K,at,bt", [(128, 64, 64, 0, 1), (128, 64, 64, 1, 1), (128, 64, 64, 1, 0),
                                         (200, 136, 72, 0, 1), (200, 136, 72, 1, 1), (200, 136, 72, 1, 0),
                                         (384, 768, 3072, 0, 1), (384, 768, 3072, 1, 1), (384, 768, 3072, 1, 0),
                                         (777, 320, 1000, 0, 1), (64, 2048, 130, 1, 1)])
def test_gemm_transposed_operands(M, N, K, at, bt):
    """sv_gemm_bf16_ex: operands given transposed in memory are staged as 64-column slabs and read MN-major by the tensor
    core (the dgrad / wgrad forms of a linear layer) — against an fp32 torch matmul of the same bf16 values."""
    from sceneverse_b200 import native
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()
    A = rnd(K, M) if at else rnd(M, K)
    B = rnd(K, N) if bt else rnd(N, K)
    want = (A.float().t() if at else A.float()) @ (B.float() if bt else B.float().t())
    got = native.gemm_ex(A, B, a_transposed=bool(at), b_transposed=bool(bt), out_dtype=torch.float32)
    assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item()
