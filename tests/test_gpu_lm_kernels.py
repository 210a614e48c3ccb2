"""Hand-written kernels of the decoder-LM slice on the B200, against fp64 torch
restatements of the same operators (floating-point kernels: tolerance stated per test)."""
import ctypes

import pytest
import torch

from fa_b200 import _lib

pytestmark = pytest.mark.gpu


def gemm(a, b, bias, M, N, K, a_mn, b_mn, c=None, accumulate=False):
    lib = _lib.load_library()
    if c is None:
        c = torch.empty(M, N, device=a.device, dtype=torch.float32)
    _lib.check(lib.fa_b200_gemm_tf32(a.data_ptr(), b.data_ptr(), c.data_ptr(), bias.data_ptr() if bias is not None else None,
                                     M, N, K, int(a_mn), int(b_mn), int(accumulate), torch.cuda.current_stream().cuda_stream))
    return c


@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (0, 0), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 384, 256), (132, 200, 100), (5632, 256, 256), (64, 10000, 256), (256, 256, 5632)])
def test_gemm_tf32_all_operand_layouts(cuda_device, M, N, K, a_mn, b_mn):
    """The tensor core truncates fp32 operands to tf32 (10 mantissa bits) and accumulates in fp32: against the fp64
    product of the TRUNCATED operands the kernel must be exact up to fp32 accumulation (max(1e-5 sqrt(K), 5e-7 K)
    absolute for U(-1,1) data); against the untruncated product the error is the tf32 rounding itself (<= 3e-3 * sqrt(K))."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = (torch.rand(M, K, generator=g, dtype=torch.float64) * 2 - 1).float()
    B = (torch.rand(K, N, generator=g, dtype=torch.float64) * 2 - 1).float()
    bias = torch.rand(N, generator=g, dtype=torch.float64).float()
    trunc = lambda t: (t.view(torch.int32) & -8192).view(torch.float32).double()  # noqa: E731  keep sign, exponent, 10 bits
    ref_t = trunc(A) @ trunc(B) + bias.double()
    ref = A.double() @ B.double() + bias.double()
    a = (A.t().contiguous() if a_mn else A.contiguous()).to(cuda_device)   # a_mn: stored [K][M]
    b = (B.contiguous() if b_mn else B.t().contiguous()).to(cuda_device)   # b_mn: stored [K][N]
    c = gemm(a, b, bias.to(cuda_device), M, N, K, a_mn, b_mn)
    got = c.double().cpu()
    acc_tol = max(1e-5 * K ** 0.5, 5e-7 * K) + 1e-5  # fp32 accumulation inside the tensor core (truncating adds)
    assert float((got - ref_t).abs().max()) <= acc_tol
    assert float((got - ref).abs().max()) <= 3e-3 * K ** 0.5
    # accumulate on top of an existing C, no bias
    c2 = gemm(a, b, None, M, N, K, a_mn, b_mn, c=c.clone(), accumulate=True)
    assert float((c2.double().cpu() - (2 * ref_t - bias.double())).abs().max()) <= 2 * acc_tol


def test_gemm_rejects_unaligned_sizes(cuda_device):
    lib = _lib.load_library()
    x = torch.zeros(16, 16, device=cuda_device)
    assert lib.fa_b200_gemm_tf32(x.data_ptr(), x.data_ptr(), x.data_ptr(), None, 16, 16, 6, 0, 1, 0, None) == _lib.FA_ERR_BAD_ARG


def test_colsum(cuda_device):
    lib = _lib.load_library()
    for M, N in ((5632, 10000), (77, 256), (4096, 32)):
        x = torch.randn(M, N, device=cuda_device)
        out = torch.full((N,), 3.0, device=cuda_device)
        _lib.check(lib.fa_b200_colsum_f32(x.data_ptr(), out.data_ptr(), M, N, 0, torch.cuda.current_stream().cuda_stream))
        ref = x.double().sum(0)
        assert float((out.double() - ref).abs().max()) <= 1e-3 * float(ref.abs().max())
        _lib.check(lib.fa_b200_colsum_f32(x.data_ptr(), out.data_ptr(), M, N, 1, torch.cuda.current_stream().cuda_stream))
        assert float((out.double() - 2 * ref).abs().max()) <= 2e-3 * float(ref.abs().max())
