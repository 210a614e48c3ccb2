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


# ----------------------------------------------------------------------------------------------
# fused HBM-bound operators vs their reference formulas in fp64 (the fixtures of the CPU suite pin those formulas
# to the reference's own modules)
# ----------------------------------------------------------------------------------------------
from fa_b200.mt import ops  # noqa: E402
from fa_b200.mt.optim import Adam  # noqa: E402
from fa_b200.mt.model import Parameter  # noqa: E402


def _grads(fn, inputs, gy):
    outs = fn(*inputs)
    (outs * gy).sum().backward()
    return outs, [t.grad for t in inputs if t.requires_grad]


@pytest.mark.parametrize("rows,dim", [(5632, 256), (77, 64), (3, 1024)])
def test_layernorm_kernels(cuda_device, rows, dim):
    torch.manual_seed(rows + dim)
    x = (torch.randn(rows, dim) * 2 + 0.5)
    w, b, gy = torch.rand(dim) + 0.5, torch.rand(dim) - 0.5, torch.rand(rows, dim) * 2 - 1
    ref_in = [t.double().requires_grad_(True) for t in (x, w, b)]
    y_ref, g_ref = _grads(lambda a, ww, bb: ops.layernorm(a, ww, bb, 1e-5), ref_in, gy.double())
    dev_in = [t.to(cuda_device).requires_grad_(True) for t in (x, w, b)]
    y, g = _grads(lambda a, ww, bb: ops.layernorm(a, ww, bb, 1e-5), dev_in, gy.to(cuda_device))
    assert float((y.double().cpu() - y_ref).abs().max()) <= 1e-5
    for got, ref in zip(g, g_ref):
        assert float((got.double().cpu() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


def test_gelu_dropout_embedding_kernels(cuda_device):
    torch.manual_seed(4)
    x, gy = torch.randn(1000, 256) * 2, torch.rand(1000, 256) * 2 - 1
    xr = x.double().requires_grad_(True)
    y_ref, (g_ref,) = _grads(ops.gelu, [xr], gy.double())
    xd = x.to(cuda_device).requires_grad_(True)
    y, (g,) = _grads(ops.gelu, [xd], gy.to(cuda_device))
    assert float((y.double().cpu() - y_ref).abs().max()) <= 2e-6 and float((g.double().cpu() - g_ref).abs().max()) <= 2e-6
    # dropout: inverted Bernoulli, same mask regenerated in the backward, fresh mask per call
    xd = torch.ones(2048, 512, device=cuda_device, requires_grad=True)
    y1 = ops.dropout(xd, 0.1, True)
    y1.sum().backward()
    keep = float((y1 != 0).float().mean())
    p_eff = 13 / 128
    assert abs(keep - (1 - p_eff)) <= 5 * (p_eff * (1 - p_eff) / y1.numel()) ** 0.5
    assert abs(float(y1.max()) - 1 / (1 - p_eff)) < 1e-6 and torch.equal(xd.grad, y1.detach())  # dy = 1: dx = mask * scale
    assert not torch.equal(ops.dropout(xd, 0.1, True), y1) and ops.dropout(xd, 0.1, False) is xd
    # embedding: gather forward, scatter-add backward with repeated indices
    w = torch.randn(300, 64)
    idx = torch.randint(0, 300, (7, 33))
    ge = torch.randn(7, 33, 64)
    wr = w.double().requires_grad_(True)
    (wr[idx] * ge.double()).sum().backward()
    wd = w.to(cuda_device).requires_grad_(True)
    yd = ops.embedding(idx.to(cuda_device), wd)
    (yd * ge.to(cuda_device)).sum().backward()
    assert torch.equal(yd.cpu(), w[idx]) and float((wd.grad.double().cpu() - wr.grad).abs().max()) <= 1e-5


@pytest.mark.parametrize("rows,C", [(5632, 10000), (9, 36), (64, 300)])
def test_softmax_loss_kernels(cuda_device, rows, C):
    torch.manual_seed(C)
    logits, target, gw = torch.randn(rows, C) * 3, torch.randint(0, C, (rows,)), torch.rand(rows)
    lr = logits.double().requires_grad_(True)
    loss_ref = ops.softmax_loss(lr, target)
    (loss_ref * gw.double()).sum().backward()
    ld = logits.to(cuda_device).requires_grad_(True)
    loss = ops.softmax_loss(ld, target.to(cuda_device))
    (loss * gw.to(cuda_device)).sum().backward()
    assert float((loss.double().cpu() - loss_ref).abs().max()) <= 2e-5
    assert float((ld.grad.double().cpu() - lr.grad).abs().max()) <= 2e-6


def test_fused_adam_kernel_matches_reference_formula(cuda_device):
    torch.manual_seed(1)
    p0, grads = torch.rand(1000) * 2 - 1, torch.randn(4, 1000)
    ref = Parameter(p0.double().clone())
    oref = Adam([ref], lr=0.02)
    flat, gflat = p0.to(cuda_device).clone(), torch.zeros(1000, device=cuda_device)
    odev = Adam([], lr=0.02, flat=(flat, gflat))
    for g in grads:
        ref.value.grad = g.double()
        oref.step()
        gflat.copy_(g)
        odev.step()
    assert float((flat.double().cpu() - ref.value.detach()).abs().max()) <= 2e-6


@pytest.mark.parametrize("B,T,H,d,K", [(2, 128, 2, 32, 64), (3, 256, 8, 32, 256), (1, 128, 2, 64, 128), (2, 128, 1, 128, 64)])
def test_head_layout_projections_equal_linear_plus_permute(cuda_device, B, T, H, d, K):
    """linear_to_heads / linear_from_heads (the GEMM's tensor maps write / read (B,H,T,d) directly) against the
    reference sequence linear -> view -> permute -> contiguous on the same kernels: same products, same k order, so
    outputs and input gradients agree to fp32 rounding; weight gradients (split-K reductions) to 1e-5 relative."""
    torch.manual_seed(B * T + d)
    E = H * d
    x = torch.randn(B * T, K, device=cuda_device)
    w, b = torch.randn(K, E, device=cuda_device) / K ** 0.5, torch.randn(E, device=cuda_device)
    wo, bo = torch.randn(E, K, device=cuda_device) / E ** 0.5, torch.randn(K, device=cuda_device)
    gy = torch.randn(B * T, K, device=cuda_device)

    def run(fused):
        leaves = [t.clone().requires_grad_(True) for t in (x, w, b, wo, bo)]
        xx, ww, bb, wwo, bbo = leaves
        if fused:
            assert ops.heads_layout_ok(xx, T, d)
            heads = ops.linear_to_heads(xx, ww, bb, B, T, H, d)
            y = ops.linear_from_heads(heads * 1.5, wwo, bbo)
        else:
            heads = ops.linear(xx, ww, bb).view(B, T, H, d).permute(0, 2, 1, 3).contiguous()
            y = ops.linear((heads * 1.5).permute(0, 2, 1, 3).contiguous().view(B * T, E), wwo, bbo)
        (y * gy).sum().backward()
        return [heads.detach(), y.detach()] + [t.grad for t in leaves]

    before = _lib.launch_count()
    fused = run(True)
    assert _lib.launch_count() - before >= 8  # 2 forward GEMMs, 4 backward GEMMs, 2 bias reductions: all ours
    plain = run(False)
    for name, a, c in zip(("heads", "y", "dx", "dw", "db", "dwo", "dbo"), fused, plain):
        assert a.shape == c.shape and float((a - c).abs().max()) <= 1e-5 * max(1.0, float(c.abs().max())), name
    # and against fp64 on the CPU within the tf32 bar
    ref = (x.double().cpu() @ w.double().cpu() + b.double().cpu()).view(B, T, H, d).permute(0, 2, 1, 3)
    assert float((fused[0].double().cpu() - ref).abs().max()) <= 3e-3 * K ** 0.5


@pytest.mark.parametrize("B,T,H,d,K", [(2, 128, 2, 32, 64), (3, 256, 8, 32, 256), (1, 128, 2, 64, 128)])
def test_fused_qkv_projection_equals_three_projections(cuda_device, B, T, H, d, K):
    """ops.qkv_to_heads (one GEMM over the interleaved weight [in][Wq | Wk | Wv], output (3, B, H, T, d)) against three
    ops.linear_to_heads on the same kernels: q, k, v and the input gradient agree to fp32 rounding (same products, the
    input gradient is summed inside one K = 3 E reduction instead of by two adds), weight / bias gradients to 1e-5 relative."""
    torch.manual_seed(7 * B + T + d)
    E = H * d
    x = torch.randn(B * T, K, device=cuda_device)
    ws = [torch.randn(K, E, device=cuda_device) / K ** 0.5 for _ in range(3)]
    bs = [torch.randn(E, device=cuda_device) for _ in range(3)]
    gy = torch.randn(3, B, H, T, d, device=cuda_device)

    xx = x.clone().requires_grad_(True)
    wqkv = torch.cat(ws, dim=1).contiguous().requires_grad_(True)
    bqkv = torch.cat(bs).contiguous().requires_grad_(True)
    before = _lib.launch_count()
    y = ops.qkv_to_heads(xx, wqkv, bqkv, B, T, H, d)
    (y * gy).sum().backward()
    assert _lib.launch_count() - before == 4  # forward GEMM, dX GEMM, dW GEMM, bias reduction

    x3 = x.clone().requires_grad_(True)
    w3 = [w.clone().requires_grad_(True) for w in ws]
    b3 = [b.clone().requires_grad_(True) for b in bs]
    ys = [ops.linear_to_heads(x3, w3[i], b3[i], B, T, H, d) for i in range(3)]
    sum((ys[i] * gy[i]).sum() for i in range(3)).backward()

    tol = lambda c: 1e-5 * max(1.0, float(c.abs().max()))  # noqa: E731
    for i in range(3):
        assert float((y[i] - ys[i]).abs().max()) <= tol(ys[i])
        assert float((wqkv.grad[:, i * E:(i + 1) * E] - w3[i].grad).abs().max()) <= tol(w3[i].grad)
        assert float((bqkv.grad[i * E:(i + 1) * E] - b3[i].grad).abs().max()) <= tol(b3[i].grad)
    assert float((xx.grad - x3.grad).abs().max()) <= 1e-4 * max(1.0, float(x3.grad.abs().max()))


@pytest.mark.parametrize("rows,dim", [(5632, 256), (77, 128), (9, 512), (1, 256)])
@pytest.mark.parametrize("with_b", [True, False])
def test_add_layernorm_fused_residual(cuda_device, rows, dim, with_b):
    """ops.add_layernorm: (a + b, LN(a + b)) in one pass; its backward adds the gradient arriving at the sum along the
    residual branch inside the LayerNorm-backward kernel (one launch, dw / db included).  Against the torch formula in
    float64, with BOTH outputs used downstream (the transformer layer's pattern)."""
    torch.manual_seed(rows + dim)
    a, b = torch.randn(rows, dim) * 2 + 0.5, torch.randn(rows, dim)
    w, bias = torch.rand(dim) + 0.5, torch.rand(dim) - 0.5
    gs, gy = torch.rand(rows, dim) * 2 - 1, torch.rand(rows, dim) * 2 - 1

    def run(dev, dt):
        ins = [t.to(device=dev, dtype=dt).requires_grad_(True) for t in ((a, b, w, bias) if with_b else (a, w, bias))]
        s, y = ops.add_layernorm(ins[0], ins[1] if with_b else None, ins[-2], ins[-1], 1e-5)
        ((s * gs.to(device=dev, dtype=dt)).sum() + (y * gy.to(device=dev, dtype=dt)).sum()).backward()
        return s, y, [t.grad for t in ins]

    s_ref, y_ref, g_ref = run("cpu", torch.float64)
    before = _lib.launch_count()
    s, y, g = run(cuda_device, torch.float32)
    assert _lib.launch_count() - before == 2  # one forward, one backward
    assert float((s.double().cpu() - s_ref).abs().max()) <= 1e-6 * max(1.0, float(s_ref.abs().max()))
    assert float((y.double().cpu() - y_ref).abs().max()) <= 1e-5
    for got, ref in zip(g, g_ref):
        assert float((got.double().cpu() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    # an unused sum costs nothing (no zero tensor is materialised for it) and gives the plain LayerNorm gradient
    x1 = a.to(cuda_device).requires_grad_(True)
    _, y1 = ops.add_layernorm(x1, None, w.to(cuda_device), bias.to(cuda_device), 1e-5)
    (y1 * gy.to(cuda_device)).sum().backward()
    x2 = a.to(cuda_device).requires_grad_(True)
    (ops.layernorm(x2, w.to(cuda_device), bias.to(cuda_device), 1e-5) * gy.to(cuda_device)).sum().backward()
    assert torch.equal(x1.grad, x2.grad)


@pytest.mark.parametrize("parts,B,H,T,d", [(3, 11, 8, 512, 32), (1, 2, 2, 128, 64), (3, 1, 3, 40, 24), (2, 3, 1, 7, 128)])
def test_colsum_heads_parts(cuda_device, parts, B, H, T, d):
    """Bias gradient of a projection whose output is in head layout: out[part][h*d + dd] = sum over (b, t)."""
    torch.manual_seed(parts + B + T)
    x = torch.randn(parts, B, H, T, d, device=cuda_device)
    out = torch.full((parts, H * d), 3.0, device=cuda_device)
    ref = x.double().sum(dim=(1, 3)).reshape(parts, H * d)
    lib = _lib.load_library()
    _lib.check(lib.fa_b200_colsum_heads_parts_f32(x.data_ptr(), out.data_ptr(), parts, B, H, T, d, 0, None))
    torch.cuda.synchronize()
    assert float((out.double() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    _lib.check(lib.fa_b200_colsum_heads_parts_f32(x.data_ptr(), out.data_ptr(), parts, B, H, T, d, 1, None))
    torch.cuda.synchronize()
    assert float((out.double() - 2 * ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


def test_packed_attention_equals_unpacked(cuda_device):
    """ops.flash_attention.packed((3, B, H, T, d)) is flash_attention(q, k, v) with dq | dk | dv written into one tensor."""
    torch.manual_seed(3)
    qkv = torch.randn(3, 2, 4, 256, 32, device=cuda_device)
    go = torch.randn(2, 4, 256, 32, device=cuda_device)
    p = qkv.clone().requires_grad_(True)
    o1 = ops.flash_attention.packed(p, True)
    (o1 * go).sum().backward()
    u = qkv.clone().requires_grad_(True)
    o2 = ops.flash_attention(u[0], u[1], u[2], True)
    (o2 * go).sum().backward()
    assert torch.equal(o1, o2) and torch.equal(p.grad, u.grad) and p.grad.is_contiguous()


def test_layernorm_backward_also_reduces_the_producer_bias_gradient(cuda_device):
    """linear(..., bias_grad=False) followed by add_layernorm(..., b_bias=bias): the bias gradient (column sums of the
    gradient of b) comes out of the LayerNorm-backward kernel; same numbers as the linear layer's own reduction."""
    torch.manual_seed(12)
    rows, K, dim = 640, 64, 256
    x, a = torch.randn(rows, K, device=cuda_device), torch.randn(rows, dim, device=cuda_device)
    W, bl = torch.randn(K, dim, device=cuda_device) / 8, torch.randn(dim, device=cuda_device)
    w, bias = torch.rand(dim, device=cuda_device) + 0.5, torch.rand(dim, device=cuda_device) - 0.5
    gs, gy = torch.randn(rows, dim, device=cuda_device), torch.randn(rows, dim, device=cuda_device)

    def run(fused):
        ins = [t.clone().requires_grad_(True) for t in (x, a, W, bl, w, bias)]
        xx, aa, WW, bb, ww, bi = ins
        assert ops.fused_bias_grad_ok(xx, rows, WW, bb)
        before = _lib.launch_count()
        b = ops.linear(xx, WW, bb, bias_grad=not fused)
        s, y = ops.add_layernorm(aa, b, ww, bi, 1e-5, bb if fused else None)
        ((s * gs).sum() + (y * gy).sum()).backward()
        return _lib.launch_count() - before, [t.grad for t in ins]

    n_plain, g_plain = run(False)
    n_fused, g_fused = run(True)
    assert n_plain - n_fused == 1  # the separate bias reduction is gone
    for gp, gf in zip(g_plain, g_fused):
        assert float((gp - gf).abs().max()) <= 1e-5 * max(1.0, float(gp.abs().max()))
    ref = g_plain[1].double().sum(0)  # gradient of b == gradient of a: its column sums
    assert float((g_fused[3].double() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))


def test_gelu_dropout_equals_gelu_then_dropout(cuda_device):
    torch.manual_seed(13)
    x, gy = torch.randn(512, 256, device=cuda_device) * 2, torch.randn(512, 256, device=cuda_device)
    st = ops.dropout_state
    calls, step = st.calls, st.step
    a = x.clone().requires_grad_(True)
    y1 = ops.gelu_dropout(a, 0.1, True)
    (y1 * gy).sum().backward()
    st.calls = calls  # replay the same call site: same mask
    b = x.clone().requires_grad_(True)
    y2 = ops.dropout(ops.gelu(b), 0.1, True)
    (y2 * gy).sum().backward()
    assert torch.equal(y1, y2) and float((y1 == 0).float().mean()) > 0.05
    assert float((a.grad - b.grad).abs().max()) <= 1e-6 * max(1.0, float(b.grad.abs().max()))
    assert torch.equal(ops.gelu_dropout(a, 0.1, False), ops.gelu(a))  # eval mode: plain gelu


@pytest.mark.parametrize("rows,C", [(5632, 10000), (12, 36), (300, 12288), (40, 20000)])
def test_softmax_loss_backward_also_reduces_the_logits_bias_gradient(cuda_device, rows, C):
    torch.manual_seed(rows)
    K = 32
    x = torch.randn(rows, K, device=cuda_device)
    W, b = torch.randn(K, C, device=cuda_device) / 4, torch.randn(C, device=cuda_device)
    target = torch.randint(0, C, (rows,), device=cuda_device)
    wts = torch.rand(rows, device=cuda_device)

    def run(fused):
        ins = [t.clone().requires_grad_(True) for t in (x, W, b)]
        assert ops.fused_logits_bias_grad_ok(ins[0], rows, ins[1], ins[2], target)
        before = _lib.launch_count()
        logits = ops.linear(ins[0], ins[1], ins[2], bias_grad=not fused)
        loss = ops.softmax_loss(logits, target, ins[2] if fused else None)
        (loss * wts).mean().backward()
        return _lib.launch_count() - before, loss, [t.grad for t in ins]

    n_plain, l_plain, g_plain = run(False)
    n_fused, l_fused, g_fused = run(True)
    assert n_plain - n_fused == 1 and torch.equal(l_plain, l_fused)
    for gp, gf in zip(g_plain, g_fused):
        assert float((gp - gf).abs().max()) <= 1e-5 * max(1e-3, float(gp.abs().max()))


@pytest.mark.parametrize("training", [True, False])
def test_embed_pos_dropout_equals_the_three_operators(cuda_device, training):
    """ops.embed_pos_dropout = dropout(embedding(idx, tok) + embedding(arange(T), pos)): same values (same mask at the
    same call site), same table gradients; rows of the position table beyond T get no gradient."""
    torch.manual_seed(21)
    B, T, V, dim = 5, 64, 300, 128
    idx = torch.randint(0, V, (B, T), device=cuda_device)
    tok, pos = torch.randn(V, dim, device=cuda_device), torch.randn(V, dim, device=cuda_device)
    gy = torch.randn(B, T, dim, device=cuda_device)
    st = ops.dropout_state
    calls = st.calls
    t1, p1 = tok.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    before = _lib.launch_count()
    y1 = ops.embed_pos_dropout(idx, t1, p1, 0.1, training)
    (y1 * gy).sum().backward()
    assert _lib.launch_count() - before == 2
    st.calls = calls
    t2, p2 = tok.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    ar = torch.arange(T, device=cuda_device).view(1, T)
    y2 = ops.dropout(ops.embedding(idx, t2) + ops.embedding(ar, p2), 0.1, training)
    (y2 * gy).sum().backward()
    assert torch.equal(y1, y2)
    assert training == bool((y1 == 0).any())
    for a, b in ((t1.grad, t2.grad), (p1.grad, p2.grad)):
        assert float((a - b).abs().max()) <= 1e-5 * max(1.0, float(b.abs().max()))
    assert float(p1.grad[T:].abs().max()) == 0.0


@pytest.mark.parametrize("training", [True, False])
def test_gelu_dropout_backward_also_reduces_the_producer_bias_gradient(cuda_device, training):
    torch.manual_seed(31)
    rows, K, dim = 648, 64, 256
    x, W, b = torch.randn(rows, K, device=cuda_device), torch.randn(K, dim, device=cuda_device) / 8, torch.randn(dim, device=cuda_device)
    gy = torch.randn(rows, dim, device=cuda_device)
    st = ops.dropout_state

    def run(fused):
        ins = [t.clone().requires_grad_(True) for t in (x, W, b)]
        calls = st.calls
        before = _lib.launch_count()
        h = ops.linear(ins[0], ins[1], ins[2], bias_grad=not fused)
        y = ops.gelu_dropout(h, 0.1, training, x_bias=ins[2] if fused else None)
        (y * gy).sum().backward()
        st.calls = calls  # the same call site for both runs: same mask
        return _lib.launch_count() - before, y, [t.grad for t in ins]

    n_plain, y_plain, g_plain = run(False)
    n_fused, y_fused, g_fused = run(True)
    assert n_plain - n_fused == 1 and torch.equal(y_plain, y_fused)
    for gp, gf in zip(g_plain, g_fused):
        assert float((gp - gf).abs().max()) <= 1e-5 * max(1.0, float(gp.abs().max()))
