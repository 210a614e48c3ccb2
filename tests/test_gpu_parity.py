"""Parity tests proper (B200, `-m gpu`): every call goes through the C ABI of libfa_b200.so.

Bars (BASELINE.md §5 / north_star):
  * fp32 path  : max-abs <= 1e-3 on O, dQ, dK, dV against the reference's vanilla attention (golden fixtures
                 produced by the reference itself + the pinned oracle)
  * bf16 path  : ||x - ref||_inf / ||ref||_inf <= 2e-2 per tensor against the fp64 oracle evaluated on the
                 same bf16-rounded inputs
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_files
from fa_b200 import _lib, device
from fa_b200 import minitorch_compat as mt
from oracle import attention_oracle as ao
from oracle import c_oracle

pytestmark = pytest.mark.gpu

FP32_ATOL = 1e-3   # north_star: max-abs 1e-3 on the fp32/TF32 path
BF16_REL = 2e-2    # north_star: relative 2e-2 on the bf16 path


def rel_inf(x, ref):
    """||x - ref||_inf / ||ref||_inf; a reference that is identically zero (e.g. dQ for T = 1) is compared with a
    1e-6 absolute floor so that fp32 rounding noise does not read as an infinite relative error."""
    return float(np.abs(np.asarray(x, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-4))


def _mt4(a, be, requires_grad=False):
    return mt.tensor_from_numpy(np.ascontiguousarray(a[None]), backend=be, requires_grad=requires_grad)


# ----------------------------------------------------------------------------------------------
# 1. the reference operator surface, on fixtures computed by the reference itself
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["flash_attn", "flash_attn_causal", "flash_attn2"])
@pytest.mark.parametrize("name", golden_files("mt_"))
def test_tensor_flash_attn_matches_reference_vanilla(cuda_device, name, method):
    """Reads like kernel_tests/test_flashattn_{fw,bw}.py: custom = q.flash_attn*(k, v, causal); baseline =
    vanilla minitorch attention (here: its recorded output)."""
    g = np.load(os.path.join(GOLDEN, name))
    causal = bool(g["causal"])
    be = mt.TensorBackend(mt.CudaKernelOps)
    q, k, v = (_mt4(g[x], be, True) for x in ("q", "k", "v"))
    out = getattr(q, method)(k, v, causal_mask=causal)
    assert out.shape == (1,) + g["q"].shape
    assert np.abs(out.to_numpy()[0] - g["o"]).max() <= FP32_ATOL
    out.backward(_mt4(g["d_out"], be))
    for t, key in ((q, "dq"), (k, "dk"), (v, "dv")):
        assert np.abs(t.grad.to_numpy()[0] - g[key]).max() <= FP32_ATOL, key


@pytest.mark.parametrize("precision", ["auto", "fp32"])
@pytest.mark.parametrize("name", golden_files("fap_"))
def test_saved_statistics_follow_each_family_convention(cuda_device, name, precision):
    """FA-1 libraries return l = sum exp(s - m), m = rowmax; FA-2 returns L in `l` and leaves `m` as passed
    (-FLT_MAX) -- pinned by the authors' flash_attn_python.py outputs.  `auto` is what a drop-in user gets (tf32 tensor
    cores where a tcgen05 kernel covers the shape: everything within the 1e-3 bar of the fp32 path); `fp32` is the exact
    CUDA-core arithmetic (statistics to 1e-5)."""
    lib = _lib.load_library()
    prev = lib.fa_b200_get_host_precision()
    lib.fa_b200_set_host_precision(_lib.FA_PREC_AUTO if precision == "auto" else _lib.FA_PREC_FP32_SIMT)
    stat_tol = 1e-3 if precision == "auto" else 1e-5
    try:
        g = np.load(os.path.join(GOLDEN, name))
        be = mt.TensorBackend(mt.CudaKernelOps)
        q, k, v = (_mt4(g[x].astype(np.float32), be) for x in ("q", "k", "v"))
        o1, l1, m1 = mt.CudaKernelOps.flash_attn_fw(q, k, v, mt.tensor(False))
        assert np.abs(o1.to_numpy()[0] - g["o_fa1"]).max() <= FP32_ATOL
        assert np.abs(m1.to_numpy()[0] - g["m_fa1"]).max() <= stat_tol
        assert rel_inf(l1.to_numpy()[0], g["l_fa1"]) <= 2 * stat_tol
        o2, l2, m2 = mt.CudaKernelOps.flash_attn2_fw(q, k, v, mt.tensor(False))
        assert np.abs(l2.to_numpy()[0] - g["L_fa2"]).max() <= 10 * stat_tol
        assert np.all(m2.to_numpy() == -np.finfo(np.float32).max)
        # each backward consumes its own convention
        dO = _mt4(g["d_out"].astype(np.float32), be)
        for fam, (o, l, m) in (("fa1", (o1, l1, m1)), ("fa2", (o2, l2, m2))):
            bw = mt.CudaKernelOps.flash_attn_bw if fam == "fa1" else mt.CudaKernelOps.flash_attn2_bw
            dq, dk, dv, _ = bw(q, k, v, o, dO, l, m, mt.tensor(False))
            for got, key in ((dq, "dq_"), (dk, "dk_"), (dv, "dv_")):
                assert np.abs(got.to_numpy()[0] - g[key + fam]).max() <= FP32_ATOL
    finally:
        lib.fa_b200_set_host_precision(prev)


def test_backward_accumulates_into_q_grad_like_reference(cuda_device):
    """src/flash_attn2_bw.cu:228 atomically adds into the q_grad the caller passed; k_grad / v_grad are overwritten."""
    lib = _lib.load_library()
    BH, T, d = 2, 48, 32
    q, k, v, g = ao.make_inputs(BH, T, d, seed=5)
    dq_ref, dk_ref, dv_ref, o_ref, L = ao.attention_backward(q, k, v, g, True, np.float64)
    o = o_ref.astype(np.float32)
    l = L.astype(np.float32)
    m = np.full((BH, T), -np.finfo(np.float32).max, np.float32)
    qg = np.full_like(q, 0.25)
    kg = np.full_like(q, 7.0)
    vg = np.full_like(q, 7.0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    _lib.check(lib.fa_b200_host_bwd(p(q), p(k), p(v), p(o), p(g), p(qg), p(kg), p(vg), p(l), p(m), BH, T, d, 1,
                                    _lib.FA_CONV_FA2, None))
    assert np.abs(qg - (dq_ref + 0.25)).max() <= FP32_ATOL
    assert np.abs(kg - dk_ref).max() <= FP32_ATOL and np.abs(vg - dv_ref).max() <= FP32_ATOL


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_host_api_chunk_pipeline_equals_single_shot(cuda_device, monkeypatch, prec):
    """The host-pointer ABI streams heads through a 3-buffer H2D / kernel / D2H pipeline; forcing 1-head chunks (7
    chunks over 3 buffers, ragged last use) must give bit-identical results to the single-chunk call."""
    lib = _lib.load_library()
    BH, T, d = 7, 200, 64
    q, k, v, g = ao.make_inputs(BH, T, d, seed=21)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    prev = lib.fa_b200_get_host_precision()
    lib.fa_b200_set_host_precision(_lib.FA_PREC_BF16_TC if prec == "bf16" else _lib.FA_PREC_FP32_SIMT)
    try:
        res = []
        for chunk in (None, str(T * d)):
            if chunk is None:
                monkeypatch.delenv("FA_B200_HOST_CHUNK_ELEMS", raising=False)
            else:
                monkeypatch.setenv("FA_B200_HOST_CHUNK_ELEMS", chunk)
            out = np.zeros_like(q)
            l = np.zeros((BH, T), np.float32)
            m = np.full((BH, T), -np.finfo(np.float32).max, np.float32)
            _lib.check(lib.fa_b200_host_fwd(p(q), p(k), p(v), p(out), p(l), p(m), BH, T, d, 1, _lib.FA_CONV_FA1, None))
            qg, kg, vg = np.full_like(q, 0.5), np.zeros_like(q), np.zeros_like(q)
            _lib.check(lib.fa_b200_host_bwd(p(q), p(k), p(v), p(out), p(g), p(qg), p(kg), p(vg), p(l), p(m), BH, T, d, 1,
                                            _lib.FA_CONV_FA1, None))
            res.append((out, l, m, kg, vg, qg))
        for a, b in zip(res[0][:5], res[1][:5]):
            assert np.array_equal(a, b)
        # dQ goes through fp32 reductions whose order is not fixed on the tensor-core path
        assert np.abs(res[0][5] - res[1][5]).max() <= (0 if prec == "fp32" else 1e-5)
        dq_r, dk_r, dv_r, o_r, _ = ao.attention_backward(q, k, v, g, True, np.float64)
        tol = FP32_ATOL if prec == "fp32" else 2e-2
        assert np.abs(res[1][0] - o_r).max() <= tol and np.abs(res[1][5] - (dq_r + 0.5)).max() <= tol
    finally:
        lib.fa_b200_set_host_precision(prev)


# ----------------------------------------------------------------------------------------------
# 2. device-resident API, bf16 tensor-core path vs the oracle on the same bf16-rounded inputs
# ----------------------------------------------------------------------------------------------
def _bf16_problem(BH, T, d, seed, scale=1.0):
    q, k, v, g = ao.make_inputs(BH, T, d, seed)
    return tuple(ao.bf16_round(x * s) for x, s in ((q, scale), (k, scale), (v, 1.0), (g, 1.0)))


def _to_dev(a, dev, bf16=True):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.to(torch.bfloat16) if bf16 else t


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("BH,T,d", [(3, 128, 128), (2, 256, 64), (2, 1000, 128), (1, 777, 64), (2, 1, 64), (2, 129, 128),
                                    (1, 2048, 128)])
def test_device_bf16_forward_backward_vs_oracle(cuda_device, BH, T, d, causal):
    import torch
    q, k, v, g = _bf16_problem(BH, T, d, seed=100 + T + d)
    dq_r, dk_r, dv_r, o_r, L_r = ao.attention_backward(q, k, v, g, causal, np.float64)
    tq, tk, tv, tg = (_to_dev(x, cuda_device) for x in (q, k, v, g))
    o, lse, m, l = device.flash_attn_fwd(tq, tk, tv, causal, want_lm=True)
    assert o.dtype == torch.bfloat16
    assert rel_inf(o.float().cpu().numpy(), o_r) <= BF16_REL
    assert np.abs(lse.cpu().numpy() - L_r).max() <= 2e-3
    assert np.abs((m + torch.log(l)).cpu().numpy() - L_r).max() <= 2e-3
    dq, dk, dv = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal)
    for got, ref, name in ((dq, dq_r, "dq"), (dk, dk_r, "dk"), (dv, dv_r, "dv")):
        assert rel_inf(got.float().cpu().numpy(), ref) <= BF16_REL, name


@pytest.fixture
def pair_forward():
    """Switch the d = 128 bf16 forward to the CTA-pair kernel (fa_fwd_pair_sm100.cu) for one test."""
    lib = _lib.load_library()
    prev = lib.fa_b200_get_fwd_pair()
    lib.fa_b200_set_fwd_pair(1)
    yield
    lib.fa_b200_set_fwd_pair(prev)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("BH,T", [(3, 128), (2, 1), (2, 129), (2, 512), (1, 513), (2, 1000), (1, 2048), (1, 2125)])
def test_pair_forward_kernel_vs_oracle(cuda_device, pair_forward, BH, T, causal):
    """The cta_group::2 forward (opt-in: measured slower than the single-CTA kernel) computes the same O, L
    and FA-1 statistics: one cluster = 512 query rows, T below / across / far beyond one cluster, ragged tails, the step
    that only the second CTA of a pair needs (rank 0 rides along with P = 0)."""
    q, k, v, g = _bf16_problem(BH, T, 128, seed=300 + T)
    o_r, p, m_r, l_r, L_r = ao.attention_forward(q, k, v, causal, np.float64)
    o, lse, m, l = device.flash_attn_fwd(*(_to_dev(x, cuda_device) for x in (q, k, v)), causal, want_lm=True)
    assert rel_inf(o.float().cpu().numpy(), o_r) <= BF16_REL
    assert np.abs(lse.cpu().numpy() - L_r).max() <= 2e-3
    import torch
    assert np.abs((m + torch.log(l)).cpu().numpy() - L_r).max() <= 2e-3


def test_pair_forward_kernel_lazy_rescale_and_dropout(cuda_device, pair_forward):
    q, k, v, g = _bf16_problem(2, 1024, 128, seed=77, scale=6.0)
    for causal in (False, True):
        o_r, p, m_r, l_r, L_r = ao.attention_forward(q, k, v, causal, np.float64)
        o, lse = device.flash_attn_fwd(*(_to_dev(x, cuda_device) for x in (q, k, v)), causal)
        assert rel_inf(o.float().cpu().numpy(), o_r) <= BF16_REL
        assert np.abs(lse.cpu().numpy() - L_r).max() <= 1e-2
    pd, seed, BH, T = 0.1, 4242, 2, 640
    q, k, v, g = _bf16_problem(BH, T, 128, seed=31)
    keep = device.dropout_mask(BH, T, pd, seed, device=cuda_device).cpu().numpy()
    _, _, _, o_r, L_r = ao.attention_dropout_backward(q, k, v, g, True, keep, ao.dropout_keep_scale(pd))
    o, lse = device.flash_attn_fwd(*(_to_dev(x, cuda_device) for x in (q, k, v)), True, p_drop=pd, seed=seed)
    assert rel_inf(o.float().cpu().numpy(), o_r) <= BF16_REL and np.abs(lse.cpu().numpy() - L_r).max() <= 2e-3


@pytest.mark.parametrize("BH,T,d,causal,p_drop,bf16", [(40, 1024, 128, True, 0.0, True), (40, 1024, 128, False, 0.0, True),
                                                      (5, 2125, 128, True, 0.0, True), (400, 200, 128, True, 0.0, True),
                                                      (37, 700, 64, True, 0.0, True), (21, 1100, 128, True, 0.1, True),
                                                      (3, 4224, 128, True, 0.0, True), (30, 1000, 64, False, 0.0, False),
                                                      (50, 600, 32, True, 0.0, False)])
def test_persistent_forward_equals_one_cta_per_block(cuda_device, BH, T, d, causal, p_drop, bf16):
    """The persistent forward (one CTA per SM draws 256-row query blocks from a library-owned counter that the kernel leaves
    zeroed; K/V ring, S / P / O hand-offs run on across items) against the same kernel launched with one CTA per block:
    O, L and the FA-1 statistics bit-identical (no cross-CTA arithmetic), twice in a row (the counter slot was reset).
    bf16 and tf32 (fp32 storage) kernels; many more items than SMs, one-step items, ragged T, dropout."""
    import torch
    lib = _lib.load_library()
    q, k, v, g = _bf16_problem(BH, T, d, seed=700 + BH + T)
    tq, tk, tv = (_to_dev(x, cuda_device, bf16=bf16) for x in (q, k, v))
    kw = dict(p_drop=p_drop, seed=99) if p_drop else {}
    prec = dict() if bf16 else dict(precision=_lib.FA_PREC_TF32_TC)
    prev = lib.fa_b200_get_fwd_persistent()
    try:
        lib.fa_b200_set_fwd_persistent(0)
        ref = device.flash_attn_fwd(tq, tk, tv, causal, want_lm=not p_drop, **kw, **prec)
        lib.fa_b200_set_fwd_persistent(1)
        for _ in range(2):
            got = device.flash_attn_fwd(tq, tk, tv, causal, want_lm=not p_drop, **kw, **prec)
            for a, b in zip(got, ref):
                assert torch.equal(a, b)
    finally:
        lib.fa_b200_set_fwd_persistent(prev)


@pytest.mark.parametrize("BH,T,d,causal,p_drop", [(40, 1024, 128, True, 0.0), (40, 1024, 128, False, 0.0), (5, 2125, 128, True, 0.0),
                                                 (300, 128, 128, True, 0.0), (37, 700, 64, True, 0.0), (21, 1100, 128, True, 0.1),
                                                 (3, 4224, 128, True, 0.0)])
def test_persistent_backward_equals_one_cta_per_work_item(cuda_device, BH, T, d, causal, p_drop):
    """The persistent backward (148 CTAs draw (head, K/V tile) items from a counter; K/V buffers swap roles from item to item,
    the first S of the next item is issued ahead of the last dK / dQ of the current one) against the same kernel launched
    with one CTA per item: dK, dV bit-identical (no atomics on their path, the same MMA sequence per item), dQ equal up
    to the order of its fp32 reduce-adds; two persistent runs give bit-identical dK, dV.  Shapes: many more items than
    SMs, one-pair items (T = 128), ragged T, d = 64 (two dO stages), dropout, one head longer than a wave."""
    import torch
    lib = _lib.load_library()
    q, k, v, g = _bf16_problem(BH, T, d, seed=900 + BH + T)
    tq, tk, tv, tg = (_to_dev(x, cuda_device) for x in (q, k, v, g))
    kw = dict(p_drop=p_drop, seed=99) if p_drop else {}
    o, lse = device.flash_attn_fwd(tq, tk, tv, causal, **kw)
    prev = lib.fa_b200_get_bwd_persistent()
    try:
        lib.fa_b200_set_bwd_persistent(0)
        dq0, dk0, dv0 = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal, **kw)
        lib.fa_b200_set_bwd_persistent(1)
        dq1, dk1, dv1 = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal, **kw)
        dq2, dk2, dv2 = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal, **kw)
    finally:
        lib.fa_b200_set_bwd_persistent(prev)
    assert torch.equal(dk0, dk1) and torch.equal(dv0, dv1)
    assert torch.equal(dk1, dk2) and torch.equal(dv1, dv2)
    ref = dq0.float()
    assert float((dq1.float() - ref).abs().max()) <= 8e-3 * float(ref.abs().max())  # one bf16 ulp of the largest element
    assert bool(torch.isfinite(dq1.float()).all())


def test_device_bf16_large_scores_exercise_lazy_rescale(cuda_device):
    """|s| up to ~40: the running max moves by far more than the 2^8 slack, so O must be rescaled in TMEM."""
    q, k, v, g = _bf16_problem(2, 1024, 128, seed=77, scale=6.0)
    for causal in (False, True):
        o_r, p, m_r, l_r, L_r = ao.attention_forward(q, k, v, causal, np.float64)
        o, lse = device.flash_attn_fwd(*(_to_dev(x, cuda_device) for x in (q, k, v)), causal)
        assert rel_inf(o.float().cpu().numpy(), o_r) <= BF16_REL
        assert np.abs(lse.cpu().numpy() - L_r).max() <= 1e-2


@pytest.mark.parametrize("d", [32, 34, 100])
def test_device_fp32_any_head_dim(cuda_device, d):
    """The exact fp32 CUDA-core path (FA_PREC_FP32_SIMT) at any head dim: the MT model's d=32, the reference test's d=34."""
    q, k, v, g = ao.make_inputs(2, 77, d, seed=d)
    dq_r, dk_r, dv_r, o_r, L_r = ao.attention_backward(q, k, v, g, True, np.float64)
    tq, tk, tv, tg = (_to_dev(x, cuda_device, bf16=False) for x in (q, k, v, g))
    o, lse = device.flash_attn_fwd(tq, tk, tv, True, precision=_lib.FA_PREC_FP32_SIMT)
    assert np.abs(o.cpu().numpy() - o_r).max() <= 1e-5
    dq, dk, dv = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, True, precision=_lib.FA_PREC_FP32_SIMT)
    for got, ref in ((dq, dq_r), (dk, dk_r), (dv, dv_r)):
        assert np.abs(got.cpu().numpy() - ref).max() <= 1e-4


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("BH,T,d", [(3, 128, 64), (2, 40, 32), (2, 327, 64), (1, 1000, 32), (2, 1, 64), (1, 1024, 64)])
def test_device_tf32_forward_vs_oracle(cuda_device, BH, T, d, causal):
    """fp32 storage, tcgen05 kind::tf32 arithmetic (what FA_PREC_AUTO selects for fp32 tensors with d in {32, 64}):
    north_star bar max-abs 1e-3 on O; the saved statistics feed the exact-fp32 backward."""
    q, k, v, g = ao.make_inputs(BH, T, d, seed=500 + T + d)
    dq_r, dk_r, dv_r, o_r, L_r = ao.attention_backward(q, k, v, g, causal, np.float64)
    tq, tk, tv, tg = (_to_dev(x, cuda_device, bf16=False) for x in (q, k, v, g))
    before = _lib.launch_count()
    o, lse, m, l = device.flash_attn_fwd(tq, tk, tv, causal, precision=_lib.FA_PREC_TF32_TC, want_lm=True)
    assert _lib.launch_count() == before + 1
    assert np.abs(o.cpu().numpy() - o_r).max() <= FP32_ATOL
    assert np.abs(lse.cpu().numpy() - L_r).max() <= FP32_ATOL
    import torch
    assert np.abs((m + torch.log(l)).cpu().numpy() - L_r).max() <= FP32_ATOL
    o_auto, _ = device.flash_attn_fwd(tq, tk, tv, causal)  # AUTO resolves to the same kernel: bit-identical
    assert torch.equal(o_auto, o)
    dq, dk, dv = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal)
    for got, ref, name in ((dq, dq_r, "dq"), (dk, dk_r, "dk"), (dv, dv_r, "dv")):
        assert np.abs(got.cpu().numpy() - ref).max() <= FP32_ATOL, name


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("d", [32, 64])
@pytest.mark.parametrize("BH,T", [(3, 1), (2, 40), (2, 65), (3, 128), (2, 327), (4, 512), (1, 1000), (1, 2125)])
def test_device_tf32_backward_vs_oracle(cuda_device, BH, T, causal, d):
    """fp32 storage, head dim 32 (the reference MT model's) and 64 (the reference kernel tests',
    kernel_tests/test_flashattn_bw.py:205-210): the backward runs on tcgen05 kind::tf32 (what FA_PREC_AUTO selects) -- prep +
    main kernel = 2 launches -- and meets the north_star bar max-abs 1e-3 against the fp64 oracle for dQ, dK, dV; with
    accumulate_dq it adds onto the caller's dq like the reference (src/flash_attn2_bw.cu:228); the exact-fp32 CUDA-core
    backward stays available through FA_PREC_FP32_SIMT and agrees to 1e-3."""
    import torch
    q, k, v, g = ao.make_inputs(BH, T, d, seed=900 + T)
    dq_r, dk_r, dv_r, o_r, L_r = ao.attention_backward(q, k, v, g, causal, np.float64)
    tq, tk, tv, tg = (_to_dev(x, cuda_device, bf16=False) for x in (q, k, v, g))
    o, lse = device.flash_attn_fwd(tq, tk, tv, causal, precision=_lib.FA_PREC_FP32_SIMT)
    before = _lib.launch_count()
    dq, dk, dv = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal, precision=_lib.FA_PREC_TF32_TC)
    assert _lib.launch_count() == before + 2
    for got, ref, name in ((dq, dq_r, "dq"), (dk, dk_r, "dk"), (dv, dv_r, "dv")):
        assert np.abs(got.cpu().numpy() - ref).max() <= FP32_ATOL, name
    dq_a, dk_a, dv_a = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal)  # AUTO: the same kernel, bit-identical dK / dV
    assert torch.equal(dk_a, dk) and torch.equal(dv_a, dv)
    assert float((dq_a - dq).abs().max()) <= 1e-5  # (dQ: fp32 reduce-adds in L2 arrive in any order)
    dq_s, dk_s, dv_s = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal, precision=_lib.FA_PREC_FP32_SIMT)
    for a, b in ((dq, dq_s), (dk, dk_s), (dv, dv_s)):
        assert float((a - b).abs().max()) <= FP32_ATOL
    # accumulate_dq through the C ABI
    lib = _lib.load_library()
    ws = device.bwd_workspace(tq, _lib.FA_PREC_TF32_TC)
    acc = torch.full_like(tq, 0.25)
    dk2, dv2 = torch.empty_like(tq), torch.empty_like(tq)
    _lib.check(lib.fa_b200_bwd(tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), o.data_ptr(), tg.data_ptr(), lse.data_ptr(),
                               acc.data_ptr(), dk2.data_ptr(), dv2.data_ptr(), _lib.FA_DTYPE_F32, _lib.FA_PREC_TF32_TC, BH, T, d,
                               1 if causal else 0, 1, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    assert np.abs(acc.cpu().numpy() - 0.25 - dq_r).max() <= FP32_ATOL
    assert torch.equal(dk2, dk) and torch.equal(dv2, dv)


def test_cfg2_full_size_tf32_forward(cuda_device):
    """configs[1] at full size (B=32 H=16 T=1024 d=64, fp32 storage, non-causal): rows of P sum to one (V = 1 gives
    O = 1), bit-exact determinism, and sampled heads against the C oracle at max-abs 1e-3."""
    import torch
    torch.manual_seed(2001)
    BH, T, d = 512, 1024, 64
    q, k, v = (torch.rand(BH, T, d, device=cuda_device) * 2 - 1 for _ in range(3))
    o, lse = device.flash_attn_fwd(q, k, v, False, precision=_lib.FA_PREC_TF32_TC)
    o2, lse2 = device.flash_attn_fwd(q, k, v, False, precision=_lib.FA_PREC_TF32_TC)
    assert torch.equal(o, o2) and torch.equal(lse, lse2)
    ones, _ = device.flash_attn_fwd(q, k, torch.ones_like(v), False, precision=_lib.FA_PREC_TF32_TC)
    assert float((ones - 1).abs().max()) <= FP32_ATOL
    for h in (0, 255, 511):
        f = lambda t: t[h:h + 1].cpu().numpy()  # noqa: E731
        r = c_oracle.attention(f(q), f(k), f(v), None, causal=False)
        assert np.abs(f(o) - r["o"]).max() <= FP32_ATOL
        assert np.abs(f(lse) - r["lse"]).max() <= FP32_ATOL


def test_host_api_bf16_precision(cuda_device):
    """The reference ABI (host fp32 buffers) computing in bf16 on the tensor cores."""
    lib = _lib.load_library()
    BH, T, d = 4, 512, 128
    q, k, v, g = _bf16_problem(BH, T, d, seed=9)
    dq_r, dk_r, dv_r, o_r, L_r = ao.attention_backward(q, k, v, g, True, np.float64)
    out = np.zeros_like(q)
    l = np.zeros((BH, T), np.float32)
    m = np.full((BH, T), -np.finfo(np.float32).max, np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    prev = lib.fa_b200_get_host_precision()
    lib.fa_b200_set_host_precision(_lib.FA_PREC_BF16_TC)
    try:
        _lib.check(lib.fa_b200_host_fwd(p(q), p(k), p(v), p(out), p(l), p(m), BH, T, d, 1, _lib.FA_CONV_FA2, None))
        assert rel_inf(out, o_r) <= BF16_REL and np.abs(l - L_r).max() <= 2e-3
        qg, kg, vg = np.zeros_like(q), np.zeros_like(q), np.zeros_like(q)
        _lib.check(lib.fa_b200_host_bwd(p(q), p(k), p(v), p(out), p(g), p(qg), p(kg), p(vg), p(l), p(m), BH, T, d, 1,
                                        _lib.FA_CONV_FA2, None))
        for got, ref in ((qg, dq_r), (kg, dk_r), (vg, dv_r)):
            assert rel_inf(got, ref) <= BF16_REL
    finally:
        lib.fa_b200_set_host_precision(prev)


# ----------------------------------------------------------------------------------------------
# 3. BASELINE.json full sizes: size-independent properties + sampled heads against the C oracle
# ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg3(cuda_device):
    """configs[2]: B=32 H=16 T=4096 d=128 bf16 causal (a quarter of the heads: 128 of 512, to bound test time)."""
    import torch
    torch.manual_seed(3001)
    BH, T, d = 128, 4096, 128
    mk = lambda: (torch.rand(BH, T, d, device=cuda_device) * 2 - 1).to(torch.bfloat16)  # noqa: E731
    return dict(BH=BH, T=T, d=d, q=mk(), k=mk(), v=mk(), g=mk())


def test_full_size_rows_of_p_sum_to_one(cuda_device, cfg3):
    import torch
    ones = torch.ones_like(cfg3["v"])
    o, lse = device.flash_attn_fwd(cfg3["q"], cfg3["k"], ones, True)
    assert float((o.float() - 1).abs().max()) <= 1e-2  # V = 1  =>  O = rowsum(P) = 1


def test_full_size_causality_and_determinism(cuda_device, cfg3):
    import torch
    q, k, v, T = cfg3["q"], cfg3["k"], cfg3["v"], cfg3["T"]
    o1, l1 = device.flash_attn_fwd(q, k, v, True)
    o2, l2 = device.flash_attn_fwd(q, k, v, True)
    assert torch.equal(o1, o2) and torch.equal(l1, l2)  # forward is deterministic
    k2, v2 = k.clone(), v.clone()
    k2[:, T // 2:] = 3.0  # the future may not leak into the first half
    v2[:, T // 2:] = -5.0
    o3, l3 = device.flash_attn_fwd(q, k2, v2, True)
    assert torch.equal(o1[:, :T // 2], o3[:, :T // 2]) and torch.equal(l1[:, :T // 2], l3[:, :T // 2])


def test_full_size_sampled_heads_vs_c_oracle(cuda_device, cfg3):
    """Two (b, h) slices of the full-size problem against the CPU oracle on identical bf16-rounded inputs."""
    q, k, v, g = cfg3["q"], cfg3["k"], cfg3["v"], cfg3["g"]
    o, lse = device.flash_attn_fwd(q, k, v, True)
    dq, dk, dv = device.flash_attn_bwd(q, k, v, o, g, lse, True)
    for h in (0, 97):
        f = lambda t: t[h:h + 1].float().cpu().numpy()  # noqa: E731
        r = c_oracle.attention(f(q), f(k), f(v), f(g), causal=True)
        assert rel_inf(f(o), r["o"].astype(np.float64)) <= BF16_REL
        assert np.abs(f(lse) - r["lse"]).max() <= 5e-3
        for got, key in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
            assert rel_inf(f(got), r[key].astype(np.float64)) <= BF16_REL, key


def test_full_size_backward_identities(cuda_device, cfg3):
    """sum_j dV_j = sum_i dO_i (rows of P sum to one) and <dQ, Q> = <dK, K> (S is bilinear), per head; and
    backward is linear in dO."""
    import torch
    q, k, v, g = cfg3["q"], cfg3["k"], cfg3["v"], cfg3["g"]
    o, lse = device.flash_attn_fwd(q, k, v, True)
    dq, dk, dv = device.flash_attn_bwd(q, k, v, o, g, lse, True)
    lhs, rhs = dv.float().sum(1), g.float().sum(1)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) <= BF16_REL
    a = (dq.float() * q.float()).sum((1, 2))
    b = (dk.float() * k.float()).sum((1, 2))
    assert float((a - b).abs().max() / a.abs().max()) <= 5e-2
    dq2, dk2, dv2 = device.flash_attn_bwd(q, k, v, o, (g.float() * 2).to(torch.bfloat16), lse, True)
    assert float((dv2.float() - 2 * dv.float()).abs().max() / dv.float().abs().max()) <= BF16_REL
    assert float((dq2.float() - 2 * dq.float()).abs().max() / dq.float().abs().max()) <= BF16_REL


def test_cfg2_shape_fp32_forward_sampled(cuda_device):
    """configs[1]: non-causal forward, T=1024, d=64, fp32 storage through the reference ABI (a slice of B*H)."""
    BH, T, d = 16, 1024, 64
    q, k, v, g = ao.make_inputs(BH, T, d, seed=2001)
    r = c_oracle.attention(q, k, v, None, causal=False)
    be = mt.TensorBackend(mt.CudaKernelOps)
    out = _mt4(q, be).flash_attn2(_mt4(k, be), _mt4(v, be), causal_mask=False)
    assert np.abs(out.to_numpy()[0] - r["o"]).max() <= FP32_ATOL


# ----------------------------------------------------------------------------------------------
# 3b. fused attention dropout (BASELINE.json configs[3]): the kernels' mask is exported and fed to the oracle
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,p,seed,head_begin", [(1, 0.1, 1, 0), (77, 0.1, 4001, 0), (256, 0.5, 2**40 + 17, 3), (513, 0.03, 9, 500)])
def test_dropout_mask_matches_oracle_restatement_bit_for_bit(cuda_device, T, p, seed, head_begin):
    keep = device.dropout_mask(3, T, p, seed, device=cuda_device, head_begin=head_begin).cpu().numpy().astype(bool)
    ref = ao.dropout_keep_mask(seed, np.arange(head_begin, head_begin + 3), T, p)
    assert np.array_equal(keep, ref)
    assert abs(device.dropout_keep_scale(p) - ao.dropout_keep_scale(p)) < 1e-7
    if T >= 256:  # Bernoulli(1 - p_eff) within 5 sigma
        p_eff = ao.dropout_threshold(p) / 128.0
        assert abs(keep.mean() - (1 - p_eff)) <= 5 * np.sqrt(p_eff * (1 - p_eff) / keep.size)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("BH,T,d", [(2, 77, 32), (2, 130, 64), (1, 200, 100)])
def test_dropout_fp32_path_vs_oracle(cuda_device, BH, T, d, causal):
    """Exact-fp32 kernels with fused dropout against the oracle evaluated with the exported mask."""
    p, seed = 0.1, 4001
    q, k, v, g = ao.make_inputs(BH, T, d, seed=d + T)
    keep = device.dropout_mask(BH, T, p, seed, device=cuda_device).cpu().numpy()
    dq_r, dk_r, dv_r, o_r, L_r = ao.attention_dropout_backward(q, k, v, g, causal, keep, ao.dropout_keep_scale(p))
    tq, tk, tv, tg = (_to_dev(x, cuda_device, bf16=False) for x in (q, k, v, g))
    o, lse = device.flash_attn_fwd(tq, tk, tv, causal, precision=_lib.FA_PREC_FP32_SIMT, p_drop=p, seed=seed)
    assert np.abs(o.cpu().numpy() - o_r).max() <= 1e-5
    assert np.abs(lse.cpu().numpy() - L_r).max() <= 1e-5  # statistics of the undropped scores
    dq, dk, dv = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal, precision=_lib.FA_PREC_FP32_SIMT, p_drop=p, seed=seed)
    for got, ref in ((dq, dq_r), (dk, dk_r), (dv, dv_r)):
        assert np.abs(got.cpu().numpy() - ref).max() <= 1e-4


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("BH,T,d,p", [(3, 128, 128, 0.1), (2, 300, 64, 0.1), (2, 1000, 128, 0.25), (4, 2, 64, 0.1), (2, 1024, 128, 0.1)])
def test_dropout_bf16_tensor_core_path_vs_oracle(cuda_device, BH, T, d, p, causal):
    """tcgen05 forward (thread = query row) and transposed backward (thread = key row) regenerate the same mask."""
    seed = 4001 + T
    q, k, v, g = _bf16_problem(BH, T, d, seed=200 + T + d)
    keep = device.dropout_mask(BH, T, p, seed, device=cuda_device).cpu().numpy()
    dq_r, dk_r, dv_r, o_r, L_r = ao.attention_dropout_backward(q, k, v, g, causal, keep, ao.dropout_keep_scale(p))
    tq, tk, tv, tg = (_to_dev(x, cuda_device) for x in (q, k, v, g))
    o, lse = device.flash_attn_fwd(tq, tk, tv, causal, p_drop=p, seed=seed)
    assert rel_inf(o.float().cpu().numpy(), o_r) <= BF16_REL
    assert np.abs(lse.cpu().numpy() - L_r).max() <= 2e-3
    dq, dk, dv = device.flash_attn_bwd(tq, tk, tv, o, tg, lse, causal, p_drop=p, seed=seed)
    for got, ref, name in ((dq, dq_r, "dq"), (dk, dk_r, "dk"), (dv, dv_r, "dv")):
        assert rel_inf(got.float().cpu().numpy(), ref) <= BF16_REL, name
    # a different seed draws a different mask; p_drop = 0 is the plain kernel
    o2, _ = device.flash_attn_fwd(tq, tk, tv, causal, p_drop=p, seed=seed + 1)
    if T > 1:
        assert not np.array_equal(o2.float().cpu().numpy(), o.float().cpu().numpy())


def test_dropout_tf32_forward_vs_oracle(cuda_device):
    p, seed, BH, T, d = 0.1, 77, 2, 333, 64
    q, k, v, g = ao.make_inputs(BH, T, d, seed=12)
    keep = device.dropout_mask(BH, T, p, seed, device=cuda_device).cpu().numpy()
    _, _, _, o_r, L_r = ao.attention_dropout_backward(q, k, v, g, True, keep, ao.dropout_keep_scale(p))
    tq, tk, tv = (_to_dev(x, cuda_device, bf16=False) for x in (q, k, v))
    o, lse = device.flash_attn_fwd(tq, tk, tv, True, precision=_lib.FA_PREC_TF32_TC, p_drop=p, seed=seed)
    assert np.abs(o.cpu().numpy() - o_r).max() <= FP32_ATOL and np.abs(lse.cpu().numpy() - L_r).max() <= FP32_ATOL


@pytest.fixture(scope="module")
def cfg4(cuda_device):
    """configs[3]: B=4 H=32 T=16384 d=128 bf16 causal, dropout 0.1 (32 of the 128 heads, to bound test time)."""
    import torch
    torch.manual_seed(4001)
    BH, T, d = 32, 16384, 128
    mk = lambda: (torch.rand(BH, T, d, device=cuda_device) * 2 - 1).to(torch.bfloat16)  # noqa: E731
    return dict(BH=BH, T=T, d=d, q=mk(), k=mk(), v=mk(), g=mk(), p=0.1, seed=4001)


def test_cfg4_long_context_dropout_properties(cuda_device, cfg4):
    """Size-independent properties at T = 16384: determinism, unbiasedness (V = 1 gives E[O] = 1 with the variance
    of a Bernoulli thinning), backward linear in dO, and sum_j dV_j = sum_i Z-weighted dO."""
    import torch
    q, k, v, g, p, seed = (cfg4[x] for x in ("q", "k", "v", "g", "p", "seed"))
    o1, l1 = device.flash_attn_fwd(q, k, v, True, p_drop=p, seed=seed)
    o2, l2 = device.flash_attn_fwd(q, k, v, True, p_drop=p, seed=seed)
    assert torch.equal(o1, o2) and torch.equal(l1, l2)
    o0, l0 = device.flash_attn_fwd(q, k, v, True)
    assert float((l1 - l0).abs().max()) == 0.0  # the statistics ignore dropout
    ones, _ = device.flash_attn_fwd(q, k, torch.ones_like(v), True, p_drop=p, seed=seed)
    late = ones[:, 4096:, 0].float()  # rows with >= 4096 keys: O = sum_j P_ij Z_ij, mean 1
    assert abs(float(late.mean()) - 1.0) <= 2e-3
    dq, dk, dv = device.flash_attn_bwd(q, k, v, o1, g, l1, True, p_drop=p, seed=seed)
    dq2, dk2, dv2 = device.flash_attn_bwd(q, k, v, o1, (g.float() * 2).to(torch.bfloat16), l1, True, p_drop=p, seed=seed)
    for a, b in ((dq2, dq), (dk2, dk), (dv2, dv)):
        # doubling dO doubles delta too (delta = rowsum(dO o O)): the whole backward is linear in dO
        assert float((a.float() - 2 * b.float()).abs().max() / b.float().abs().max()) <= BF16_REL


def test_cfg4_sampled_head_vs_c_oracle_with_exported_mask(cuda_device, cfg4):
    """One (b, h) slice of the long-context problem against the CPU oracle fed with the mask the kernels drew."""
    q, k, v, g, p, seed, T = (cfg4[x] for x in ("q", "k", "v", "g", "p", "seed", "T"))
    o, lse = device.flash_attn_fwd(q, k, v, True, p_drop=p, seed=seed)
    dq, dk, dv = device.flash_attn_bwd(q, k, v, o, g, lse, True, p_drop=p, seed=seed)
    h = 19
    keep = device.dropout_mask(1, T, p, seed, device=cuda_device, head_begin=h).cpu().numpy()
    f = lambda t: t[h:h + 1].float().cpu().numpy()  # noqa: E731
    c_oracle.use_all_cores()
    r = c_oracle.attention(f(q), f(k), f(v), f(g), causal=True, keep=keep, keep_scale=ao.dropout_keep_scale(p))
    assert rel_inf(f(o), r["o"].astype(np.float64)) <= BF16_REL
    assert np.abs(f(lse) - r["lse"]).max() <= 5e-3
    for got, key in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
        assert rel_inf(f(got), r[key].astype(np.float64)) <= BF16_REL, key


# ----------------------------------------------------------------------------------------------
# 4. error behaviour and launch accounting
# ----------------------------------------------------------------------------------------------
def test_errors_are_reported_not_swallowed(cuda_device):
    import torch
    x = torch.zeros(2, 16, 96, device=cuda_device, dtype=torch.bfloat16)
    with pytest.raises(_lib.FaError):
        device.flash_attn_fwd(x, x, x, False)  # bf16 storage needs d in {64, 128}
    y = torch.zeros(2, 16, 64, device=cuda_device, dtype=torch.float32)
    lse = torch.zeros(2, 16, device=cuda_device)
    lib = _lib.load_library()
    rc = lib.fa_b200_bwd(y.data_ptr(), y.data_ptr(), y.data_ptr(), y.data_ptr(), y.data_ptr(), lse.data_ptr(),
                         y.data_ptr(), y.data_ptr(), y.data_ptr(), _lib.FA_DTYPE_F32, _lib.FA_PREC_AUTO, 2, 16, 64, 0, 0,
                         None, 0, None)
    assert rc == _lib.FA_ERR_WORKSPACE


def test_launch_counter_counts_kernels(cuda_device):
    import torch
    before = _lib.launch_count()
    x = (torch.rand(2, 256, 64, device=cuda_device) - 0.5).to(torch.bfloat16)
    o, lse = device.flash_attn_fwd(x, x, x, True)
    torch.cuda.synchronize()
    assert _lib.launch_count() == before + 1
