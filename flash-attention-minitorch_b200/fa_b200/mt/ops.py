"""Device-resident operators of the decoder-only LM (BASELINE.json configs[4]).

Every function restates one piece of the reference's model path on tensors that never leave the GPU (the
reference stages every op through host numpy storage, minitorch/cuda_kernel_ops.py:302-380):

    linear        x @ W + b                      minitorch/modules_basic.py:117-122  (W is (in, out))
    layernorm     w * (x - mean) / sqrt(var + eps) + b, biased variance     modules_basic.py:166-175
    gelu          tanh approximation             minitorch/nn.py:208-214
    dropout       x * Bernoulli(1-p) / (1-p)     minitorch/modules_basic.py:77-80
    embedding     one_hot(idx) @ W  == W[idx]    minitorch/modules_basic.py:46-53
    add_layernorm     (a + b, layernorm(a + b)): a residual addition of TransformerLayer.forward and the LayerNorm after it
    gelu_dropout      dropout(gelu(x))           FeedForward.forward, minitorch/modules_transfomer.py
    embed_pos_dropout dropout(tok[idx] + pos[arange(T)])   DecoderLM.forward, minitorch/modules_transfomer.py
    softmax_loss  logsumexp(logits) - logits[target], no reduction           minitorch/nn.py:256-274
    flash_attention  q.flash_attn2(k, v, causal) minitorch/modules_transfomer.py:109-120 -> this repo's kernels

On CUDA fp32 tensors every operator runs a hand-written kernel of libfa_b200.so: attention on the flash kernels, the
linear layers (forward, dX, dW) on the tf32 tcgen05 GEMM, LayerNorm / GELU / dropout / embedding / loss as one fused pass
each (rows N1-N4 of SURVEY.md §8f), with the residual additions, the gradient sums at the residual branch points and the
bias gradients of the linear layers folded into the LayerNorm / GELU / loss kernels that touch the same tensors.  torch
owns device memory, streams and the autograd tape only; the loss weighting / mean and reshapes are the remaining ATen calls.
"""
from __future__ import annotations

import math
import os

import torch

from .. import _lib, device


class _FlashAttn2(torch.autograd.Function):
    """Flash_Attn2 (minitorch/tensor_functions.py:476-485) on device tensors: forward saves (q, k, v, o, L),
    backward recomputes P from L.  (B, H, T, d) contiguous, fp32 or bf16."""

    @staticmethod
    def forward(ctx, q, k, v, causal, p_drop, seed):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        o, lse = device.flash_attn_fwd(q, k, v, bool(causal), p_drop=float(p_drop), seed=int(seed))
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.causal, ctx.p_drop, ctx.seed = bool(causal), float(p_drop), int(seed)
        return o

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, o, lse = ctx.saved_tensors
        dq, dk, dv = device.flash_attn_bwd(q, k, v, o, d_out.contiguous(), lse, ctx.causal, p_drop=ctx.p_drop,
                                           seed=ctx.seed)
        return dq, dk, dv, None, None, None


class _FlashAttn2Packed(torch.autograd.Function):
    """The same on a packed (3, B, H, T, d) tensor -- the output of the fused q | k | v projection: the backward kernels
    write dq, dk, dv straight into the three slices of ONE (3, B, H, T, d) gradient, which the projection's backward
    GEMMs read as it is (no unbind / stack copy in between)."""

    @staticmethod
    def forward(ctx, qkv, causal):
        q, k, v = qkv[0], qkv[1], qkv[2]
        o, lse = device.flash_attn_fwd(q, k, v, bool(causal))
        ctx.save_for_backward(qkv, o, lse)
        ctx.causal = bool(causal)
        return o

    @staticmethod
    def backward(ctx, d_out):
        qkv, o, lse = ctx.saved_tensors
        dqkv = torch.empty_like(qkv)
        device.flash_attn_bwd(qkv[0], qkv[1], qkv[2], o, d_out.contiguous(), lse, ctx.causal, out=(dqkv[0], dqkv[1], dqkv[2]))
        return dqkv, None


def flash_attention_packed(qkv, causal: bool = True):
    """flash_attention(qkv[0], qkv[1], qkv[2]) for a contiguous (3, B, H, T, d) CUDA tensor."""
    if not qkv.is_cuda:
        raise RuntimeError("flash_attention needs CUDA tensors: the product has no CPU attention path")
    return _FlashAttn2Packed.apply(qkv.contiguous(), causal)


def flash_attention(q, k, v, causal: bool = True, p_drop: float = 0.0, seed: int = 0):
    """softmax(q k^T / sqrt(d) [+ causal mask]) v on the B200 kernels; q, k, v: (B, H, T, d) CUDA tensors.
    fp32 tensors take the fp32 path (tf32 tensor-core forward and -- at d = 32 -- backward where covered, else the exact
    fp32 CUDA-core kernels; 1e-3 bar)."""
    if not q.is_cuda:
        raise RuntimeError("flash_attention needs CUDA tensors: the product has no CPU attention path")
    return _FlashAttn2.apply(q, k, v, causal, p_drop, seed)


flash_attention.packed = flash_attention_packed


def flash_attention_bf16(q, k, v, causal: bool = True, p_drop: float = 0.0, seed: int = 0):
    """The bf16 tensor-core path for fp32 model tensors (2e-2 relative bar): cast to bf16, run forward AND backward on
    tcgen05, cast back.  Head dims below 64 are zero-padded to 64 (the scores and the softmax scale are computed for
    the padded dim, so q is pre-multiplied by sqrt(64 / d)); the padding columns of O and of the gradients are zero
    and sliced away."""
    if not q.is_cuda:
        raise RuntimeError("flash_attention needs CUDA tensors: the product has no CPU attention path")
    d = q.shape[-1]
    if d in (64, 128):
        qb, kb, vb = q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
        return _FlashAttn2.apply(qb, kb, vb, causal, p_drop, seed).to(q.dtype)
    if d > 64:
        raise ValueError(f"bf16 tensor-core path needs d <= 64 or d = 128 (got {d})")
    pad = 64 - d
    fix = math.sqrt(64.0 / d)  # kernels scale by 1/sqrt(64); the model's scale is 1/sqrt(d)
    qb = torch.nn.functional.pad(q * fix, (0, pad)).to(torch.bfloat16)
    kb = torch.nn.functional.pad(k, (0, pad)).to(torch.bfloat16)
    vb = torch.nn.functional.pad(v, (0, pad)).to(torch.bfloat16)
    return _FlashAttn2.apply(qb, kb, vb, causal, p_drop, seed)[..., :d].to(q.dtype)


# ----------------------------------------------------------------------------------------------
# Kernel-backed operators.  CUDA fp32 tensors run this repo's kernels (libfa_b200.so: tf32 tcgen05 GEMM, fused
# LayerNorm / GELU / dropout / embedding / loss); anything else (the float64 CPU tensors of the unit tests) evaluates
# the reference formula with torch, which is also what the GPU tests compare the kernels against.
# ----------------------------------------------------------------------------------------------
def _on_kernels(*tensors) -> bool:
    """True: every tensor is CUDA fp32 -> this repo's kernels.  False: every tensor lives on the CPU (unit tests).  A CUDA
    tensor of any other dtype, or a CUDA / CPU mix, is an error: the product has no ATen / cuBLAS route on the GPU."""
    on = [t.is_cuda for t in tensors]
    if not any(on):
        return False
    if not all(on) or any(t.dtype != torch.float32 for t in tensors):
        raise TypeError("fa_b200.mt.ops: CUDA operands must all be float32 CUDA tensors (got " +
                        ", ".join(f"{t.device.type}/{str(t.dtype).replace('torch.', '')}" for t in tensors) +
                        "); there is no library fallback on the GPU")
    return True


def _st():
    return torch.cuda.current_stream().cuda_stream


def _gemm(a, b, out_shape, M, N, K, a_mn, b_mn, bias=None, into=None):
    """C = A B (+ bias) into a fresh tensor, or C += A B when `into` is given (a parameter's gradient buffer)."""
    c = into if into is not None else torch.empty(out_shape, device=a.device, dtype=torch.float32)
    _lib.check(_lib.load_library().fa_b200_gemm_tf32(a.data_ptr(), b.data_ptr(), c.data_ptr(),
                                                     bias.data_ptr() if bias is not None else None, M, N, K, a_mn, b_mn,
                                                     1 if into is not None else 0, _st()))
    return c


class grad_sink:
    """With flattened parameters (Module.flatten_parameters) every Parameter.value.grad is a view of ONE static flat
    buffer that zero_grad() clears with a single memset.  The weight / bias gradient kernels then accumulate STRAIGHT into
    those views (GEMM epilogue `C +=`, column sums with accumulate) and the autograd Functions return None for the
    parameter inputs -- autograd's AccumulateGrad would otherwise launch one extra `grad += dW` kernel per parameter per
    step (~70 of the ~280 launches of a step, 13 % of its device time) and allocate every dW / db twice.

    The switch is per tensor, not per process: flatten_parameters marks the views it creates (`mark`), every other tensor
    -- another model in the same process, a parameter of a model that was never flattened -- keeps ordinary autograd
    semantics (torch.autograd.grad, gradient hooks, retain_grad)."""
    ATTR = "_fa_b200_flat_grad"

    @staticmethod
    def mark(t) -> None:
        setattr(t, grad_sink.ATTR, True)


def _sink(p):
    """The gradient buffer the kernels may accumulate into for parameter tensor `p`, or None (ordinary autograd)."""
    if (p is not None and getattr(p, grad_sink.ATTR, False) and p.is_leaf and p.requires_grad and p.grad is not None
            and p.grad.is_contiguous()):
        return p.grad
    return None


def _colsum(dy2d, M, N, into=None):
    db = into if into is not None else torch.empty(N, device=dy2d.device, dtype=torch.float32)
    _lib.check(_lib.load_library().fa_b200_colsum_f32(dy2d.data_ptr(), db.data_ptr(), M, N, 1 if into is not None else 0, _st()))
    return db


class _Linear(torch.autograd.Function):
    """y = x @ W + b with W (in, out): forward and both backward products on the tf32 tensor-core GEMM, no transposes."""

    @staticmethod
    def forward(ctx, x, w, b, bias_grad=True):
        x, w = x.contiguous(), w.contiguous()
        M, K = x.shape
        N = w.shape[1]
        ctx.save_for_backward(x, w)
        ctx.bias = b if bias_grad else None  # bias_grad False: add_layernorm(..., b_bias=) reduces it (see there)
        return _gemm(x, w, (M, N), M, N, K, 0, 1, b.contiguous() if b is not None else None)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        M, K = x.shape
        N = w.shape[1]
        dx = _gemm(dy, w, (M, K), M, K, N, 0, 0)      # dX = dY W^T : W rows are the K-major B operand
        gw = _sink(w)
        dw = _gemm(x, dy, (K, N), K, N, M, 1, 1, into=gw)  # dW = X^T dY : both operands MN-major
        db = None
        if ctx.bias is not None:
            gb = _sink(ctx.bias)
            db = _colsum(dy, M, N, into=gb)
            if gb is not None:
                db = None
        return dx, (None if gw is not None else dw), db, None


def linear(x, weight, bias=None, bias_grad: bool = True):
    """x @ W + b.  bias_grad=False (kernel path only): this call leaves the bias gradient to the add_layernorm that
    consumes its output (pass the bias there as b_bias) -- check fused_bias_grad_ok first."""
    M, K = x.shape
    N = weight.shape[1]
    if _on_kernels(x, weight) and M % 4 == 0 and N % 4 == 0 and K % 4 == 0:
        return _Linear.apply(x, weight, bias, bias_grad)
    assert bias_grad, "bias_grad=False needs the kernel path"
    y = x @ weight
    return y if bias is None else y + bias


def fused_bias_grad_ok(like, M: int, weight, bias) -> bool:
    """True when linear / linear_from_heads of an (M, in) input on the device / dtype of `like` run on the kernels AND an
    add_layernorm of their output width can take over the bias reduction (FA_MT_FUSED_BIAS_GRAD=0 turns it off)."""
    K, N = weight.shape
    return (bias is not None and os.environ.get("FA_MT_FUSED_BIAS_GRAD", "1") != "0" and _on_kernels(like, weight, bias)
            and M % 4 == 0 and K % 4 == 0 and bool(_lib.load_library().fa_b200_layernorm_fused_width(N)))


def _gemm_heads(a, b, c, bias, M, N, K, a_mn, b_mn, a_h, b_h, c_h, T, H, d, accumulate=0):
    _lib.check(_lib.load_library().fa_b200_gemm_tf32_heads(a.data_ptr(), b.data_ptr(), c.data_ptr(),
                                                           bias.data_ptr() if bias is not None else None, M, N, K, a_mn, b_mn,
                                                           accumulate, a_h, b_h, c_h, T, H, d, _st()))
    return c


def fused_qkv_enabled() -> bool:
    return os.environ.get("FA_MT_FUSED_QKV", "1") != "0"  # A/B switch


def heads_layout_ok(x, T: int, d: int) -> bool:
    """The TMA-permuting GEMM needs whole 32-column slabs per head and row tiles inside one batch."""
    if os.environ.get("FA_MT_HEAD_LAYOUT", "1") == "0":  # A/B switch: projections as linear + permute + contiguous
        return False
    return x.is_cuda and x.dtype == torch.float32 and T % 128 == 0 and d % 32 == 0


class _LinearToHeads(torch.autograd.Function):
    """(B*T, in) @ W(in, H*d) + b written DIRECTLY as (B, H, T, d): the projection and the
    .view(B,T,H,d).permute(0,2,1,3).contiguous() of project_to_query_key_value (modules_transfomer.py:79-89) in one
    kernel -- the output tensor map does the permutation.  Backward reads dY out of head layout the same way."""

    @staticmethod
    def forward(ctx, x, w, b, B, T, H, d):
        x, w = x.contiguous(), w.contiguous()
        M, K = x.shape
        y = torch.empty(B, H, T, d, device=x.device, dtype=torch.float32)
        _gemm_heads(x, w, y, b.contiguous() if b is not None else None, M, H * d, K, 0, 1, 0, 0, 1, T, H, d)
        ctx.save_for_backward(x, w)
        ctx.dims, ctx.bias = (B, T, H, d), b
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        B, T, H, d = ctx.dims
        dy = dy.contiguous()
        M, K = x.shape
        E = H * d
        dx = _gemm_heads(dy, w, torch.empty(M, K, device=x.device, dtype=torch.float32), None, M, K, E, 0, 0, 1, 0, 0, T, H, d)
        gw = _sink(w)
        dw = _gemm_heads(x, dy, gw if gw is not None else torch.empty(K, E, device=x.device, dtype=torch.float32), None, K, E, M,
                         1, 1, 0, 1, 0, T, H, d, accumulate=1 if gw is not None else 0)
        db = None
        if ctx.bias is not None:
            gb = _sink(ctx.bias)
            db = gb if gb is not None else torch.empty(E, device=x.device, dtype=torch.float32)
            _lib.check(_lib.load_library().fa_b200_colsum_heads_f32(dy.data_ptr(), db.data_ptr(), B, H, T, d,
                                                                    1 if gb is not None else 0, _st()))
            if gb is not None:
                db = None
        return dx, (None if gw is not None else dw), db, None, None, None, None


class _LinearFromHeads(torch.autograd.Function):
    """O(B, H, T, d) read as (B*T, H*d) @ W + b: the permute(0,2,1,3).contiguous().view of MultiHeadAttention.forward
    (modules_transfomer.py:154-157) folded into the A-operand tensor map; backward writes dO straight into head layout."""

    @staticmethod
    def forward(ctx, o, w, b, bias_grad=True):
        o, w = o.contiguous(), w.contiguous()
        B, H, T, d = o.shape
        M, E, N = B * T, H * d, w.shape[1]
        y = torch.empty(M, N, device=o.device, dtype=torch.float32)
        _gemm_heads(o, w, y, b.contiguous() if b is not None else None, M, N, E, 0, 1, 1, 0, 0, T, H, d)
        ctx.save_for_backward(o, w)
        ctx.bias = b if bias_grad else None
        return y

    @staticmethod
    def backward(ctx, dy):
        o, w = ctx.saved_tensors
        B, H, T, d = o.shape
        M, E, N = B * T, H * d, w.shape[1]
        dy = dy.contiguous()
        do = _gemm_heads(dy, w, torch.empty_like(o), None, M, E, N, 0, 0, 0, 0, 1, T, H, d)
        gw = _sink(w)
        dw = _gemm_heads(o, dy, gw if gw is not None else torch.empty(E, N, device=o.device, dtype=torch.float32), None, E, N, M,
                         1, 1, 1, 0, 0, T, H, d, accumulate=1 if gw is not None else 0)
        db = None
        if ctx.bias is not None:
            gb = _sink(ctx.bias)
            db = _colsum(dy, M, N, into=gb)
            if gb is not None:
                db = None
        return do, (None if gw is not None else dw), db, None


class _QKVToHeads(torch.autograd.Function):
    """The three projections of project_to_query_key_value (modules_transfomer.py:79-89) as ONE GEMM: x (B*T, in) times the
    interleaved weight [in][Wq | Wk | Wv] (+ [bq | bk | bv]) written directly as (3, B, H, T, d) -- q, k, v are its three
    slices, already in attention layout.  Backward: one input-gradient GEMM over K = 3 H d reading (dq | dk | dv) out of
    head layout, one weight-gradient GEMM (N = 3 H d) and one bias reduction: 4 launches per layer instead of 12, and no
    summation of three input gradients."""

    @staticmethod
    def forward(ctx, x, wqkv, bqkv, B, T, H, d):
        x = x.contiguous()
        M, K = x.shape
        E3 = 3 * H * d
        y = torch.empty(3, B, H, T, d, device=x.device, dtype=torch.float32)
        _gemm_heads(x, wqkv, y, bqkv, M, E3, K, 0, 1, 0, 0, 1, T, H, d)
        ctx.save_for_backward(x, wqkv)
        ctx.dims, ctx.bias = (B, T, H, d), bqkv
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wqkv = ctx.saved_tensors
        B, T, H, d = ctx.dims
        dy = dy.contiguous()
        M, K = x.shape
        E3 = 3 * H * d
        dx = _gemm_heads(dy, wqkv, torch.empty(M, K, device=x.device, dtype=torch.float32), None, M, K, E3, 0, 0, 1, 0, 0, T, H, d)
        gw = _sink(wqkv)
        dw = _gemm_heads(x, dy, gw if gw is not None else torch.empty(K, E3, device=x.device, dtype=torch.float32), None, K, E3, M,
                         1, 1, 0, 1, 0, T, H, d, accumulate=1 if gw is not None else 0)
        db = None
        if ctx.bias is not None:
            gb = _sink(ctx.bias)
            db = gb if gb is not None else torch.empty(E3, device=x.device, dtype=torch.float32)
            _lib.check(_lib.load_library().fa_b200_colsum_heads_parts_f32(dy.data_ptr(), db.data_ptr(), 3, B, H, T, d,
                                                                          1 if gb is not None else 0, _st()))
            if gb is not None:
                db = None
        return dx, (None if gw is not None else dw), db, None, None, None, None


def qkv_to_heads(x, wqkv, bqkv, B: int, T: int, H: int, d: int):
    """x (B*T, in), interleaved weight (in, 3 H d) -> (3, B, H, T, d); callers check heads_layout_ok first."""
    return _QKVToHeads.apply(x, wqkv, bqkv, B, T, H, d)


def linear_to_heads(x, weight, bias, B: int, T: int, H: int, d: int):
    """x (B*T, in) -> (B, H, T, d) = heads of x @ W + b.  Fused (no permute copy) when the layout allows it."""
    if heads_layout_ok(x, T, d) and _on_kernels(x, weight) and x.shape[1] % 4 == 0:
        return _LinearToHeads.apply(x, weight, bias, B, T, H, d)
    return linear(x, weight, bias).view(B, T, H, d).permute(0, 2, 1, 3).contiguous()


def linear_from_heads(o, weight, bias, bias_grad: bool = True):
    """o (B, H, T, d) -> (B*T, out) = merge_heads(o) @ W + b.  (bias_grad: see linear.)"""
    B, H, T, d = o.shape
    if heads_layout_ok(o, T, d) and _on_kernels(o, weight) and weight.shape[1] % 4 == 0:
        return _LinearFromHeads.apply(o, weight, bias, bias_grad)
    return linear(o.permute(0, 2, 1, 3).contiguous().view(B * T, H * d), weight, bias, bias_grad)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        x = x.contiguous()
        rows, dim = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        _lib.check(_lib.load_library().fa_b200_layernorm_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                                             rstd.data_ptr(), rows, dim, float(eps), _st()))
        ctx.save_for_backward(x, w, mean, rstd)
        ctx.bias = b
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        rows, dim = x.shape
        dx = torch.empty_like(x)
        gw, gb = _sink(w), _sink(ctx.bias)
        lib = _lib.load_library()
        if gw is not None and gb is not None:  # straight into the flat gradient views (grad_sink)
            _lib.check(lib.fa_b200_layernorm_bwd_accum(x.data_ptr(), w.data_ptr(), dy.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                       dx.data_ptr(), gw.data_ptr(), gb.data_ptr(), rows, dim, _st()))
            return dx, None, None, None
        dw, db = torch.empty_like(w), torch.empty_like(w)
        _lib.check(lib.fa_b200_layernorm_bwd(x.data_ptr(), w.data_ptr(), dy.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                             dx.data_ptr(), dw.data_ptr(), db.data_ptr(), rows, dim, _st()))
        return dx, dw, db, None


class _AddLayerNorm(torch.autograd.Function):
    """(a, b) -> (s, y) with s = a + b, y = LayerNorm(s): the residual sum of the layer before and the LayerNorm that
    consumes it, one pass forward (b None: s is a itself).  Backward: ds = LN'(dy) + d_s, where d_s is the gradient that
    reaches s along the residual branch -- one launch including the dw / db reduction, and no separate addition of the two
    branch gradients."""

    @staticmethod
    def forward(ctx, a, b, w, bias, eps, b_bias=None):
        a = a.contiguous()
        rows, dim = a.shape
        y = torch.empty_like(a)
        s = torch.empty_like(a) if b is not None else a
        mean = torch.empty(rows, device=a.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        _lib.check(_lib.load_library().fa_b200_add_layernorm_fwd(a.data_ptr(), b.contiguous().data_ptr() if b is not None else None,
                                                                 w.data_ptr(), bias.data_ptr(), s.data_ptr() if b is not None else None,
                                                                 y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, dim, float(eps), _st()))
        ctx.save_for_backward(s, w, mean, rstd)
        ctx.bias, ctx.has_b, ctx.b_bias = bias, b is not None, (b_bias if b is not None else None)
        ctx.set_materialize_grads(False)  # an unused sum (the model's final ln) must not cost a zero tensor
        return s, y

    @staticmethod
    def backward(ctx, d_s, dy):
        s, w, mean, rstd = ctx.saved_tensors
        rows, dim = s.shape
        bb = ctx.b_bias
        if dy is None:  # only the sum was used
            return d_s, (d_s if ctx.has_b else None), None, None, None, (d_s.sum(0) if bb is not None and d_s is not None else None)
        dy = dy.contiguous()
        ds = torch.empty_like(s)
        gw, gb, gc = _sink(w), _sink(ctx.bias), _sink(bb)
        sink = gw is not None and gb is not None and (bb is None or gc is not None)
        dw, db = (gw, gb) if sink else (torch.empty_like(w), torch.empty_like(w))
        dc = None if bb is None else (gc if sink else torch.empty_like(w))
        _lib.check(_lib.load_library().fa_b200_layernorm_bwd_res(s.data_ptr(), w.data_ptr(), dy.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                                 d_s.contiguous().data_ptr() if d_s is not None else None, ds.data_ptr(),
                                                                 dw.data_ptr(), db.data_ptr(), dc.data_ptr() if dc is not None else None,
                                                                 rows, dim, 1 if sink else 0, _st()))
        return ds, (ds if ctx.has_b else None), (None if sink else dw), (None if sink else db), None, (None if sink else dc)


def add_layernorm(a, b, weight, bias, eps: float, b_bias=None):
    """(a + b, LayerNorm(a + b)) for (rows, dim) tensors; b may be None (then the first result is a).  The pair the
    transformer layers are built from: the sum is the residual stream, the LayerNorm output feeds the next sub-block.
    b_bias: the bias of the linear layer that produced b with bias_grad=False -- the gradient of b is also that layer's dy,
    so its bias gradient (the column sums) comes out of this backward kernel for free."""
    if _on_kernels(a, weight, bias) and (b is None or _on_kernels(b)) and _lib.load_library().fa_b200_layernorm_fused_width(a.shape[1]):
        return _AddLayerNorm.apply(a, b, weight, bias, eps, b_bias)
    assert b_bias is None, "b_bias needs the kernel path (fused_bias_grad_ok)"
    s = a if b is None else a + b
    return s, layernorm(s, weight, bias, eps)


def layernorm(x, weight, bias, eps: float):
    if _on_kernels(x, weight, bias) and x.shape[1] % 4 == 0:
        return _LayerNorm.apply(x, weight, bias, eps)
    mean = x.mean(dim=1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=1, keepdim=True)
    return weight * ((x - mean) / torch.sqrt(var + eps)) + bias


class _Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        _lib.check(_lib.load_library().fa_b200_gelu_fwd(x.data_ptr(), y.data_ptr(), x.numel(), _st()))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        _lib.check(_lib.load_library().fa_b200_gelu_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), _st()))
        return dx


def gelu(x):
    if _on_kernels(x) and x.numel() % 4 == 0:
        return _Gelu.apply(x)
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * (x ** 3))))


class _Dropout(torch.autograd.Function):
    """Mask = f(seed, step counter, element index), regenerated in the backward: no mask tensor is ever materialised."""

    @staticmethod
    def forward(ctx, x, p, seed, step):
        x = x.contiguous()
        y = torch.empty_like(x)
        _lib.check(_lib.load_library().fa_b200_dropout_apply(x.data_ptr(), y.data_ptr(), x.numel(), float(p), int(seed),
                                                             step.data_ptr() if step is not None else None, _st()))
        ctx.p, ctx.seed, ctx.step = float(p), int(seed), step
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        _lib.check(_lib.load_library().fa_b200_dropout_apply(dy.data_ptr(), dx.data_ptr(), dy.numel(), ctx.p, ctx.seed,
                                                             ctx.step.data_ptr() if ctx.step is not None else None, _st()))
        return dx, None, None, None


class _GeluDropout(torch.autograd.Function):
    """dropout(gelu(x)) as one pass each way; the mask is the one `dropout` would draw at this call site (p = 0: none).
    x_bias (optional): the bias of the linear layer that produced x with bias_grad=False -- its gradient, the column sums
    of dx, is reduced in the backward pass."""

    @staticmethod
    def forward(ctx, x, p, seed, step, x_bias=None):
        x = x.contiguous()
        y = torch.empty_like(x)
        _lib.check(_lib.load_library().fa_b200_gelu_dropout_fwd(x.data_ptr(), y.data_ptr(), x.numel(), float(p), int(seed),
                                                                step.data_ptr() if step is not None else None, _st()))
        ctx.save_for_backward(x)
        ctx.p, ctx.seed, ctx.step, ctx.x_bias = float(p), int(seed), step, x_bias
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        lib = _lib.load_library()
        step = ctx.step.data_ptr() if ctx.step is not None else None
        bb = ctx.x_bias
        if bb is None:
            _lib.check(lib.fa_b200_gelu_dropout_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), ctx.p, ctx.seed, step, _st()))
            return dx, None, None, None, None
        gc = _sink(bb)
        dc = gc if gc is not None else torch.empty_like(bb)
        rows, dim = x.shape
        _lib.check(lib.fa_b200_gelu_dropout_bwd_colsum(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), dc.data_ptr(), rows, dim, ctx.p, ctx.seed,
                                                       step, 1 if gc is not None else 0, _st()))
        return dx, None, None, None, (None if gc is not None else dc)


class DropoutState:
    """Seeds of the fused dropout.  Every call site gets its own (base seed, call index) pair; `step` is a device-side
    counter mixed in by the kernel, so a captured CUDA graph (fixed call indices) still draws fresh masks per replay:
    the trainer bumps it once per step."""

    def __init__(self):
        self.base_seed = 0x5EED
        self.calls = 0
        self.steps = {}  # device -> torch.int64 tensor of one element living on that device

    @property
    def step(self):
        """The counter of the current CUDA device (None before the first dropout call on it)."""
        return self.steps.get(torch.device("cuda", torch.cuda.current_device())) if torch.cuda.is_available() else None

    def counter(self, device):
        dev = torch.device(device)
        if dev not in self.steps:
            self.steps[dev] = torch.zeros(1, dtype=torch.int64, device=dev)
        return self.steps[dev]

    def begin_step(self):
        self.calls = 0
        for c in self.steps.values():
            c.add_(1)


dropout_state = DropoutState()


def _next_dropout_seed(x):
    st = dropout_state
    st.calls += 1
    return (st.base_seed << 32) ^ st.calls, st.counter(x.device)


def gelu_dropout(x, p: float, training: bool, generator=None, x_bias=None):
    """dropout(gelu(x), p): one kernel each way on the kernel path, the two reference operators otherwise.
    x_bias: see _GeluDropout (x 2-D, width accepted by fused_bias_grad_ok)."""
    drop = training and p > 0.0
    if (drop or x_bias is not None) and _on_kernels(x) and x.numel() % 4 == 0:
        seed, step = _next_dropout_seed(x) if drop else (0, None)
        return _GeluDropout.apply(x, p if drop else 0.0, seed, step, x_bias)
    assert x_bias is None, "x_bias needs the kernel path (fused_bias_grad_ok)"
    return dropout(gelu(x), p, training, generator)


def dropout(x, p: float, training: bool, generator=None):
    if not training or p <= 0.0:
        return x
    if _on_kernels(x) and x.numel() % 4 == 0:
        seed, step = _next_dropout_seed(x)
        return _Dropout.apply(x, p, seed, step)
    mask = (torch.rand(x.shape, device=x.device, generator=generator) < (1.0 - p)).to(x.dtype)
    return (x * mask) * (1.0 / (1.0 - p))


class _Embedding(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, w):
        idx = idx.contiguous().long()
        w = w.contiguous()
        y = torch.empty(tuple(idx.shape) + (w.shape[1],), device=w.device, dtype=torch.float32)
        _lib.check(_lib.load_library().fa_b200_embedding_fwd(idx.data_ptr(), w.data_ptr(), y.data_ptr(), idx.numel(), w.shape[1], _st()))
        ctx.save_for_backward(idx)
        ctx.vocab, ctx.weight = w.shape[0], w
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        dy = dy.contiguous()
        dim = dy.shape[-1]
        gw = _sink(ctx.weight)
        if gw is not None:  # scatter-add straight into the (already zeroed) flat gradient view: no 10 MB memset + add
            _lib.check(_lib.load_library().fa_b200_embedding_bwd_accum(idx.data_ptr(), dy.data_ptr(), gw.data_ptr(), idx.numel(), dim,
                                                                       ctx.vocab, _st()))
            return None, None
        dw = torch.empty(ctx.vocab, dim, device=dy.device, dtype=torch.float32)
        _lib.check(_lib.load_library().fa_b200_embedding_bwd(idx.data_ptr(), dy.data_ptr(), dw.data_ptr(), idx.numel(), dim, ctx.vocab, _st()))
        return None, dw


class _EmbedPosDropout(torch.autograd.Function):
    """dropout(tok_w[idx] + pos_w[arange(T)]) for idx (B, T): one gather pass forward, one scatter pass backward."""

    @staticmethod
    def forward(ctx, idx, tok_w, pos_w, p, seed, step):
        idx = idx.contiguous().long()
        B, T = idx.shape
        dim = tok_w.shape[1]
        y = torch.empty(B, T, dim, device=tok_w.device, dtype=torch.float32)
        _lib.check(_lib.load_library().fa_b200_embed_pos_dropout_fwd(idx.data_ptr(), tok_w.contiguous().data_ptr(), pos_w.contiguous().data_ptr(),
                                                                     y.data_ptr(), idx.numel(), T, dim, float(p), int(seed),
                                                                     step.data_ptr() if step is not None else None, _st()))
        ctx.save_for_backward(idx)
        ctx.tables, ctx.p, ctx.seed, ctx.step = (tok_w, pos_w), float(p), int(seed), step
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        tok_w, pos_w = ctx.tables
        B, T = idx.shape
        dim = tok_w.shape[1]
        gt, gp = _sink(tok_w), _sink(pos_w)
        sink = gt is not None and gp is not None
        dt, dp = (gt, gp) if sink else (torch.empty_like(tok_w), torch.empty_like(pos_w))
        _lib.check(_lib.load_library().fa_b200_embed_pos_dropout_bwd(idx.data_ptr(), dy.contiguous().data_ptr(), dt.data_ptr(), dp.data_ptr(),
                                                                     idx.numel(), T, dim, tok_w.shape[0], pos_w.shape[0], ctx.p, ctx.seed,
                                                                     ctx.step.data_ptr() if ctx.step is not None else None,
                                                                     1 if sink else 0, _st()))
        return None, (None if sink else dt), (None if sink else dp), None, None, None


validate_ids = True  # range-check token ids / class labels on the kernel path (one reduction + host sync per call site)


def check_ids(idx, n: int, what: str) -> None:
    """The kernels index tables and logit rows with these integers unchecked (gather, scatter-add, row[target]): an id
    outside [0, n) would read or atomically write out of bounds.  The reference raises IndexError for such ids (numpy
    fancy indexing in nn.one_hot, minitorch/nn.py:244-246); so do we -- outside CUDA-graph capture, where a host sync is
    allowed (GraphedStep runs its warm-up steps eagerly, so every training run is checked at least once)."""
    if not validate_ids or idx.numel() == 0 or (idx.is_cuda and torch.cuda.is_current_stream_capturing()):
        return
    lo, hi = int(idx.min()), int(idx.max())
    if lo < 0 or hi >= n:
        raise IndexError(f"{what}: ids must lie in [0, {n}) (found min {lo}, max {hi})")


def embed_pos_dropout(idx, tok_weight, pos_weight, p: float, training: bool, generator=None):
    """dropout(tok_weight[idx] + pos_weight[arange(T)], p) for idx (B, T) -- the model's input stage; one kernel each way on
    the kernel path, the reference operators otherwise."""
    B, T = idx.shape
    if (_on_kernels(tok_weight, pos_weight) and idx.is_cuda and tok_weight.shape[1] % 4 == 0 and tok_weight.shape[1] == pos_weight.shape[1]
            and pos_weight.shape[0] >= T and tok_weight.is_contiguous() and pos_weight.is_contiguous()):
        check_ids(idx, tok_weight.shape[0], "embed_pos_dropout")
        drop = training and p > 0.0
        seed, step = _next_dropout_seed(tok_weight) if drop else (0, None)
        return _EmbedPosDropout.apply(idx, tok_weight, pos_weight, p if drop else 0.0, seed, step)
    pos = torch.arange(T, device=idx.device).view(1, T)
    return dropout(embedding(idx, tok_weight) + embedding(pos, pos_weight), p, training, generator)


def embedding(idx, weight):
    if _on_kernels(weight) and idx.is_cuda and weight.shape[1] % 4 == 0:
        check_ids(idx, weight.shape[0], "embedding")
        return _Embedding.apply(idx, weight)
    return weight[idx]


class _SoftmaxLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, logits_bias=None):
        logits, target = logits.contiguous(), target.contiguous().long()
        rows, C = logits.shape
        loss = torch.empty(rows, device=logits.device, dtype=torch.float32)
        lse = torch.empty_like(loss)
        _lib.check(_lib.load_library().fa_b200_softmax_loss_fwd(logits.data_ptr(), target.data_ptr(), loss.data_ptr(), lse.data_ptr(), rows, C, _st()))
        ctx.save_for_backward(logits, target, lse)
        ctx.logits_bias = logits_bias
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits, target, lse = ctx.saved_tensors
        rows, C = logits.shape
        dlogits = torch.empty_like(logits)
        lib = _lib.load_library()
        bb = ctx.logits_bias
        if bb is None:
            _lib.check(lib.fa_b200_softmax_loss_bwd(logits.data_ptr(), target.data_ptr(), lse.data_ptr(), dloss.contiguous().data_ptr(),
                                                    dlogits.data_ptr(), rows, C, _st()))
            return dlogits, None, None
        gc = _sink(bb)  # the logits layer's bias gradient = column sums of dlogits, reduced in the same pass
        dc = gc if gc is not None else torch.empty(C, device=logits.device, dtype=torch.float32)
        _lib.check(lib.fa_b200_softmax_loss_bwd_colsum(logits.data_ptr(), target.data_ptr(), lse.data_ptr(), dloss.contiguous().data_ptr(),
                                                       dlogits.data_ptr(), dc.data_ptr(), rows, C, 1 if gc is not None else 0, _st()))
        return dlogits, None, (None if gc is not None else dc)


def softmax_loss(logits, target, logits_bias=None):
    """(minibatch, C) logits, (minibatch,) integer targets -> (minibatch,) losses.
    logits_bias: the bias of the linear layer that produced `logits` with bias_grad=False (fused_logits_bias_grad_ok):
    its gradient is reduced inside the loss-backward kernel."""
    if _on_kernels(logits) and target.is_cuda and logits.shape[1] % 4 == 0:
        check_ids(target, logits.shape[1], "softmax_loss")
        return _SoftmaxLoss.apply(logits, target, logits_bias)
    assert logits_bias is None, "logits_bias needs the kernel path"
    return torch.logsumexp(logits, dim=1) - logits.gather(1, target.view(-1, 1).long()).view(-1)


def fused_logits_bias_grad_ok(like, M: int, weight, bias, target) -> bool:
    """The loss backward may take over the bias gradient of linear(like-shaped (M, in) input, weight, bias)."""
    K, N = weight.shape
    return (bias is not None and os.environ.get("FA_MT_FUSED_LOGITS_BIAS_GRAD", "1") != "0" and _on_kernels(like, weight, bias)
            and target.is_cuda and M % 4 == 0 and K % 4 == 0 and bool(_lib.load_library().fa_b200_softmax_loss_colsum_width(N)))


def launch_count() -> int:
    return _lib.launch_count()
