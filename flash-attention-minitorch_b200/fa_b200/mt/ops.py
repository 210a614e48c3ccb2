"""Device-resident operators of the decoder-only LM (BASELINE.json configs[4]).

Every function restates one piece of the reference's model path on tensors that never leave the GPU (the
reference stages every op through host numpy storage, minitorch/cuda_kernel_ops.py:302-380):

    linear        x @ W + b                      minitorch/modules_basic.py:117-122  (W is (in, out))
    layernorm     w * (x - mean) / sqrt(var + eps) + b, biased variance     modules_basic.py:166-175
    gelu          tanh approximation             minitorch/nn.py:208-214
    dropout       x * Bernoulli(1-p) / (1-p)     minitorch/modules_basic.py:77-80
    embedding     one_hot(idx) @ W  == W[idx]    minitorch/modules_basic.py:46-53
    softmax_loss  logsumexp(logits) - logits[target], no reduction           minitorch/nn.py:256-274
    flash_attention  q.flash_attn2(k, v, causal) minitorch/modules_transfomer.py:109-120 -> this repo's kernels

Attention always runs on this repo's CUDA kernels (fa_b200.device; there is no CPU path).  The other operators
are cuBLAS / ATen library calls on the device for now -- rows N2-N4 of SURVEY.md §8(f), next to be replaced by
hand-written kernels; they are already device-resident (row N1), which is what removes the per-op PCIe round trip.
torch is the owner of device memory, streams and autograd bookkeeping only.
"""
from __future__ import annotations

import math

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


def flash_attention(q, k, v, causal: bool = True, p_drop: float = 0.0, seed: int = 0):
    """softmax(q k^T / sqrt(d) [+ causal mask]) v on the B200 kernels; q, k, v: (B, H, T, d) CUDA tensors."""
    if not q.is_cuda:
        raise RuntimeError("flash_attention needs CUDA tensors: the product has no CPU attention path")
    return _FlashAttn2.apply(q, k, v, causal, p_drop, seed)


def linear(x, weight, bias=None):
    y = x @ weight
    return y if bias is None else y + bias


def layernorm(x, weight, bias, eps: float):
    mean = x.mean(dim=1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=1, keepdim=True)
    return weight * ((x - mean) / torch.sqrt(var + eps)) + bias


def gelu(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * (x ** 3))))


def dropout(x, p: float, training: bool, generator=None):
    if not training or p <= 0.0:
        return x
    mask = (torch.rand(x.shape, device=x.device, generator=generator) < (1.0 - p)).to(x.dtype)
    return (x * mask) * (1.0 / (1.0 - p))


def embedding(idx, weight):
    return weight[idx]


def softmax_loss(logits, target):
    """(minibatch, C) logits, (minibatch,) integer targets -> (minibatch,) losses."""
    return torch.logsumexp(logits, dim=1) - logits.gather(1, target.view(-1, 1).long()).view(-1)


def launch_count() -> int:
    return _lib.launch_count()
