"""Test-only helpers for the decoder-LM slice: a plain torch restatement of causal attention that CPU tests inject
into the model in place of the CUDA kernels (the product has no CPU attention path)."""
import math

import torch


def vanilla_attention(q, k, v, causal=True):
    """softmax((q @ kT)/sqrt(d) + M) @ v, M = -FLT_MAX * triu(1)  (minitorch/modules_transfomer.py:121-127)."""
    T, d = q.shape[-2], q.shape[-1]
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(T, T, dtype=torch.bool, device=q.device), 1), -torch.finfo(s.dtype).max)
    return torch.softmax(s, dim=-1) @ v
