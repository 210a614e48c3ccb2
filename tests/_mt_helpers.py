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


def reference_forward(model, idx):
    """The reference's DecoderLM.forward in eval mode, restated operator by operator from the named parameters of `model`
    (minitorch/modules_transfomer.py: DecoderLM.forward, TransformerLayer.forward = x + attn(ln_1(x)) then + ff(ln_2(.)),
    MultiHeadAttention.forward, FeedForward.forward; LayerNorm1d / GELU: modules_basic.py:166-175, nn.py:208-214).
    Independent of the model's own forward code (its fused residual / LayerNorm / q|k|v paths)."""
    P = {n: p.value.detach() for n, p in model.named_parameters()}
    B, T = idx.shape
    E, H = model.n_embd, model.t_layer_1.attention.n_head

    def ln(x, name):
        mean = x.mean(dim=-1, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
        return P[name + ".weights"] * (x - mean) / torch.sqrt(var + 1e-5) + P[name + ".bias"]

    def lin(x, name):
        return x @ P[name + ".weights"] + P[name + ".bias"]

    def heads(x):
        return x.view(B, T, H, E // H).permute(0, 2, 1, 3)

    x = P["token_embeddings.weights"][idx] + P["position_embeddings.weights"][torch.arange(T)].unsqueeze(0)
    for i in range(1, 5):
        L = f"t_layer_{i}."
        h = ln(x, L + "ln_1")
        q, k, v = (heads(lin(h, L + f"attention.{n}_projection")) for n in "qkv")
        o = vanilla_attention(q, k, v, True).permute(0, 2, 1, 3).reshape(B, T, E)
        x = x + lin(o, L + "attention.out_projection")
        h = lin(ln(x, L + "ln_2"), L + "ff.linear_in")
        h = 0.5 * h * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (h + 0.044715 * h ** 3)))
        x = x + lin(h, L + "ff.linear_out")
    return lin(ln(x, "ln"), "lm_head")
