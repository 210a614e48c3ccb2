"""Decoder-only Pre-LN transformer of the reference's machine-translation run, device-resident.

Same module tree, parameter names, shapes and initialisation as the reference (minitorch/modules_transfomer.py,
modules_basic.py), so a state dict maps one to one:

    DecoderLM(n_vocab, n_embd, n_head, n_positions, p_dropout, ln_eps, bias)      modules_transfomer.py:253-357
      token_embeddings / position_embeddings : Embedding(n_vocab, n_embd), N(0,1)  modules_basic.py:17-53
      t_layer_1..4 : TransformerLayer  (x + attn(ln_1(x)); out + ff(ln_2(out)))    modules_transfomer.py:204-250
        attention  : MultiHeadAttention, causal, flash path                        modules_transfomer.py:14-157
        ff         : FeedForward(n_embd, middle_dim=256): linear, GELU, dropout, linear   modules_transfomer.py:160-201
      dropout, ln (LayerNorm1d), lm_head : Linear(n_embd, n_vocab)
    Linear weights are (in, out), U(-1/sqrt(in), 1/sqrt(in)) (RParam, modules_basic.py:82-85)

`attention_fn(q, k, v, causal)` defaults to this repo's CUDA kernels; tests on a CPU-only box inject a plain
restatement to exercise the host logic (the product path itself has no CPU attention).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import torch

from . import ops


class Parameter:
    def __init__(self, value: torch.Tensor):
        self.value = value.requires_grad_(True)


class Module:
    training = True

    def named_parameters(self, prefix: str = "") -> Iterator[Tuple[str, Parameter]]:
        for name, v in vars(self).items():
            if isinstance(v, Parameter):
                yield prefix + name, v
            elif isinstance(v, Module):
                yield from v.named_parameters(prefix + name + ".")

    def parameters(self) -> List[Parameter]:
        return [p for _, p in self.named_parameters()]

    def modules(self):
        yield self
        for v in vars(self).values():
            if isinstance(v, Module):
                yield from v.modules()

    def train(self):
        for m in self.modules():
            m.training = True

    def eval(self):
        for m in self.modules():
            m.training = False

    def __call__(self, *a, **kw):
        return self.forward(*a, **kw)

    def flatten_parameters(self, fuse: Optional[bool] = None):
        """Move every parameter into ONE flat buffer (values) with a twin flat buffer for gradients; each
        Parameter.value becomes a view of the first and its .grad a view of the second (the gradient kernels accumulate
        into it in place, ops.grad_sink).  One optimizer kernel and one all-reduce then cover the whole model.
        Modules may ask for several parameters to be laid out together (Module.fuse_parameters): the q / k / v projection
        weights of an attention layer are interleaved row by row into one (in, 3 E) matrix -- each stays addressable as a
        strided view -- so that the layer runs ONE projection GEMM.  `fuse`: None = fuse on CUDA tensors only (default),
        True / False = always / never (the CPU tests exercise the fused layout with True).  Returns (values, grads)."""
        params = self.parameters()
        ref = params[0].value
        # Every parameter / fused group starts on a 32-element (128-byte) boundary: the kernels read weights and biases
        # with float4 loads and as TMA tensors (16-byte aligned base), so one parameter whose size is not a multiple of
        # four (say a 10002-entry vocabulary bias) must not misalign everything laid out after it.
        align = lambda n: (n + 31) // 32 * 32  # noqa: E731

        # Layout units, clustered per top-level child of this module (token_embeddings, ..., t_layer_1, ..., ln, lm_head):
        # a child's own parameters first, then its fused groups -- so that every child is ONE contiguous range of the flat
        # buffers (Module.flat_ranges), which is what lets the trainer all-reduce a layer's gradients as one bucket the
        # moment its backward has finished.
        units = []  # ("p", Parameter) | ("g", members, on_fused)
        self._flat_children = []  # (attribute name, begin, end) in layout order
        for name, child in vars(self).items():
            if isinstance(child, Parameter):
                subtree_params, subtree_groups = [child], []
            elif isinstance(child, Module):
                subtree_groups = [] if fuse is False else [g for m in child.modules() for g in m.fuse_parameters()
                                                           if fuse or all(p.value.is_cuda for p in g[0])]
                grouped = {id(p) for g in subtree_groups for p in g[0]}
                subtree_params = [p for p in child.parameters() if id(p) not in grouped]
            else:
                continue
            self._flat_children.append([name, len(units), None])
            units += [("p", p) for p in subtree_params] + [("g", g[0], g[1]) for g in subtree_groups]
            self._flat_children[-1][2] = len(units)

        def unit_numel(u):
            return u[1].value.numel() if u[0] == "p" else sum(p.value.numel() for p in u[1])

        offsets, off = [], 0
        for u in units:
            offsets.append(off)
            off += align(unit_numel(u))
        total = off
        flat = torch.zeros(total, device=ref.device, dtype=ref.dtype)
        gflat = torch.zeros_like(flat)

        def attach(p, view, gview):
            view.copy_(p.value.detach())
            p.value = view.requires_grad_(True)
            p.value.grad = gview
            ops.grad_sink.mark(p.value)  # the gradient kernels may accumulate straight into gview

        for u, off in zip(units, offsets):
            if u[0] == "p":
                p = u[1]
                n = p.value.numel()
                attach(p, flat[off:off + n].view(p.value.shape), gflat[off:off + n].view(p.value.shape))
                continue
            members, on_fused = u[1], u[2]
            rows = members[0].value.shape[0] if members[0].value.dim() == 2 else 1
            widths = [p.value.numel() // rows for p in members]
            n = rows * sum(widths)
            block, gblock = flat[off:off + n].view(rows, sum(widths)), gflat[off:off + n].view(rows, sum(widths))
            col = 0
            for p, w in zip(members, widths):  # column slices: strided views for matrices, contiguous ones for vectors
                attach(p, block[:, col:col + w].view(p.value.shape) if rows == 1 else block[:, col:col + w],
                       gblock[:, col:col + w].view(p.value.shape) if rows == 1 else gblock[:, col:col + w])
                col += w
            fused = (block if rows > 1 else block.view(-1)).requires_grad_(True)
            fused.grad = gblock if rows > 1 else gblock.view(-1)
            ops.grad_sink.mark(fused)
            on_fused(fused)
        # element ranges of the flat buffers per top-level child (padding included)
        self._flat_children = [(name, offsets[lo] if lo < len(offsets) else total, offsets[hi] if hi < len(offsets) else total)
                               for name, lo, hi in self._flat_children if hi > lo]
        self._flat = (flat, gflat)
        return self._flat

    def fuse_parameters(self):
        """[(parameters to lay out side by side, callback receiving the fused leaf tensor)] -- see flatten_parameters."""
        return []

    def flat_ranges(self) -> Dict[str, Tuple[int, int]]:
        """{top-level child name: [begin, end) of the flat parameter / gradient buffers} after flatten_parameters."""
        return {name: (a, b) for name, a, b in getattr(self, "_flat_children", [])}


def _rparam(in_size: int, shape, device, dtype, gen) -> Parameter:
    r = (2.0 / math.sqrt(in_size)) * (torch.rand(shape, generator=gen, dtype=torch.float32) - 0.5)
    return Parameter(r.to(device=device, dtype=dtype))


class Embedding(Module):
    def __init__(self, num_embeddings, embedding_dim, device, dtype, gen):
        self.weights = Parameter(torch.randn((num_embeddings, embedding_dim), generator=gen).to(device=device, dtype=dtype))

    def forward(self, idx):
        return ops.embedding(idx, self.weights.value)


class Linear(Module):
    def __init__(self, in_size, out_size, bias, device, dtype, gen):
        self.weights = _rparam(in_size, (in_size, out_size), device, dtype, gen)
        if bias:
            self.bias = _rparam(in_size, (out_size,), device, dtype, gen)
        self.has_bias = bool(bias)

    def forward(self, x, bias_grad: bool = True):
        return ops.linear(x, self.weights.value, self.bias.value if self.has_bias else None, bias_grad)

    def fused_bias_grad_ok(self, like, M: int) -> bool:
        """The LayerNorm that consumes this layer's (M, out) output may take over its bias gradient (ops.add_layernorm
        b_bias); `like`: any tensor on the device / of the dtype of the input."""
        return self.has_bias and ops.fused_bias_grad_ok(like, M, self.weights.value, self.bias.value)


class LayerNorm1d(Module):
    def __init__(self, dim, eps, device, dtype):
        self.eps = eps
        self.weights = Parameter(torch.ones(dim, device=device, dtype=dtype))
        self.bias = Parameter(torch.zeros(dim, device=device, dtype=dtype))

    def forward(self, x):
        return ops.layernorm(x, self.weights.value, self.bias.value, self.eps)

    def add_forward(self, a, b=None, b_bias=None):
        """(a + b, LayerNorm(a + b)) in one pass -- the residual sum and the LayerNorm that follows it (b may be None).
        b_bias: see ops.add_layernorm."""
        return ops.add_layernorm(a, b, self.weights.value, self.bias.value, self.eps, b_bias)


class MultiHeadAttention(Module):
    def __init__(self, n_embd, n_head, causal, bias, device, dtype, gen, attention_fn):
        self.n_embd, self.n_head, self.causal = n_embd, n_head, causal
        self.attn_hidden_dim = n_embd // n_head
        self.q_projection = Linear(n_embd, n_embd, bias, device, dtype, gen)
        self.k_projection = Linear(n_embd, n_embd, bias, device, dtype, gen)
        self.v_projection = Linear(n_embd, n_embd, bias, device, dtype, gen)
        self.out_projection = Linear(n_embd, n_embd, bias, device, dtype, gen)
        self.attention_fn = attention_fn

    def fuse_parameters(self):
        """q | k | v weights interleaved into one (in, 3 E) matrix, their biases side by side (flatten_parameters)."""
        projs = (self.q_projection, self.k_projection, self.v_projection)
        groups = [([p.weights for p in projs], lambda t: setattr(self, "_wqkv", t))]
        if all(p.has_bias for p in projs):
            groups.append(([p.bias for p in projs], lambda t: setattr(self, "_bqkv", t)))
        return groups

    def forward(self, x, bias_grad: bool = True):
        """bias_grad=False: the caller hands out_projection's bias to the add_layernorm that consumes the output."""
        B, T, E = x.shape
        H, d = self.n_head, self.attn_hidden_dim
        x2 = x.reshape(B * T, E)
        # flash path of project_to_query_key_value: K keeps the (B, H, T, d) layout (modules_transfomer.py:83-84).
        # ops.linear_to_heads / linear_from_heads write / read that layout directly when T % 128 == 0 and d % 32 == 0
        # (the GEMM's tensor maps permute); otherwise they are linear + permute + contiguous as in the reference.
        bias = lambda lin: lin.bias.value if lin.has_bias else None  # noqa: E731
        wqkv = getattr(self, "_wqkv", None)
        if wqkv is not None and ops.fused_qkv_enabled():
            # flattened parameters: the three projections are ONE GEMM over the interleaved weight -- straight into a
            # (3, B, H, T, d) tensor when the head-layout GEMM applies, else (B*T, 3E) followed by one permute copy
            bqkv = getattr(self, "_bqkv", None)
            if ops.heads_layout_ok(x2, T, d):
                qkv = ops.qkv_to_heads(x2, wqkv, bqkv, B, T, H, d)
            else:
                qkv = ops.linear(x2, wqkv, bqkv).view(B, T, 3, H, d).permute(2, 0, 3, 1, 4).contiguous()
            packed = getattr(self.attention_fn, "packed", None)
            if packed is not None:  # dq | dk | dv come back as one (3, B, H, T, d) gradient: no unbind / stack copies
                o = packed(qkv, self.causal)
                return ops.linear_from_heads(o, self.out_projection.weights.value, bias(self.out_projection), bias_grad).view(B, T, E)
            q, k, v = qkv.unbind(0)
        else:
            q = ops.linear_to_heads(x2, self.q_projection.weights.value, bias(self.q_projection), B, T, H, d)
            k = ops.linear_to_heads(x2, self.k_projection.weights.value, bias(self.k_projection), B, T, H, d)
            v = ops.linear_to_heads(x2, self.v_projection.weights.value, bias(self.v_projection), B, T, H, d)
        o = self.attention_fn(q, k, v, self.causal)
        return ops.linear_from_heads(o, self.out_projection.weights.value, bias(self.out_projection), bias_grad).view(B, T, E)


class FeedForward(Module):
    def __init__(self, n_embd, middle_dim, p_dropout, bias, device, dtype, gen):
        self.linear_in = Linear(n_embd, middle_dim, bias, device, dtype, gen)
        self.linear_out = Linear(middle_dim, n_embd, bias, device, dtype, gen)
        self.p_dropout = p_dropout

    def forward(self, x, bias_grad: bool = True):
        """bias_grad=False: the caller hands linear_out's bias to the add_layernorm that consumes the output."""
        B, T, E = x.shape
        x2 = x.reshape(B * T, E)
        fuse = self.linear_in.fused_bias_grad_ok(x2, B * T)  # linear_in's bias gradient comes out of the gelu backward pass
        h = ops.gelu_dropout(self.linear_in(x2, bias_grad=not fuse), self.p_dropout, self.training,
                             x_bias=self.linear_in.bias.value if fuse else None)
        return self.linear_out(h, bias_grad).view(B, T, E)


class TransformerLayer(Module):
    def __init__(self, n_embd, n_head, p_dropout, ln_eps, bias, device, dtype, gen, attention_fn):
        self.ln_1 = LayerNorm1d(n_embd, ln_eps, device, dtype)
        self.ln_2 = LayerNorm1d(n_embd, ln_eps, device, dtype)
        self.attention = MultiHeadAttention(n_embd, n_head, True, bias, device, dtype, gen, attention_fn)
        self.ff = FeedForward(n_embd, 256, p_dropout, bias, device, dtype, gen)

    def forward(self, x):
        s, f, _ = self.forward_pending(x, None, None, fuse_tail=False)
        return s + f

    def forward_pending(self, x, delta, delta_bias=None, fuse_tail: bool = True):
        """The layer applied to x + delta (delta may be None), returned as (s, f, f_bias) with output = s + f: every
        residual addition is left to the LayerNorm that follows it (ops.add_layernorm: one pass forward, one backward),
        the last one to the next layer's ln_1 or the model's final ln -- the caller passes (s, f, f_bias) on as
        (x, delta, delta_bias).  f_bias / delta_bias: the bias of the linear layer that produced f / delta when its
        gradient is left to that LayerNorm's backward kernel (only with fuse_tail), else None.
        (x + attn(ln_1(x)), then + ff(ln_2(.)): modules_transfomer.py TransformerLayer.forward.)"""
        B, T, E = x.shape
        M = B * T
        s1, h = self.ln_1.add_forward(x.reshape(M, E), None if delta is None else delta.reshape(M, E), delta_bias)
        outp = self.attention.out_projection
        fuse_a = outp.fused_bias_grad_ok(h, M)
        a = self.attention(h.view(B, T, E), bias_grad=not fuse_a)
        s2, h2 = self.ln_2.add_forward(s1, a.reshape(M, E), outp.bias.value if fuse_a else None)
        fuse_f = fuse_tail and self.ff.linear_out.fused_bias_grad_ok(h2, M)
        f = self.ff(h2.view(B, T, E), bias_grad=not fuse_f)
        return s2.view(B, T, E), f, (self.ff.linear_out.bias.value if fuse_f else None)


class DecoderLM(Module):
    def __init__(self, n_vocab: int, n_embd: int, n_head: int, n_positions: int, p_dropout: float = 0.1,
                 ln_eps: float = 1e-5, bias: bool = True, device="cuda", dtype=torch.float32, seed: int = 11111,
                 attention_fn: Optional[Callable] = None):
        gen = torch.Generator().manual_seed(seed)
        attention_fn = attention_fn or ops.flash_attention
        self.n_vocab, self.n_embd, self.n_positions, self.p_dropout = n_vocab, n_embd, n_positions, p_dropout
        self.token_embeddings = Embedding(n_vocab, n_embd, device, dtype, gen)
        self.position_embeddings = Embedding(n_vocab, n_embd, device, dtype, gen)  # sized n_vocab, as in the reference
        self.t_layer_1 = TransformerLayer(n_embd, n_head, p_dropout, ln_eps, bias, device, dtype, gen, attention_fn)
        self.t_layer_2 = TransformerLayer(n_embd, n_head, p_dropout, ln_eps, bias, device, dtype, gen, attention_fn)
        self.t_layer_3 = TransformerLayer(n_embd, n_head, p_dropout, ln_eps, bias, device, dtype, gen, attention_fn)
        self.t_layer_4 = TransformerLayer(n_embd, n_head, p_dropout, ln_eps, bias, device, dtype, gen, attention_fn)
        self.ln = LayerNorm1d(n_embd, ln_eps, device, dtype)
        self.lm_head = Linear(n_embd, n_vocab, bias, device, dtype, gen)

    def forward(self, idx, head_bias_grad: bool = True):
        """Logits (B, T, n_vocab).  head_bias_grad=False: the caller (loss_fn) hands lm_head's bias to ops.softmax_loss,
        whose backward kernel reduces its gradient."""
        B, T = idx.shape
        # dropout(token_embeddings(idx) + position_embeddings(arange(T))): one gather pass, one scatter pass back
        x = ops.embed_pos_dropout(idx, self.token_embeddings.weights.value, self.position_embeddings.weights.value,
                                  self.p_dropout, self.training)
        delta = delta_bias = None  # the residual stream is x + delta; each addition happens inside the next LayerNorm's kernel
        hook = getattr(self, "grad_ready_hook", None)
        layers = (self.t_layer_1, self.t_layer_2, self.t_layer_3, self.t_layer_4)
        if hook is not None and x.requires_grad:
            # gradient of the embedding stage's output: layer 1 is done, only the embedding scatter-add is left
            x.register_hook(lambda g, b=len(layers): hook(b))
        for k, layer in enumerate(layers):
            x, delta, delta_bias = layer.forward_pending(x, delta, delta_bias)
            if hook is not None and x.requires_grad:
                # fires when the gradient of layer k's output exists, i.e. when everything downstream of it -- layers
                # k+1.., the final LayerNorm and lm_head -- has finished its backward: the trainer starts the all-reduce of
                # that part's gradients (bucket `len(layers) - 1 - k`, see train.gradient_buckets) while the rest runs
                x.register_hook(lambda g, b=len(layers) - 1 - k: hook(b))
        _, x = self.ln.add_forward(x.reshape(B * T, self.n_embd), delta.reshape(B * T, self.n_embd), delta_bias)
        return self.lm_head(x, head_bias_grad).view(B, T, self.n_vocab)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {n: p.value.detach() for n, p in self.named_parameters()}


def loss_fn(batch, model: DecoderLM):
    """MLE loss averaged over all positions (project/run_machine_translation.py:158-177)."""
    ids, labels = batch["input_ids"], batch["labels"]
    head = model.lm_head
    fuse = head.has_bias and ops.fused_logits_bias_grad_ok(head.weights.value, ids.numel(), head.weights.value, head.bias.value, labels)
    logits = model(ids, head_bias_grad=not fuse)
    B, T, V = logits.shape
    loss = ops.softmax_loss(logits.view(B * T, V), labels.view(B * T), head.bias.value if fuse else None)
    return (loss * batch["label_token_weights"].view(B * T)).mean()
