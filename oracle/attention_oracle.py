"""CPU restatement of the reference's attention path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker.  The product (flash-attention-minitorch_b200/) never does.

Parity status: PINNED.  The reference ships no golden vectors; this oracle is pinned
against outputs of the reference itself, generated in the build container by oracle/gen_golden.py
(which imports /root/reference/minitorch and kernel_tests/flash_attn_python.py) and committed under
tests/golden/.  tests/test_oracle_golden.py re-checks the pin on every run.

What is restated (reference file:line):
  * vanilla attention  softmax((q @ kT)/sqrt(d) + M, dim=3) @ v     minitorch/modules_transfomer.py:121-127
      - causal mask M = -FLT_MAX * triu(ones, 1)                     minitorch/modules_transfomer.py:59-65
      - softmax = exp(x - max) / sum                                 minitorch/nn.py:105-123
      - backward of matmul / softmax chain (autodiff)                minitorch/tensor_functions.py:405-423
  * FA-1 tiled forward, statistics l = sum exp(s - m), m = rowmax   kernel_tests/flash_attn_python.py:16-56,
                                                                     src/flash_attn_fw.cu:146-284
  * FA-2 tiled forward, statistic L = m + log(l)                     kernel_tests/flash_attn_python.py:59-98,
                                                                     src/flash_attn2_fw.cu:152-294
  * FA-1 / FA-2 backward from saved statistics                       kernel_tests/flash_attn_python.py:100-192,
                                                                     src/flash_attn_bw.cu:143-245, src/flash_attn2_bw.cu:143-245
  * in-kernel causal sentinel -1e7 for masked scores                 src/flash_attn_fw.cu:152-159

All functions take numpy arrays shaped (BH, T, d) -- the reference flattens (B, H) -> B*H before
calling its kernels (minitorch/cuda_kernel_ops.py:542-547) -- and compute in the dtype given
(`dtype=np.float32` follows the reference's arithmetic, `np.float64` is the high-precision check).
"""
from __future__ import annotations

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)


def causal_mask_additive(T: int, dtype=np.float32) -> np.ndarray:
    """-FLT_MAX above the diagonal, 0 elsewhere (minitorch/modules_transfomer.py:63)."""
    return (-FLT_MAX * np.triu(np.ones((T, T), dtype=np.float64), 1)).astype(dtype)


def softmax_lastdim(x: np.ndarray) -> np.ndarray:
    """exp(x - max) / sum(exp(x - max))   (minitorch/nn.py:105-123)."""
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def attention_forward(q, k, v, causal: bool, dtype=np.float64):
    """Vanilla attention forward.  Returns (O, P, m, l, L):
    P  (BH,T,T) probabilities, m row max of the scaled scores, l = sum exp(s - m), L = m + log l."""
    q = np.asarray(q, dtype=dtype)
    k = np.asarray(k, dtype=dtype)
    v = np.asarray(v, dtype=dtype)
    BH, T, d = q.shape
    s = (q @ k.transpose(0, 2, 1)) / dtype(np.sqrt(d))
    if causal:
        # adding -FLT_MAX to an O(1) score leaves exactly -FLT_MAX in fp32; exp() of it is exactly 0
        s = np.where(np.triu(np.ones((T, T), dtype=bool), 1)[None], dtype(-FLT_MAX) if dtype == np.float32 else -np.inf, s)
    m = s.max(axis=-1)
    e = np.exp(s - m[..., None])
    l = e.sum(axis=-1)
    p = e / l[..., None]
    o = p @ v
    return o, p, m, l, m + np.log(l)


def attention_backward(q, k, v, d_out, causal: bool, dtype=np.float64):
    """Gradients of sum(O * d_out) w.r.t. q, k, v for vanilla attention.  Returns (dQ, dK, dV, O, L)."""
    q = np.asarray(q, dtype=dtype)
    k = np.asarray(k, dtype=dtype)
    v = np.asarray(v, dtype=dtype)
    g = np.asarray(d_out, dtype=dtype)
    d = q.shape[-1]
    tau = dtype(1.0 / np.sqrt(d))
    o, p, m, l, L = attention_forward(q, k, v, causal, dtype)
    dv = p.transpose(0, 2, 1) @ g
    dp = g @ v.transpose(0, 2, 1)
    delta = (g * o).sum(axis=-1)
    ds = p * (dp - delta[..., None])
    dq = tau * (ds @ k)
    dk = tau * (ds.transpose(0, 2, 1) @ q)
    return dq, dk, dv, o, L


# ----------------------------------------------------------------------------------------------
# Attention dropout (not on the reference's flash path; convention of minitorch/modules_basic.py:77-80:
# mask ~ Bernoulli(1 - p), x * mask * 1/(1 - p)), applied to P after the softmax normalisation.
# dropout_keep_mask restates flash-attention-minitorch_b200/csrc/fa_dropout.cuh bit for bit.
# ----------------------------------------------------------------------------------------------
def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def dropout_threshold(p: float) -> int:
    """7-bit threshold the kernels use: keep iff (byte & 0x7f) >= thr; p_eff = thr / 128."""
    return 0 if not p > 0 else min(127, int(np.float32(p) * np.float32(128.0) + np.float32(0.5)))


def dropout_keep_scale(p: float) -> float:
    return float(np.float32(128.0) / (np.float32(128.0) - np.float32(dropout_threshold(p))))


def dropout_keep_mask(seed: int, heads, T: int, p: float) -> np.ndarray:
    """keep[len(heads), T, T] (bool) for the given head indices (b*H + h)."""
    heads = np.atleast_1d(np.asarray(heads, dtype=np.uint32))
    with np.errstate(over="ignore"):
        lo, hi = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
        hkey = _mix32(lo ^ _mix32(hi + np.uint32(0x9E3779B9) * (heads + np.uint32(1))))
        i = np.arange(T, dtype=np.uint32)
        rterm = (i >> np.uint32(1)) * np.uint32(0x9E3779B1)
        cterm = (i >> np.uint32(1)) * np.uint32(0x85EBCA77)
        rkey = _mix32(hkey[:, None] ^ rterm[None, :])  # two rounds: the row term is mixed before the column term comes in
        bits = _mix32(rkey[:, :, None] ^ cterm[None, None, :])
        sh = np.uint32(8) * ((i & np.uint32(1))[:, None] * np.uint32(2) + (i & np.uint32(1))[None, :])
        byte = (bits >> sh[None]) & np.uint32(0x7F)
    return byte >= np.uint32(dropout_threshold(p))


def attention_dropout_backward(q, k, v, d_out, causal: bool, keep: np.ndarray, keep_scale: float, dtype=np.float64):
    """Vanilla attention with dropout Z = keep * keep_scale on P.  Returns (dQ, dK, dV, O, L); L is the logsumexp
    of the undropped scores.  O = (P o Z) V, dV = (P o Z)^T dO, dP = (dO V^T) o Z, dS = P o (dP - rowsum(dO o O))."""
    q = np.asarray(q, dtype=dtype)
    k = np.asarray(k, dtype=dtype)
    v = np.asarray(v, dtype=dtype)
    g = np.asarray(d_out, dtype=dtype)
    tau = dtype(1.0 / np.sqrt(q.shape[-1]))
    _, p, m, l, L = attention_forward(q, k, v, causal, dtype)
    z = keep.astype(dtype) * dtype(keep_scale)
    pz = p * z
    o = pz @ v
    dv = pz.transpose(0, 2, 1) @ g
    dp = (g @ v.transpose(0, 2, 1)) * z
    delta = (g * o).sum(axis=-1)
    ds = p * (dp - delta[..., None])
    return tau * (ds @ k), tau * (ds.transpose(0, 2, 1) @ q), dv, o, L


# ----------------------------------------------------------------------------------------------
# Tiled restatements of the reference's flash kernels (slow: pure numpy over tiles; small cases only)
# ----------------------------------------------------------------------------------------------
def _scores_tile(Qi, Kj, i0, j0, tau, causal, dtype):
    s = tau * (Qi @ Kj.T)
    if causal:
        rows = i0 + np.arange(Qi.shape[0])[:, None]
        cols = j0 + np.arange(Kj.shape[0])[None, :]
        s = np.where(cols <= rows, s, dtype(-1e7))  # src/flash_attn_fw.cu:152-159
    return s


def flash1_forward(q, k, v, causal: bool, Br: int = 16, Bc: int = 16, dtype=np.float32):
    """FA-1 forward: K/V-outer, Q-inner; O, l, m merged in place (flash_attn_python.py:16-56).
    Returns (O, l, m) with l = sum exp(s - m), m = row max."""
    q = np.asarray(q, dtype=dtype)
    k = np.asarray(k, dtype=dtype)
    v = np.asarray(v, dtype=dtype)
    BH, T, d = q.shape
    tau = dtype(np.sqrt(1.0 / d))
    O = np.zeros_like(q)
    l = np.zeros((BH, T), dtype=dtype)
    m = np.full((BH, T), -FLT_MAX, dtype=dtype)  # caller-provided initial state, cuda_kernel_ops.py:537-539
    for b in range(BH):
        for j0 in range(0, T, Bc):
            Kj, Vj = k[b, j0:j0 + Bc], v[b, j0:j0 + Bc]
            for i0 in range(0, T, Br):
                if causal and j0 > i0 + Br - 1:
                    continue  # -DCAUSAL_BLOCKSPARSE tile skip, src/flash_attn_fw.cu:88-92 (results identical)
                sl = slice(i0, i0 + Br)
                S = _scores_tile(q[b, sl], Kj, i0, j0, tau, causal, dtype)
                m_ij = S.max(axis=1)
                P = np.exp(S - m_ij[:, None])
                l_ij = P.sum(axis=1)
                m_new = np.maximum(m[b, sl], m_ij)
                a = np.exp(m[b, sl] - m_new) * l[b, sl]
                c = np.exp(m_ij - m_new)
                l_new = a + c * l_ij
                O[b, sl] = (a[:, None] * O[b, sl] + c[:, None] * (P @ Vj)) / l_new[:, None]
                m[b, sl] = m_new
                l[b, sl] = l_new
    return O, l, m


def flash2_forward(q, k, v, causal: bool, Br: int = 16, Bc: int = 16, dtype=np.float32):
    """FA-2 forward: Q-outer, K/V-inner, running (m, l); writes L = m + log l (flash_attn_python.py:59-98)."""
    q = np.asarray(q, dtype=dtype)
    k = np.asarray(k, dtype=dtype)
    v = np.asarray(v, dtype=dtype)
    BH, T, d = q.shape
    tau = dtype(np.sqrt(1.0 / d))
    O = np.zeros_like(q)
    L = np.zeros((BH, T), dtype=dtype)
    for b in range(BH):
        for i0 in range(0, T, Br):
            sl = slice(i0, i0 + Br)
            Qi = q[b, sl]
            Oi = np.zeros_like(Qi)
            li = np.zeros(Qi.shape[0], dtype=dtype)
            mi = np.full(Qi.shape[0], -np.inf, dtype=dtype)
            for j0 in range(0, T, Bc):
                if causal and j0 > i0 + Br - 1:
                    continue
                S = _scores_tile(Qi, k[b, j0:j0 + Bc], i0, j0, tau, causal, dtype)
                m_prev = mi
                mi = np.maximum(mi, S.max(axis=1))
                P = np.exp(S - mi[:, None])
                corr = np.exp(m_prev - mi)
                li = corr * li + P.sum(axis=1)
                Oi = corr[:, None] * Oi + P @ v[b, j0:j0 + Bc]
            O[b, sl] = Oi / li[:, None]
            L[b, sl] = mi + np.log(li)
    return O, L


def flash_backward(q, k, v, o, d_out, causal: bool, *, l=None, m=None, L=None, Br: int = 4, Bc: int = 4, dtype=np.float32):
    """FA-1 (given l, m: P = exp(S - m)/l) or FA-2 (given L: P = exp(S - L)) backward, K/V-outer
    (flash_attn_python.py:100-192; src/flash_attn_bw.cu:166, src/flash_attn2_bw.cu:168)."""
    q = np.asarray(q, dtype=dtype)
    k = np.asarray(k, dtype=dtype)
    v = np.asarray(v, dtype=dtype)
    o = np.asarray(o, dtype=dtype)
    g = np.asarray(d_out, dtype=dtype)
    BH, T, d = q.shape
    tau = dtype(np.sqrt(1.0 / d))
    dq, dk, dv = np.zeros_like(q), np.zeros_like(q), np.zeros_like(q)
    for b in range(BH):
        for j0 in range(0, T, Bc):
            js = slice(j0, j0 + Bc)
            for i0 in range(0, T, Br):
                if causal and j0 > i0 + Br - 1:
                    continue
                sl = slice(i0, i0 + Br)
                S = _scores_tile(q[b, sl], k[b, js], i0, j0, tau, causal, dtype)
                if L is not None:
                    P = np.exp(S - np.asarray(L, dtype=dtype)[b, sl][:, None])
                else:
                    P = np.exp(S - np.asarray(m, dtype=dtype)[b, sl][:, None]) / np.asarray(l, dtype=dtype)[b, sl][:, None]
                dv[b, js] += P.T @ g[b, sl]
                dP = g[b, sl] @ v[b, js].T
                D = (g[b, sl] * o[b, sl]).sum(axis=1)
                dS = P * (dP - D[:, None])
                dq[b, sl] += tau * (dS @ k[b, js])
                dk[b, js] += tau * (dS.T @ q[b, sl])
    return dq, dk, dv


# ----------------------------------------------------------------------------------------------
# helpers shared by tests and bench
# ----------------------------------------------------------------------------------------------
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (the storage rounding of the bf16 path)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def make_inputs(BH: int, T: int, d: int, seed: int, grad_seed: int | None = None, ones_grad: bool = False):
    """U(-1,1) inputs as the reference's kernel tests draw them (test_utils.py:104-105), seeded."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(-1, 1, (BH, T, d)).astype(np.float32)
    k = rng.uniform(-1, 1, (BH, T, d)).astype(np.float32)
    v = rng.uniform(-1, 1, (BH, T, d)).astype(np.float32)
    if ones_grad:
        g = np.ones((BH, T, d), dtype=np.float32)  # kernel_tests/test_flashattn_bw.py:32
    else:
        g = np.random.default_rng(seed + 1 if grad_seed is None else grad_seed).uniform(-1, 1, (BH, T, d)).astype(np.float32)
    return q, k, v, g


def attention_flops(BH: int, T: int, d: int, causal: bool, backward: bool = True, forward: bool = True) -> float:
    """F_fwd = 4*BH*T^2*d, F_bwd = 10*BH*T^2*d, causal = exactly half."""
    f = (4.0 if forward else 0.0) + (10.0 if backward else 0.0)
    return f * BH * float(T) * T * d * (0.5 if causal else 1.0)
