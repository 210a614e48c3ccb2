"""Device-resident calls (fa_b200_fwd / fa_b200_bwd) on torch CUDA tensors.

torch is used only as the owner of device memory and streams; the arithmetic happens in
libfa_b200.so.  Tensors are (BH, T, d) or (B, H, T, d), contiguous, fp32 or bf16."""
from __future__ import annotations

import math

from . import _lib


def _prep(t):
    import torch
    if not t.is_cuda:
        raise ValueError("device API needs CUDA tensors (use the host-pointer API for host buffers)")
    if not t.is_contiguous():
        raise ValueError("tensors must be contiguous (BH, T, d)")
    if t.dtype == torch.bfloat16:
        return _lib.FA_DTYPE_BF16
    if t.dtype == torch.float32:
        return _lib.FA_DTYPE_F32
    raise ValueError(f"unsupported dtype {t.dtype}")


def _dims(q):
    *lead, T, d = q.shape
    return int(math.prod(lead)) if lead else 1, int(T), int(d)


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def flash_attn_fwd(q, k, v, causal: bool, precision: int = _lib.FA_PREC_AUTO, want_lm: bool = False, out=None,
                   p_drop: float = 0.0, seed: int = 0):
    """Returns (o, lse) or (o, lse, m, l) -- lse/m/l are fp32 with shape q.shape[:-1].
    `out=(o, lse)` reuses preallocated outputs.  `p_drop > 0` fuses attention dropout (mask = f(seed, head, i, j))."""
    import torch
    lib = _lib.load_library()
    dt = _prep(q)
    assert _prep(k) == dt and _prep(v) == dt and k.shape == q.shape and v.shape == q.shape
    BH, T, d = _dims(q)
    if out is not None:
        o, lse = out
    else:
        o = torch.empty_like(q)
        lse = torch.empty(q.shape[:-1], dtype=torch.float32, device=q.device)
    if p_drop > 0.0:
        assert not want_lm, "the FA-1 (l, m) statistics are not returned by the dropout entry point"
        _lib.check(lib.fa_b200_fwd_dropout(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), dt,
                                           precision, BH, T, d, 1 if causal else 0, float(p_drop), int(seed), _stream()))
        return o, lse
    m = torch.empty_like(lse) if want_lm else None
    l = torch.empty_like(lse) if want_lm else None
    _lib.check(lib.fa_b200_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
                               m.data_ptr() if want_lm else None, l.data_ptr() if want_lm else None,
                               dt, precision, BH, T, d, 1 if causal else 0, _stream()))
    return (o, lse, m, l) if want_lm else (o, lse)


def bwd_workspace(q, precision: int = _lib.FA_PREC_AUTO):
    import torch
    lib = _lib.load_library()
    BH, T, d = _dims(q)
    nbytes = int(lib.fa_b200_bwd_workspace_bytes(_prep(q), precision, BH, T, d))
    return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=q.device)


def flash_attn_bwd(q, k, v, o, d_out, lse, causal: bool, precision: int = _lib.FA_PREC_AUTO, workspace=None,
                   out=None, p_drop: float = 0.0, seed: int = 0):
    """Returns (dq, dk, dv) in the dtype of q.  (p_drop, seed) must be those of the forward call."""
    import torch
    lib = _lib.load_library()
    dt = _prep(q)
    for t in (k, v, o, d_out):
        assert _prep(t) == dt and t.shape == q.shape
    BH, T, d = _dims(q)
    if workspace is None:
        workspace = bwd_workspace(q, precision)
    dq, dk, dv = out if out is not None else (torch.empty_like(q), torch.empty_like(q), torch.empty_like(q))
    if p_drop > 0.0:
        _lib.check(lib.fa_b200_bwd_dropout(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_out.data_ptr(),
                                           lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dt, precision, BH,
                                           T, d, 1 if causal else 0, 0, float(p_drop), int(seed), workspace.data_ptr(),
                                           workspace.numel(), _stream()))
        return dq, dk, dv
    _lib.check(lib.fa_b200_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_out.data_ptr(),
                               lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dt, precision, BH, T, d,
                               1 if causal else 0, 0, workspace.data_ptr(), workspace.numel(), _stream()))
    return dq, dk, dv


def dropout_mask(n_heads: int, T: int, p_drop: float, seed: int, device="cuda", head_begin: int = 0):
    """keep[n_heads, T, T] (uint8, 1 = kept) for heads [head_begin, head_begin + n_heads) exactly as the kernels
    draw it -- for parity checks against the oracle."""
    import torch
    lib = _lib.load_library()
    keep = torch.empty((n_heads, T, T), dtype=torch.uint8, device=device)
    _lib.check(lib.fa_b200_dropout_mask(keep.data_ptr(), head_begin, n_heads, T, float(p_drop), int(seed), _stream()))
    return keep


def dropout_keep_scale(p_drop: float) -> float:
    return float(_lib.load_library().fa_b200_dropout_keep_scale(float(p_drop)))
