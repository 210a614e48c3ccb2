"""Host-side mirror of minitorch's flash-attention operator surface, backed by the B200 library.

The reference is pure Python over host numpy float32 storage; on a GPU box the reference package
itself is not available (it cannot be vendored), so this module re-states -- with the same names,
argument meaning and error behaviour -- exactly the slice of minitorch the flash path touches:

    Tensor.flash_attn / flash_attn_causal / flash_attn2(k, v, causal_mask=False)   minitorch/tensor.py:422-429
    Flash_Attn / Flash_Attn2 / Flash_Attn_Causal (Function.apply, Context)         minitorch/tensor_functions.py:46-68, 462-497
    TensorBackend flash wiring                                                      minitorch/tensor_ops.py:102-108
    CudaKernelOps.flash_attn{,2,_causal}_{fw,bw} and the two *_generic marshallers  minitorch/cuda_kernel_ops.py:527-677
    tensor(), tensor_from_numpy(), Tensor.backward()/grad                           minitorch/tensor_functions.py:578-633

It is NOT a re-implementation of minitorch (no map/zip/reduce, no general autodiff): the only
differentiable ops are the three flash functions plus `sum`, enough for the reference's kernel
tests (`out.sum().backward()`, kernel_tests/test_flashattn_bw.py) to be written the same way.

With the real minitorch no Python changes are needed at all: point its cwd-relative loader at the
six libraries in flash-attention-minitorch_b200/minitorch/cuda_kernels/ (see INTEGRATION.md).
"""
from __future__ import annotations

import ctypes
from typing import Any, Optional, Sequence, Tuple

import numpy as np

from . import _lib

datatype = np.float32


# ----------------------------------------------------------------------------------------------
# storage + tensor
# ----------------------------------------------------------------------------------------------
class TensorData:
    """Flat float32 storage + shape + strides (element units), like minitorch/tensor_data.py."""

    def __init__(self, storage: np.ndarray, shape: Sequence[int], strides: Optional[Sequence[int]] = None):
        self._storage = np.ascontiguousarray(storage, dtype=datatype).reshape(-1)
        self.shape = tuple(int(s) for s in shape)
        if strides is None:
            strides, acc = [], 1
            for s in reversed(self.shape):
                strides.append(acc)
                acc *= s
            strides = tuple(reversed(strides))
        self._strides = np.array(strides, dtype=np.int64)
        self._shape = np.array(self.shape, dtype=np.int64)
        assert int(np.prod(self.shape, dtype=np.int64)) == self._storage.size or len(self.shape) == 0

    def is_contiguous(self) -> bool:
        last = 10 ** 18
        for s in self._strides:
            if s > last:
                return False
            last = s
        return True

    def to_numpy(self) -> np.ndarray:
        if len(self.shape) == 0:
            return self._storage[:1].reshape(())
        return np.lib.stride_tricks.as_strided(self._storage, self.shape, tuple(int(s) * 4 for s in self._strides)).copy()


class Context:
    def __init__(self, no_grad: bool = False):
        self.no_grad = no_grad
        self.saved_values: Tuple[Any, ...] = ()

    def save_for_backward(self, *values: Any) -> None:
        if not self.no_grad:
            self.saved_values = values


class History:
    def __init__(self, last_fn=None, ctx=None, inputs: Sequence["Tensor"] = ()):
        self.last_fn, self.ctx, self.inputs = last_fn, ctx, inputs


class Tensor:
    def __init__(self, v: TensorData, back: Optional[History] = None, backend: Optional["TensorBackend"] = None):
        self._tensor = v
        self.history = back
        self.backend = backend if backend is not None else default_backend()
        self.f = self.backend
        self.grad: Optional[Tensor] = None

    # -- structure --------------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return self._tensor.shape

    @property
    def dims(self) -> int:
        return len(self._tensor.shape)

    def to_numpy(self) -> np.ndarray:
        return self._tensor.to_numpy()

    def requires_grad_(self, x: bool) -> None:
        self.history = History() if x else None

    def requires_grad(self) -> bool:
        return self.history is not None

    def detach(self) -> "Tensor":
        return Tensor(self._tensor, backend=self.backend)

    def is_leaf(self) -> bool:
        return self.history is not None and self.history.last_fn is None

    def contiguous(self) -> "Tensor":
        if self._tensor.is_contiguous():
            return self
        return Tensor(TensorData(self.to_numpy(), self.shape), backend=self.backend)

    def view(self, *shape: int) -> "Tensor":
        assert self._tensor.is_contiguous(), "Must be contiguous to view"
        shape = tuple(int(s) for s in (shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape))
        return Tensor(TensorData(self._tensor._storage, shape), back=None, backend=self.backend)

    def permute(self, *order: int) -> "Tensor":
        td = self._tensor
        new = TensorData(td._storage, tuple(td.shape[o] for o in order), tuple(int(td._strides[o]) for o in order))
        return Tensor(new, backend=self.backend)

    def zeros(self, shape: Optional[Sequence[int]] = None) -> "Tensor":
        shape = self.shape if shape is None else tuple(shape)
        return Tensor(TensorData(np.zeros(int(np.prod(shape, dtype=np.int64)), dtype=datatype), shape), backend=self.backend)

    def item(self) -> float:
        return float(self._tensor._storage[0])

    # -- the flash-attention surface (minitorch/tensor.py:422-429) ---------------------------
    def flash_attn(self, k: "Tensor", v: "Tensor", causal_mask: bool = False) -> "Tensor":
        return Flash_Attn.apply(self, k, v, tensor(causal_mask))

    def flash_attn_causal(self, k: "Tensor", v: "Tensor", causal_mask: bool = False) -> "Tensor":
        return Flash_Attn_Causal.apply(self, k, v, tensor(causal_mask))

    def flash_attn2(self, k: "Tensor", v: "Tensor", causal_mask: bool = False) -> "Tensor":
        return Flash_Attn2.apply(self, k, v, tensor(causal_mask))

    # -- just enough autodiff for `out.sum().backward()` / `out.backward(grad)` ---------------
    def sum(self) -> "Tensor":
        return Sum.apply(self)

    def backward(self, grad_output: Optional["Tensor"] = None) -> None:
        if grad_output is None:
            assert self.shape in ((1,), ()), "Must provide grad_output if non-scalar"
            grad_output = Tensor(TensorData(np.ones(1, dtype=datatype), (1,)), backend=self.backend)
        _backpropagate(self, grad_output)


class Function:
    @classmethod
    def apply(cls, *vals: Tensor) -> Tensor:
        need_grad = any(v.requires_grad() for v in vals)
        raw_vals = [v.detach() for v in vals]
        ctx = Context(not need_grad)
        c = cls.forward(ctx, *raw_vals)
        back = History(cls, ctx, vals) if need_grad else None
        return Tensor(c._tensor, back, backend=c.backend)


def _backpropagate(t: Tensor, grad: Tensor) -> None:
    """Chain rule for the (tree-shaped) graphs the flash tests build."""
    if t.history is None:
        return
    if t.is_leaf():
        if t.grad is None:
            t.grad = grad
        else:
            t.grad = Tensor(TensorData(t.grad.to_numpy() + grad.to_numpy(), t.shape), backend=t.backend)
        return
    h = t.history
    grads = h.last_fn.backward(h.ctx, grad)
    if not isinstance(grads, tuple):
        grads = (grads,)
    for inp, g in zip(h.inputs, grads):
        if inp.requires_grad():
            _backpropagate(inp, g)


class Sum(Function):
    @staticmethod
    def forward(ctx: Context, a: Tensor) -> Tensor:
        ctx.save_for_backward(a.shape, a.backend)
        return Tensor(TensorData(np.array([a.to_numpy().sum(dtype=np.float64)], dtype=datatype), (1,)), backend=a.backend)

    @staticmethod
    def backward(ctx: Context, grad_output: Tensor):
        shape, backend = ctx.saved_values
        g = np.full(shape, grad_output.item(), dtype=datatype)
        return Tensor(TensorData(g, shape), backend=backend)


# minitorch/tensor_functions.py:462-497, verbatim structure
class Flash_Attn(Function):
    @staticmethod
    def forward(ctx: Context, q: Tensor, k: Tensor, v: Tensor, causal_mask: Tensor) -> Tensor:
        o, l, m = q.f.flash_attn_fw(q, k, v, causal_mask)
        ctx.save_for_backward(q, k, v, o, l, m, causal_mask)
        return o

    @staticmethod
    def backward(ctx: Context, out_grad: Tensor):
        q, k, v, o, l, m, causal_mask = ctx.saved_values
        return out_grad.f.flash_attn_bw(q, k, v, o, out_grad, l, m, causal_mask)


class Flash_Attn2(Function):
    @staticmethod
    def forward(ctx: Context, q: Tensor, k: Tensor, v: Tensor, causal_mask: Tensor) -> Tensor:
        o, l, m = q.f.flash_attn2_fw(q, k, v, causal_mask)
        ctx.save_for_backward(q, k, v, o, l, m, causal_mask)
        return o

    @staticmethod
    def backward(ctx: Context, out_grad: Tensor):
        q, k, v, o, l, m, causal_mask = ctx.saved_values
        return out_grad.f.flash_attn2_bw(q, k, v, o, out_grad, l, m, causal_mask)


class Flash_Attn_Causal(Function):
    @staticmethod
    def forward(ctx: Context, q: Tensor, k: Tensor, v: Tensor, causal_mask: Tensor) -> Tensor:
        o, l, m = q.f.flash_attn_causal_fw(q, k, v, causal_mask)
        ctx.save_for_backward(q, k, v, o, l, m, causal_mask)
        return o

    @staticmethod
    def backward(ctx: Context, out_grad: Tensor):
        q, k, v, o, l, m, causal_mask = ctx.saved_values
        return out_grad.f.flash_attn_causal_bw(q, k, v, o, out_grad, l, m, causal_mask)


# ----------------------------------------------------------------------------------------------
# the ctypes boundary (minitorch/cuda_kernel_ops.py:26-35, 527-677)
# ----------------------------------------------------------------------------------------------
_shims: dict = {}


def _shim(name: str) -> ctypes.CDLL:
    """ctypes.CDLL of one reference-named library.  The reference loads them cwd-relative at import
    (cuda_kernel_ops.py:30-35); here the path is resolved next to this package and loading is lazy."""
    if name not in _shims:
        import os
        path = _lib.shim_path(name)
        if not os.path.exists(path):
            raise _lib.LibraryMissingError(f"{path} not found: run flash-attention-minitorch_b200/compile_cuda.sh")
        _lib.load_library()  # make the dependency resident first (also validates it exists)
        _shims[name] = ctypes.CDLL(path)
    return _shims[name]


def _current_stream() -> int:
    """The reference passes torch.cuda.current_stream().cuda_stream (cuda_kernel_ops.py:535); without
    torch the default stream (0) is the same thing."""
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_stream().cuda_stream)
    except Exception:
        pass
    return 0


_f32_1d = np.ctypeslib.ndpointer(dtype=datatype, ndim=1, flags="C_CONTIGUOUS")


class CudaKernelOps:
    @staticmethod
    def flash_attn_fw_generic(q: Tensor, k: Tensor, v: Tensor, causal_mask: Tensor, generic_lib):
        causal_mask = int(causal_mask._tensor._storage.item()) == 1
        batch_size, nhead, from_len, to_len = q.shape  # NB: `to_len` is the head dim
        assert q.shape == k.shape
        assert q.shape == v.shape
        assert (q._tensor._strides - k._tensor._strides).sum() == 0
        assert (q._tensor._strides - v._tensor._strides).sum() == 0
        stream = _current_stream()

        out = q.zeros(q.shape)
        l = k.zeros((batch_size, nhead, from_len))
        m = tensor_from_numpy(np.ones(l.shape) * (-np.finfo(datatype).max), backend=q.backend)

        q = q.contiguous().view(int(np.prod(q.shape[:-2])), q.shape[-2], q.shape[-1])
        k = k.contiguous().view(int(np.prod(k.shape[:-2])), k.shape[-2], k.shape[-1])
        v = v.contiguous().view(int(np.prod(v.shape[:-2])), v.shape[-2], v.shape[-1])
        out = out.contiguous().view(int(np.prod(out.shape[:-2])), out.shape[-2], out.shape[-1])

        generic_lib.launch_flash_attn_fw.argtypes = [_f32_1d] * 6 + [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                                    ctypes.c_bool, ctypes.c_void_p]
        generic_lib.launch_flash_attn_fw.restype = None
        generic_lib.launch_flash_attn_fw(q._tensor._storage, k._tensor._storage, v._tensor._storage,
                                         out._tensor._storage, l._tensor._storage, m._tensor._storage,
                                         batch_size * nhead, from_len, to_len, causal_mask, stream)
        out = out.view(batch_size, nhead, from_len, to_len)
        return out, l, m

    @staticmethod
    def flash_attn_bw_generic(q: Tensor, k: Tensor, v: Tensor, out: Tensor, out_grad: Tensor, l: Tensor, m: Tensor,
                              causal_mask: Tensor, generic_lib):
        causal_mask_ = int(causal_mask._tensor._storage.item()) == 1
        batch_size, nhead, from_len, to_len = q.shape
        assert q.shape == k.shape
        assert q.shape == v.shape
        assert q.shape == out.shape
        assert q.shape == out_grad.shape
        assert l.shape == (batch_size, nhead, from_len)
        assert m.shape == (batch_size, nhead, from_len)
        stream = _current_stream()

        flat = lambda t: t.contiguous().view(int(np.prod(t.shape[:-2])), t.shape[-2], t.shape[-1])  # noqa: E731
        q, k, v, out, out_grad = flat(q), flat(k), flat(v), flat(out), flat(out_grad)
        l, m = l.contiguous(), m.contiguous()
        q_grad, k_grad, v_grad = q.zeros(q.shape), k.zeros(k.shape), v.zeros(v.shape)

        generic_lib.launch_flash_attn_bw.argtypes = [_f32_1d] * 10 + [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                                     ctypes.c_bool, ctypes.c_void_p]
        generic_lib.launch_flash_attn_bw.restype = None
        generic_lib.launch_flash_attn_bw(q._tensor._storage, k._tensor._storage, v._tensor._storage,
                                         out._tensor._storage, out_grad._tensor._storage, q_grad._tensor._storage,
                                         k_grad._tensor._storage, v_grad._tensor._storage, l._tensor._storage,
                                         m._tensor._storage, batch_size * nhead, from_len, to_len, causal_mask_, stream)
        q_grad = q_grad.view(batch_size, nhead, from_len, to_len)
        k_grad = k_grad.view(batch_size, nhead, from_len, to_len)
        v_grad = v_grad.view(batch_size, nhead, from_len, to_len)
        return q_grad, k_grad, v_grad, causal_mask  # 4th value: dummy "grad" of the bool tensor (:653)

    @staticmethod
    def flash_attn_fw(q, k, v, causal_mask):
        return CudaKernelOps.flash_attn_fw_generic(q, k, v, causal_mask, _shim("flash_attn_fw"))

    @staticmethod
    def flash_attn_bw(q, k, v, out, out_grad, l, m, causal_mask):
        return CudaKernelOps.flash_attn_bw_generic(q, k, v, out, out_grad, l, m, causal_mask, _shim("flash_attn_bw"))

    @staticmethod
    def flash_attn2_fw(q, k, v, causal_mask):
        return CudaKernelOps.flash_attn_fw_generic(q, k, v, causal_mask, _shim("flash_attn2_fw"))

    @staticmethod
    def flash_attn2_bw(q, k, v, out, out_grad, l, m, causal_mask):
        return CudaKernelOps.flash_attn_bw_generic(q, k, v, out, out_grad, l, m, causal_mask, _shim("flash_attn2_bw"))

    @staticmethod
    def flash_attn_causal_fw(q, k, v, causal_mask):
        return CudaKernelOps.flash_attn_fw_generic(q, k, v, causal_mask, _shim("flash_attn_causal_fw"))

    @staticmethod
    def flash_attn_causal_bw(q, k, v, out, out_grad, l, m, causal_mask):
        return CudaKernelOps.flash_attn_bw_generic(q, k, v, out, out_grad, l, m, causal_mask,
                                                   _shim("flash_attn_causal_bw"))


class TensorBackend:
    """minitorch/tensor_ops.py:102-108: attribute wiring ops.X -> backend.X.  An `ops` class without
    the flash attributes raises AttributeError here, as in the reference."""

    def __init__(self, ops):
        self.flash_attn_fw = ops.flash_attn_fw
        self.flash_attn_bw = ops.flash_attn_bw
        self.flash_attn2_fw = ops.flash_attn2_fw
        self.flash_attn2_bw = ops.flash_attn2_bw
        self.flash_attn_causal_fw = ops.flash_attn_causal_fw
        self.flash_attn_causal_bw = ops.flash_attn_causal_bw
        self.cuda = True


_default_backend: Optional[TensorBackend] = None


def default_backend() -> TensorBackend:
    global _default_backend
    if _default_backend is None:
        _default_backend = TensorBackend(CudaKernelOps)
    return _default_backend


# ----------------------------------------------------------------------------------------------
# constructors (minitorch/tensor_functions.py:578-633)
# ----------------------------------------------------------------------------------------------
def tensor(ls: Any, backend: Optional[TensorBackend] = None, requires_grad: bool = False) -> Tensor:
    arr = np.array(ls, dtype=datatype)  # a python bool becomes a shape-() tensor holding 0./1.
    t = Tensor(TensorData(arr.reshape(-1), arr.shape), backend=backend)
    t.requires_grad_(requires_grad)
    return t


def tensor_from_numpy(ls: np.ndarray, backend: Optional[TensorBackend] = None, requires_grad: bool = False) -> Tensor:
    if ls.dtype != datatype:
        ls = ls.astype(datatype)
    res = Tensor(TensorData(ls.flatten(), ls.shape), backend=backend)  # flatten() copies, like the reference
    res.requires_grad_(requires_grad)
    return res
