"""fa_b200 -- host side of the B200 flash-attention drop-in for minitorch.

Layout (everything the flash-attention path needs and nothing else):
  _lib.py              ctypes binding of lib/libfa_b200.so (include/fa_b200.h); fails loudly if the
                       CUDA library has not been built -- there is no CPU fallback
  minitorch_compat.py  mirror of the reference operator surface: Tensor.flash_attn / flash_attn_causal /
                       flash_attn2, the Flash_Attn* Functions, CudaKernelOps.flash_attn*_{fw,bw},
                       TensorBackend wiring (reference minitorch/tensor.py:422-429,
                       tensor_functions.py:462-497, cuda_kernel_ops.py:527-677, tensor_ops.py:102-108)
  device.py            device-resident calls on raw device pointers / torch CUDA tensors
  sharding.py          batch x head partitioning across ranks (no collective on the attention path)
"""
from ._lib import load_library, library_path, LibraryMissingError  # noqa: F401
from . import sharding  # noqa: F401

__all__ = ["load_library", "library_path", "LibraryMissingError", "sharding"]
