"""ctypes binding of libfa_b200.so -- one function per declaration in include/fa_b200.h."""
from __future__ import annotations

import ctypes
import os

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # flash-attention-minitorch_b200/
LIB_DIR = os.path.join(PKG_ROOT, "lib")
SHIM_DIR = os.path.join(PKG_ROOT, "minitorch", "cuda_kernels")

# status codes / enums of include/fa_b200.h
FA_OK, FA_ERR_BAD_ARG, FA_ERR_CUDA, FA_ERR_WORKSPACE, FA_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
FA_DTYPE_F32, FA_DTYPE_BF16 = 0, 1
FA_PREC_AUTO, FA_PREC_FP32_SIMT, FA_PREC_TF32_TC, FA_PREC_BF16_TC = 0, 1, 2, 3
FA_CONV_FA1, FA_CONV_FA2 = 1, 2

SHIM_NAMES = ("flash_attn_fw", "flash_attn_bw", "flash_attn_causal_fw", "flash_attn_causal_bw",
              "flash_attn2_fw", "flash_attn2_bw")


class LibraryMissingError(RuntimeError):
    """Raised when the CUDA library is not built.  The product has no CPU fallback by design."""


class FaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fa_b200 error {code}: {msg}")
        self.code = code


def library_path() -> str:
    # FA_B200_LIB: developer override used for same-box A/B runs of two builds (tools/lm_bench.py, bench.py)
    return os.environ.get("FA_B200_LIB") or os.path.join(LIB_DIR, "libfa_b200.so")


def shim_path(name: str) -> str:
    return os.path.join(SHIM_DIR, name + ".so")


_lib = None


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise LibraryMissingError(
            f"{path} not found: build it with flash-attention-minitorch_b200/compile_cuda.sh "
            "(or __graft_entry__.build()); this package has no CPU fallback")
    lib = ctypes.CDLL(path)
    vp, f32p, ci, sz = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.fa_b200_version.restype = ci
    lib.fa_b200_last_error.restype = ctypes.c_char_p
    lib.fa_b200_launch_count.restype = ctypes.c_ulonglong
    lib.fa_b200_fwd.argtypes = [vp, vp, vp, vp, f32p, f32p, f32p, ci, ci, ci, ci, ci, ci, vp]
    lib.fa_b200_fwd.restype = ci
    lib.fa_b200_bwd_workspace_bytes.argtypes = [ci, ci, ci, ci, ci]
    lib.fa_b200_bwd_workspace_bytes.restype = sz
    lib.fa_b200_bwd.argtypes = [vp, vp, vp, vp, vp, f32p, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, sz, vp]
    lib.fa_b200_bwd.restype = ci
    u64, f32 = ctypes.c_ulonglong, ctypes.c_float
    lib.fa_b200_fwd_dropout.argtypes = [vp, vp, vp, vp, f32p, ci, ci, ci, ci, ci, ci, f32, u64, vp]
    lib.fa_b200_fwd_dropout.restype = ci
    lib.fa_b200_bwd_dropout.argtypes = [vp, vp, vp, vp, vp, f32p, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, f32, u64, vp, sz, vp]
    lib.fa_b200_bwd_dropout.restype = ci
    lib.fa_b200_dropout_mask.argtypes = [vp, ci, ci, ci, f32, u64, vp]
    lib.fa_b200_dropout_mask.restype = ci
    lib.fa_b200_dropout_keep_scale.argtypes = [f32]
    lib.fa_b200_dropout_keep_scale.restype = f32
    lib.fa_b200_dropout_effective_p.argtypes = [f32]
    lib.fa_b200_dropout_effective_p.restype = f32
    lib.fa_b200_gemm_tf32.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
    lib.fa_b200_gemm_tf32.restype = ci
    if hasattr(lib, "fa_b200_gemm_tf32_heads"):  # (absent only in older builds loaded through FA_B200_LIB for A/B runs)
        lib.fa_b200_gemm_tf32_heads.argtypes = [vp, vp, vp, vp] + [ci] * 12 + [vp]
        lib.fa_b200_gemm_tf32_heads.restype = ci
        lib.fa_b200_colsum_heads_f32.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp]
        lib.fa_b200_colsum_heads_f32.restype = ci
    if hasattr(lib, "fa_b200_colsum_heads_parts_f32"):
        lib.fa_b200_colsum_heads_parts_f32.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp]
        lib.fa_b200_colsum_heads_parts_f32.restype = ci
    lib.fa_b200_colsum_f32.argtypes = [vp, vp, ci, ci, ci, vp]
    lib.fa_b200_colsum_f32.restype = ci
    lib.fa_b200_add_layernorm_fwd.argtypes = [vp] * 8 + [ci, ci, f32, vp]
    lib.fa_b200_add_layernorm_fwd.restype = ci
    lib.fa_b200_layernorm_bwd_res.argtypes = [vp] * 10 + [ci, ci, ci, vp]
    for name in ("fa_b200_gelu_dropout_fwd", "fa_b200_gelu_dropout_bwd"):
        fn = getattr(lib, name)
        fn.argtypes = ([vp, vp] if name.endswith("fwd") else [vp, vp, vp]) + [sz, f32, u64, vp, vp]
        fn.restype = ci
    lib.fa_b200_layernorm_bwd_res.restype = ci
    lib.fa_b200_softmax_loss_bwd_colsum.argtypes = [vp] * 6 + [ci, ci, ci, vp]
    lib.fa_b200_softmax_loss_bwd_colsum.restype = ci
    lib.fa_b200_softmax_loss_colsum_width.argtypes = [ci]
    lib.fa_b200_softmax_loss_colsum_width.restype = ci
    lib.fa_b200_embed_pos_dropout_fwd.argtypes = [vp, vp, vp, vp, sz, ci, ci, f32, u64, vp, vp]
    lib.fa_b200_embed_pos_dropout_fwd.restype = ci
    lib.fa_b200_embed_pos_dropout_bwd.argtypes = [vp, vp, vp, vp, sz, ci, ci, sz, sz, f32, u64, vp, ci, vp]
    lib.fa_b200_embed_pos_dropout_bwd.restype = ci
    lib.fa_b200_gelu_dropout_bwd_colsum.argtypes = [vp, vp, vp, vp, ci, ci, f32, u64, vp, ci, vp]
    lib.fa_b200_gelu_dropout_bwd_colsum.restype = ci
    lib.fa_b200_layernorm_fused_width.argtypes = [ci]
    lib.fa_b200_layernorm_fused_width.restype = ci
    for name, args in (("layernorm_fwd", [vp] * 6 + [ci, ci, f32, vp]), ("layernorm_bwd", [vp] * 8 + [ci, ci, vp]), ("layernorm_bwd_accum", [vp] * 8 + [ci, ci, vp]),
                       ("gelu_fwd", [vp, vp, sz, vp]), ("gelu_bwd", [vp, vp, vp, sz, vp]),
                       ("dropout_apply", [vp, vp, sz, f32, u64, vp, vp]), ("embedding_fwd", [vp, vp, vp, sz, ci, vp]),
                       ("embedding_bwd", [vp, vp, vp, sz, ci, sz, vp]), ("embedding_bwd_accum", [vp, vp, vp, sz, ci, sz, vp]), ("softmax_loss_fwd", [vp] * 4 + [ci, ci, vp]),
                       ("softmax_loss_bwd", [vp] * 5 + [ci, ci, vp]),
                       ("adam_step", [vp] * 4 + [sz, f32, f32, f32, f32, f32, vp])):
        fn = getattr(lib, "fa_b200_" + name)
        fn.argtypes, fn.restype = args, ci
    lib.fa_b200_host_fwd.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
    lib.fa_b200_host_fwd.restype = ci
    lib.fa_b200_host_bwd.argtypes = [vp] * 10 + [ci, ci, ci, ci, ci, vp]
    lib.fa_b200_host_bwd.restype = ci
    lib.fa_b200_set_host_precision.argtypes = [ci]
    lib.fa_b200_set_host_precision.restype = None
    lib.fa_b200_get_host_precision.restype = ci
    if hasattr(lib, "fa_b200_set_fwd_pair"):
        lib.fa_b200_set_fwd_pair.argtypes = [ci]
        lib.fa_b200_set_fwd_pair.restype = None
        lib.fa_b200_get_fwd_pair.restype = ci
    if hasattr(lib, "fa_b200_set_fwd_persistent"):
        lib.fa_b200_set_fwd_persistent.argtypes = [ci]
        lib.fa_b200_set_fwd_persistent.restype = None
        lib.fa_b200_get_fwd_persistent.argtypes = []
        lib.fa_b200_get_fwd_persistent.restype = ci
    if hasattr(lib, "fa_b200_set_bwd_persistent"):
        lib.fa_b200_set_bwd_persistent.argtypes = [ci]
        lib.fa_b200_set_bwd_persistent.restype = None
        lib.fa_b200_get_bwd_persistent.argtypes = []
        lib.fa_b200_get_bwd_persistent.restype = ci
    if hasattr(lib, "fa_b200_set_host_zero_qgrad"):
        lib.fa_b200_set_host_zero_qgrad.argtypes = [ci]
        lib.fa_b200_set_host_zero_qgrad.restype = None
    lib.fa_b200_release_arena.restype = None
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != FA_OK:
        raise FaError(rc, load_library().fa_b200_last_error().decode())


def launch_count() -> int:
    return int(load_library().fa_b200_launch_count())
