"""The C-ABI libraries load and export every symbol include/fa_b200.h declares (no compute: CPU only)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from fa_b200 import _lib

HEADER = os.path.join(ROOT, "include", "fa_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    ref_block = re.search(r"#ifdef FA_B200_DECLARE_REFERENCE_ABI(.*?)#endif", src, flags=re.S).group(1)
    main = src.replace(ref_block, "")
    pat = re.compile(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;")
    return sorted(set(pat.findall(main))), sorted(set(pat.findall(ref_block)))


def test_header_declares_expected_surface():
    main, ref = declared_functions()
    assert ref == ["launch_flash_attn_bw", "launch_flash_attn_fw"]
    for name in ("fa_b200_fwd", "fa_b200_bwd", "fa_b200_bwd_workspace_bytes", "fa_b200_host_fwd", "fa_b200_host_bwd"):
        assert name in main


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    main, _ = declared_functions()
    missing = [n for n in main if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.fa_b200_version() >= 100
    assert lib.fa_b200_launch_count() == 0 or lib.fa_b200_launch_count() > 0  # callable without a GPU


@pytest.mark.parametrize("name", _lib.SHIM_NAMES)
def test_reference_named_shims_export_reference_symbols(name):
    """Loader contract of minitorch/cuda_kernel_ops.py:30-35: six libraries, same symbol names."""
    path = _lib.shim_path(name)
    assert os.path.exists(path), f"{path} missing: run compile_cuda.sh"
    _lib.load_library()
    shim = ctypes.CDLL(path)
    sym = "launch_flash_attn_fw" if name.endswith("_fw") else "launch_flash_attn_bw"
    other = "launch_flash_attn_bw" if name.endswith("_fw") else "launch_flash_attn_fw"
    assert hasattr(shim, sym)
    assert not hasattr(shim, other)


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = _lib.load_library()
    assert lib.fa_b200_fwd(None, None, None, None, None, None, None, 0, 0, 1, 1, 1, 0, None) == _lib.FA_ERR_BAD_ARG
    assert b"null" in lib.fa_b200_last_error()
    one = ctypes.c_void_p(16)
    assert lib.fa_b200_fwd(one, one, one, one, None, None, None, 0, 0, 4, 0, 64, 0, None) == _lib.FA_ERR_BAD_ARG
    assert lib.fa_b200_fwd(one, one, one, one, None, None, None, 0, 0, 4, 16, 129, 0, None) == _lib.FA_ERR_BAD_ARG
    assert lib.fa_b200_fwd(one, one, one, one, None, None, None, 7, 0, 4, 16, 64, 0, None) == _lib.FA_ERR_BAD_ARG
    assert lib.fa_b200_bwd_workspace_bytes(_lib.FA_DTYPE_F32, _lib.FA_PREC_FP32_SIMT, 2, 128, 64) >= 2 * 128 * 4
    assert lib.fa_b200_bwd_workspace_bytes(_lib.FA_DTYPE_F32, _lib.FA_PREC_FP32_SIMT, 0, 128, 64) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_DIR", str(tmp_path))
    with pytest.raises(_lib.LibraryMissingError):
        _lib.load_library()
