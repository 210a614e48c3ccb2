"""Loader contract of the UNMODIFIED reference Python (minitorch/cuda_kernel_ops.py:21-35) against the six shims.

INTEGRATION.md Option A, executed: a scratch working directory gets `minitorch/cuda_kernels/` with our six
reference-named flash libraries (cp, as a user would) plus the reference's own three non-flash libraries, `lib/` gets
libfa_b200.so (the shims carry RUNPATH $ORIGIN/../../lib), a stub `pycuda` package stands in for the one import the
reference makes but never uses on this path, and a fresh interpreter imports the reference's `minitorch.cuda_kernel_ops`
from /root/reference with that directory as cwd (the loader's paths are cwd-relative).  Import + symbol binding only: no
CUDA call is made, so it runs in the build container; it is skipped where /root/reference does not exist (the GPU box)."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT
from fa_b200 import _lib

REF = "/root/reference"
REF_BUILT = os.path.join(ROOT, "oracle", "_ref")
AUX = ("combine", "softmax_kernel", "layernorm_kernel")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "minitorch")), reason="needs the reference checkout (build container only)")
def test_unmodified_reference_loader_binds_the_six_shims(tmp_path):
    missing = [n for n in AUX if not os.path.exists(os.path.join(REF_BUILT, n + ".so"))]
    if missing:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j3", "ref_aux"], check=True, capture_output=True)
    w = tmp_path / "checkout"
    kdir = w / "minitorch" / "cuda_kernels"
    kdir.mkdir(parents=True)
    (w / "lib").mkdir()
    for name in _lib.SHIM_NAMES:  # INTEGRATION.md Option A, step 2
        shutil.copy(_lib.shim_path(name), kdir / (name + ".so"))
    shutil.copy(_lib.library_path(), w / "lib" / "libfa_b200.so")  # step 3
    for name in AUX:  # what the reference's own `make combine layernorm softmax` provides
        shutil.copy(os.path.join(REF_BUILT, name + ".so"), kdir / (name + ".so"))
    stub = tmp_path / "stub" / "pycuda"
    stub.mkdir(parents=True)
    for f in ("__init__.py", "autoinit.py", "driver.py"):
        (stub / f).write_text("")
    code = textwrap.dedent("""
        import ctypes, os, sys
        import minitorch
        import minitorch.cuda_kernel_ops as cko            # the reference's loader: nine cwd-relative CDLL calls
        fw = [cko.lib_flash_attn_fw, cko.lib_flash_attn2_fw, cko.lib_flash_attn_causal_fw]
        bw = [cko.lib_flash_attn_bw, cko.lib_flash_attn2_bw, cko.lib_flash_attn_causal_bw]
        assert all(hasattr(l, "launch_flash_attn_fw") for l in fw) and all(hasattr(l, "launch_flash_attn_bw") for l in bw)
        for l in fw + bw:                                  # the copies really are ours and found libfa_b200.so through their RUNPATH
            assert os.path.realpath(l._name).startswith(os.path.realpath(os.getcwd())), l._name
        ours = ctypes.CDLL(os.path.join("lib", "libfa_b200.so"))
        assert ours.fa_b200_version() >= 100
        ops = cko.CudaKernelOps
        for n in ("flash_attn_fw", "flash_attn_bw", "flash_attn2_fw", "flash_attn2_bw", "flash_attn_causal_fw", "flash_attn_causal_bw"):
            assert callable(getattr(ops, n))
        be = minitorch.TensorBackend(ops)                  # tensor_ops.py:102-108 wires the six entry points
        assert be.flash_attn2_fw is not None and be.flash_attn_causal_bw is not None
        print("LOADER-CONTRACT-OK", len(fw + bw))
    """)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REF, str(tmp_path / "stub")]), NUMBA_DISABLE_CUDA="1",
               NUMBA_CACHE_DIR=str(tmp_path / "nbcache"))
    env.pop("LD_LIBRARY_PATH", None)  # the RUNPATH of the copied shims must be enough
    r = subprocess.run([sys.executable, "-c", code], cwd=str(w), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "LOADER-CONTRACT-OK 6" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
