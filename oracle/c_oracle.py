"""ctypes wrapper of oracle/liboracle.so (the plain-C restatement).  TEST INFRASTRUCTURE ONLY -- see
oracle/attention_oracle.c for what it restates; only tests/, smoke() and bench.py's CPU legs use it."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "attention_oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    return path


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        _LIB.oracle_attention.argtypes = [f32p, f32p, f32p, ctypes.c_void_p, f32p, f32p, f32p, f32p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _LIB.oracle_attention.restype = ctypes.c_int
        _LIB.oracle_attention_dropout.argtypes = list(_LIB.oracle_attention.argtypes) + [ctypes.c_void_p, ctypes.c_float]
        _LIB.oracle_attention_dropout.restype = ctypes.c_int
        _LIB.oracle_num_threads.restype = ctypes.c_int
        _LIB.oracle_set_num_threads.argtypes = [ctypes.c_int]
        _LIB.oracle_set_num_threads.restype = None
    return _LIB


def use_all_cores() -> int:
    """Use every core this process may run on (torchrun pins OMP_NUM_THREADS=1).  Returns the thread count."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().oracle_set_num_threads(n)
    return num_threads()


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def attention(q, k, v, d_out=None, causal=False, keep=None, keep_scale=1.0):
    """fp32 vanilla attention on all host cores.  Returns dict(o, m, l, lse[, dq, dk, dv]).
    keep (uint8 [BH,T,T], 1 = kept) + keep_scale apply dropout Z = keep * keep_scale to P."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    BH, T, d = q.shape
    out = {"o": np.empty_like(q), "m": np.empty((BH, T), np.float32), "l": np.empty((BH, T), np.float32),
           "lse": np.empty((BH, T), np.float32)}
    gp = dqp = dkp = dvp = None
    if d_out is not None:
        g = np.ascontiguousarray(d_out, dtype=np.float32)
        out.update(dq=np.empty_like(q), dk=np.empty_like(q), dv=np.empty_like(q))
        gp, dqp, dkp, dvp = (a.ctypes.data_as(ctypes.c_void_p) for a in (g, out["dq"], out["dk"], out["dv"]))
    if keep is not None:
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        assert keep.shape == (BH, T, T)
        rc = lib().oracle_attention_dropout(q, k, v, gp, out["o"], out["m"], out["l"], out["lse"], dqp, dkp, dvp,
                                            BH, T, d, 1 if causal else 0, keep.ctypes.data_as(ctypes.c_void_p),
                                            float(keep_scale))
    else:
        rc = lib().oracle_attention(q, k, v, gp, out["o"], out["m"], out["l"], out["lse"], dqp, dkp, dvp,
                                    BH, T, d, 1 if causal else 0)
    if rc != 0:
        raise MemoryError("oracle_attention: allocation failed")
    return out
