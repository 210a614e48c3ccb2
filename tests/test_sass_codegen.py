"""Codegen guards on the shipped library (CPU only: cuobjdump on lib/libfa_b200.so, nothing is launched).

The attention kernels are Blackwell-native only if ptxas emitted what the source means: tcgen05.mma as UTCHMMA, TMEM
traffic as LDTM / STTM, TMA as UTMALDG / UTMAREDG, and no legacy HMMA.  The second half guards a regression that cost the
backward 9 % in round 2: with one MMA operand that ptxas cannot prove warp-uniform it builds ALL operands
of the issue loops in vector registers -- ELECT + four R2UR.BROADCAST per UTCHMMA and BRA.DIV divergence checks.
"""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "flash-attention-minitorch_b200", "lib", "libfa_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump (CUDA toolkit) not available")


@pytest.fixture(scope="module")
def kernels():
    """demangled kernel name -> {opcode: count}"""
    assert os.path.exists(LIB), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            out[name] = {}
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            for key in {m.group(1), m.group(1).split(".")[0]}:
                out[name][key] = out[name].get(key, 0) + 1
    return out


def find(kernels, *parts):
    hits = [k for k in kernels if all(p in k for p in parts)]
    assert hits, parts
    return hits


def test_attention_kernels_are_tcgen05_tmem_tma(kernels):
    for k in find(kernels, "fa_fwd_kernel<128, __nv_bfloat16") + find(kernels, "fa_bwd_kernel<128, __nv_bfloat16"):
        c = kernels[k]
        assert c.get("UTCHMMA", 0) > 0 and c.get("LDTM", 0) > 0 and c.get("STTM", 0) > 0 and c.get("UTMALDG", 0) > 0, k
        assert c.get("HMMA", 0) == 0 and c.get("HGMMA", 0) == 0, k
    for k in find(kernels, "fa_bwd_kernel<128, __nv_bfloat16"):
        assert kernels[k].get("UTMAREDG", 0) > 0, k  # dQ: cp.reduce.async.bulk.tensor


def test_backward_issue_loops_run_on_the_uniform_datapath(kernels):
    # regressed state: R2UR 290 / UTCHMMA 42 = 6.9 per MMA, BRA.DIV 28; fixed: R2UR 132 / UTCHMMA 48 = 2.75 (what is left are
    # the TMEM addresses of the compute and drain warps), BRA.DIV 0.  Every template instance of the backward goes through the same issuer code.
    for k in find(kernels, "fa_bwd_kernel<"):
        c = kernels[k]
        assert c.get("BRA.DIV", 0) == 0, (k, c.get("BRA.DIV"))
        assert c.get("R2UR", 0) <= 3.5 * c["UTCHMMA"], (k, c.get("R2UR"), c["UTCHMMA"])


def test_backward_timeline_hooks_are_compiled_out(kernels):
    # clock64() of a timeline hook shows up as CS2R ..., SR_CLOCKLO; the watchdog of the mbarrier waits reads SR_GLOBALTIMERLO
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    body, keep = [], False
    for line in sass.splitlines():
        if "Function :" in line:
            keep = "fa_bwd_kernel" in line and "tf32" not in line
        elif keep:
            body.append(line)
    assert not [l for l in body if "SR_CLOCKLO" in l]
