"""tools/fuzz_persistent.py [cases] [seed]: random shapes through the persistent forward / backward against the same kernels
launched with one CTA per work item (bit-equality of O, LSE, dK, dV; dQ up to the order of its fp32 reduce-adds).  GPU only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-minitorch_b200")]
from fa_b200 import _lib, device  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
lib = _lib.load_library()
dev = torch.device("cuda:0")
bad = 0
for c in range(cases):
    d = int(rng.choice([64, 128]))
    T = int(rng.choice([rng.integers(1, 300), rng.integers(300, 3000), 128 * rng.integers(1, 24)]))
    BH = int(max(1, min(rng.integers(1, 400), 6_000_000 // (T * d))))
    causal = bool(rng.integers(0, 2))
    p_drop = float(rng.choice([0.0, 0.0, 0.1, 0.5]))
    g = torch.Generator(device=dev).manual_seed(c)
    q, k, v, do = (torch.randn(BH, T, d, device=dev, generator=g).to(torch.bfloat16) for _ in range(4))
    kw = dict(p_drop=p_drop, seed=7 + c) if p_drop else {}
    out = {}
    for mode in (0, 1, 1):
        lib.fa_b200_set_fwd_persistent(mode)
        lib.fa_b200_set_bwd_persistent(mode)
        o, lse = device.flash_attn_fwd(q, k, v, causal, **kw)
        dq, dk, dv = device.flash_attn_bwd(q, k, v, o, do, lse, causal, **kw)
        torch.cuda.synchronize()
        if mode not in out:
            out[mode] = (o, lse, dq, dk, dv)
            continue
        ref = out[0]
        ok = all(torch.equal(a, b) for a, b in zip((o, lse, dk, dv), (ref[0], ref[1], ref[3], ref[4])))
        scale = float(ref[2].float().abs().max()) + 1e-30
        ok = ok and float((dq.float() - ref[2].float()).abs().max()) <= 8e-3 * scale and bool(torch.isfinite(dq.float()).all())
        if not ok:
            bad += 1
            print(f"MISMATCH BH={BH} T={T} d={d} causal={causal} p_drop={p_drop}")
    lib.fa_b200_set_fwd_persistent(1)
    lib.fa_b200_set_bwd_persistent(1)
print(f"fuzz_persistent: {cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
