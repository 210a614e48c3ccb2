#!/usr/bin/env python
"""Why is the forward 2.4 ms inside bench.py and 2.0 ms in tools/fa_check?  Times 10 forwards (CUDA events) cold, right after a
fw+bw loop, and after a pause, sampling the SM clock with NVML around each."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "flash-attention-minitorch_b200"))
from fa_b200 import device  # noqa: E402
import pynvml  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
clk = lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)  # noqa: E731
pw = lambda: pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0  # noqa: E731
dev = torch.device("cuda:0")
BH, T, d = 512, 4096, 128
mk = lambda: (torch.rand(BH, T, d, device=dev) * 2 - 1).to(torch.bfloat16)  # noqa: E731
q, k, v, g = mk(), mk(), mk(), mk()
outs = (torch.empty_like(q), torch.empty(BH, T, dtype=torch.float32, device=dev))
grads = (torch.empty_like(q), torch.empty_like(q), torch.empty_like(q))
ws = device.bwd_workspace(q)


def timed(fn, reps, label):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    c_mid, p_mid = clk(), pw()
    b.record()
    torch.cuda.synchronize()
    print(f"{label:34s} {a.elapsed_time(b) / reps:7.3f} ms per call   SM clock while queued {c_mid} MHz, {p_mid:.0f} W, after {clk()} MHz")


fwd = lambda: device.flash_attn_fwd(q, k, v, True, out=outs)  # noqa: E731
o, lse = fwd()
bwd = lambda: device.flash_attn_bwd(q, k, v, o, g, lse, True, workspace=ws, out=grads)  # noqa: E731
timed(fwd, 10, "forward x10, cold")
timed(fwd, 50, "forward x50")
timed(bwd, 10, "backward x10")
timed(lambda: (fwd(), bwd()), 20, "fw+bw x20")
timed(fwd, 10, "forward x10 right after fw+bw")
time.sleep(3)
timed(fwd, 10, "forward x10 after 3 s pause")
q2 = (torch.rand(BH, T, d, device=dev)).to(torch.bfloat16)  # different score statistics (all positive)
timed(lambda: device.flash_attn_fwd(q2, k, v, True, out=outs), 10, "forward x10, q in (0,1)")
