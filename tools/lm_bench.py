#!/usr/bin/env python
"""Per-kernel roofline numbers of the decoder-LM slice at the shapes of the reference's MT run (tokens = 11 x 512,
n_embd 256, vocab 10 000): CUDA-event time of 50 launches replayed from a CUDA graph, algorithmic bytes / FLOPs, fraction of the
measured peak (MEASURED_PEAKS.json: HBM copy GB/s; tf32 tensor peak taken as half the measured bf16 rate).
    python tools/lm_bench.py > gpurun_out/lm_bench.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-minitorch_b200"))
from fa_b200 import _lib  # noqa: E402

lib = _lib.load_library()
dev = torch.device("cuda:0")
peaks = {"hbm_gbs": 6547.2, "bf16_tflops": 1638.1}
try:
    peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except Exception:
    pass
HBM, TF32 = peaks["hbm_gbs"], peaks["bf16_tflops"] / 2
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
P = lambda t: t.data_ptr()  # noqa: E731


def timeit(fn, iters=50):
    """Device time per launch: `iters` launches captured into one CUDA graph and replayed (a Python launch loop would
    measure the ~5 us host cost per ctypes call, not the kernels)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side, capture_error_mode="relaxed"):
        for _ in range(iters):
            fn()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


rows = []


def report(name, seconds, nbytes=None, flops=None):
    r = {"kernel": name, "us": seconds * 1e6}
    if nbytes is not None:
        r.update(bound="hbm", algorithmic_MB=nbytes / 1e6, achieved_GBs=nbytes / seconds / 1e9, frac=nbytes / seconds / 1e9 / HBM)
    if flops is not None:
        r.update(bound="tensor(tf32)", GFLOP=flops / 1e9, achieved_TFLOPs=flops / seconds / 1e12, frac=flops / seconds / 1e12 / TF32)
    rows.append(r)
    print(json.dumps(r), flush=True)


M, E, V, MID = 11 * 512, 256, 10000, 256
f32 = dict(device=dev, dtype=torch.float32)
x, w, bias = torch.randn(M, E, **f32), torch.randn(E, E, **f32), torch.randn(E, **f32)
wv, bv = torch.randn(E, V, **f32), torch.randn(V, **f32)
dy, dyv = torch.randn(M, E, **f32), torch.randn(M, V, **f32)
y, yv = torch.empty(M, E, **f32), torch.empty(M, V, **f32)
dw, dwv = torch.empty(E, E, **f32), torch.empty(E, V, **f32)


def gemm(a, b, c, bi, m, n, k, amn, bmn):
    return lambda: _lib.check(lib.fa_b200_gemm_tf32(P(a), P(b), P(c), P(bi) if bi is not None else None, m, n, k, amn, bmn, 0, st()))


report("gemm Y=XW+b   5632x256x256", timeit(gemm(x, w, y, bias, M, E, E, 0, 1)), flops=2.0 * M * E * E)
report("gemm logits   5632x10000x256", timeit(gemm(x, wv, yv, bv, M, V, E, 0, 1)), flops=2.0 * M * V * E, )
report("gemm dX=dY W^T 5632x256x10000", timeit(gemm(dyv, wv, y, None, M, E, V, 0, 0)), flops=2.0 * M * V * E)
report("gemm dW=X^T dY 256x10000x5632", timeit(gemm(x, dyv, dwv, None, E, V, M, 1, 1)), flops=2.0 * M * V * E)
report("gemm dW=X^T dY 256x256x5632 (split-K)", timeit(gemm(x, dy, dw, None, E, E, M, 1, 1)), flops=2.0 * M * E * E)
db = torch.empty(V, **f32)
report("colsum 5632x10000", timeit(lambda: _lib.check(lib.fa_b200_colsum_f32(P(dyv), P(db), M, V, 0, st()))), nbytes=4.0 * M * V)

lw, lb = torch.rand(E, **f32) + 0.5, torch.randn(E, **f32)
mean, rstd = torch.empty(M, **f32), torch.empty(M, **f32)
dx, dlw, dlb = torch.empty(M, E, **f32), torch.empty(E, **f32), torch.empty(E, **f32)
report("layernorm fwd 5632x256", timeit(lambda: _lib.check(lib.fa_b200_layernorm_fwd(P(x), P(lw), P(lb), P(y), P(mean), P(rstd), M, E, 1e-5, st()))),
       nbytes=2.0 * 4 * M * E)
report("layernorm bwd 5632x256", timeit(lambda: _lib.check(lib.fa_b200_layernorm_bwd(P(x), P(lw), P(dy), P(mean), P(rstd), P(dx), P(dlw), P(dlb), M, E, st()))),
       nbytes=5.0 * 4 * M * E)
x2, ssum, dcol = torch.randn(M, E, **f32), torch.empty(M, E, **f32), torch.zeros(E, **f32)
report("residual add + layernorm fwd 5632x256", timeit(lambda: _lib.check(lib.fa_b200_add_layernorm_fwd(P(x), P(x2), P(lw), P(lb), P(ssum), P(y), P(mean), P(rstd), M, E, 1e-5, st()))),
       nbytes=4.0 * 4 * M * E)
report("layernorm bwd + branch grad + dw/db + producer bias grad (accumulating) 5632x256",
       timeit(lambda: _lib.check(lib.fa_b200_layernorm_bwd_res(P(x), P(lw), P(dy), P(mean), P(rstd), P(x2), P(dx), P(dlw), P(dlb), P(dcol), M, E, 1, st()))),
       nbytes=4.0 * 4 * M * E)
report("gelu+dropout fwd 5632x256", timeit(lambda: _lib.check(lib.fa_b200_gelu_dropout_fwd(P(x), P(y), M * MID, 0.1, 7, None, st()))), nbytes=2.0 * 4 * M * MID)
report("gelu+dropout bwd 5632x256", timeit(lambda: _lib.check(lib.fa_b200_gelu_dropout_bwd(P(x), P(dy), P(dx), M * MID, 0.1, 7, None, st()))), nbytes=3.0 * 4 * M * MID)
report("gelu fwd 5632x256", timeit(lambda: _lib.check(lib.fa_b200_gelu_fwd(P(x), P(y), M * MID, st()))), nbytes=2.0 * 4 * M * MID)
report("gelu bwd 5632x256", timeit(lambda: _lib.check(lib.fa_b200_gelu_bwd(P(x), P(dy), P(dx), M * MID, st()))), nbytes=3.0 * 4 * M * MID)
report("dropout 5632x256", timeit(lambda: _lib.check(lib.fa_b200_dropout_apply(P(x), P(y), M * MID, 0.1, 7, None, st()))), nbytes=2.0 * 4 * M * MID)
target = torch.randint(0, V, (M,), device=dev)
loss, lse, gs = torch.empty(M, **f32), torch.empty(M, **f32), torch.full((M,), 1.0 / M, **f32)
report("softmax_loss fwd 5632x10000", timeit(lambda: _lib.check(lib.fa_b200_softmax_loss_fwd(P(yv), P(target), P(loss), P(lse), M, V, st()))),
       nbytes=4.0 * M * V)
report("softmax_loss bwd 5632x10000", timeit(lambda: _lib.check(lib.fa_b200_softmax_loss_bwd(P(yv), P(target), P(lse), P(gs), P(dyv), M, V, st()))),
       nbytes=2.0 * 4 * M * V)
dbv = torch.zeros(V, **f32)
report("softmax_loss bwd + lm_head bias grad 5632x10000", timeit(lambda: _lib.check(lib.fa_b200_softmax_loss_bwd_colsum(P(yv), P(target), P(lse), P(gs), P(dyv), P(dbv), M, V, 1, st()))),
       nbytes=2.0 * 4 * M * V)
qkv = torch.randn(3, 11, 8, 512, 32, **f32)
dbq = torch.zeros(3 * E, **f32)
report("bias grad of the fused q|k|v projection (3,11,8,512,32)", timeit(lambda: _lib.check(lib.fa_b200_colsum_heads_parts_f32(P(qkv), P(dbq), 3, 11, 8, 512, 32, 1, st()))),
       nbytes=4.0 * 3 * M * E)
n = 9273616
p_, g_, m_, v_ = (torch.randn(n, **f32) for _ in range(4))
v_.abs_()
report("adam 9.27M parameters", timeit(lambda: _lib.check(lib.fa_b200_adam_step(P(p_), P(g_), P(m_), P(v_), n, 1e-3, 0.9, 0.999, 0.1, 1e-8, st()))),
       nbytes=7.0 * 4 * n)
idx = torch.randint(0, V, (11, 512), device=dev)
emb, ey = torch.randn(V, E, **f32), torch.empty(11, 512, E, **f32)
report("embedding fwd 5632 rows", timeit(lambda: _lib.check(lib.fa_b200_embedding_fwd(P(idx), P(emb), P(ey), M, E, st()))), nbytes=2.0 * 4 * M * E)
demb = torch.empty(V, E, **f32)
report("embedding bwd (zero 10000x256 + scatter)", timeit(lambda: _lib.check(lib.fa_b200_embedding_bwd(P(idx), P(ey), P(demb), M, E, V, st()))),
       nbytes=4.0 * (V * E + 2 * M * E))
with open(os.path.join(ROOT, "gpurun_out", "lm_bench.md"), "w") as f:
    f.write("| kernel | time (us) | bound | achieved | fraction of measured peak |\n|---|---|---|---|---|\n")
    for r in rows:
        ach = f"{r['achieved_GBs']:.0f} GB/s" if r["bound"] == "hbm" else f"{r['achieved_TFLOPs']:.1f} TFLOP/s"
        f.write(f"| {r['kernel']} | {r['us']:.1f} | {r['bound']} | {ach} | {r['frac']:.2f} |\n")
