"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

    NUMBA_DISABLE_CUDA=1 NUMBA_CACHE_DIR=/tmp/nbcache PYTHONPATH=/root/reference:/root/reference/kernel_tests \
        python oracle/gen_golden.py

Two reference code paths are executed, unmodified, by import (nothing is copied):
  1. minitorch vanilla attention forward + backward on the numba CPU backend -- the numerics oracle the
     reference's own kernel tests compare against (kernel_tests/test_flashattn_fw.py:60-76,
     test_flashattn_bw.py:152-174): softmax((q @ kT)/sqrt(d) [+ M], dim) @ v ; out.sum().backward() or a
     weighted sum for a non-trivial out_grad.  Inputs are flattened to (B*H, T, d) because FastOps'
     matmul only indexes 3 dims (fast_ops.py:336-349).
  2. kernel_tests/flash_attn_python.py: the authors' torch restatements of FA-1 / FA-2 forward and
     backward (non-causal, one head), which pin the (l, m) and L conventions.
Each case stores inputs and reference outputs as float32/float64 arrays; files are a few hundred KB.
"""
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def fastops_backend():
    import minitorch
    from minitorch.fast_ops import FastOps

    class FastOpsShim(FastOps):
        pass

    for n in ["attn_softmax_fw", "attn_softmax_bw", "layernorm_fw", "layernorm_bw", "flash_attn_fw", "flash_attn_bw",
              "flash_attn2_fw", "flash_attn2_bw", "flash_attn_causal_fw", "flash_attn_causal_bw"]:
        setattr(FastOpsShim, n, staticmethod(lambda *a, **k: None))
    return minitorch, minitorch.TensorBackend(FastOpsShim)


def minitorch_case(minitorch, be, name, BH, T, d, causal, seed, ones_grad):
    rng = np.random.default_rng(seed)
    q = rng.uniform(-1, 1, (BH, T, d)).astype(np.float32)
    k = rng.uniform(-1, 1, (BH, T, d)).astype(np.float32)
    v = rng.uniform(-1, 1, (BH, T, d)).astype(np.float32)
    g = np.ones((BH, T, d), np.float32) if ones_grad else np.random.default_rng(seed + 1).uniform(-1, 1, (BH, T, d)).astype(np.float32)
    tq = minitorch.tensor_from_numpy(q.copy(), backend=be, requires_grad=True)
    tkT = minitorch.tensor_from_numpy(np.ascontiguousarray(k.transpose(0, 2, 1)), backend=be, requires_grad=True)
    tv = minitorch.tensor_from_numpy(v.copy(), backend=be, requires_grad=True)
    s = (tq @ tkT) / np.sqrt(d)
    if causal:
        datatype = np.float32
        M = -np.finfo(datatype).max * np.triu(np.ones((1, T, T), dtype=datatype), 1)  # modules_transfomer.py:63
        s = s + minitorch.tensor_from_numpy(M, backend=be)
    o = minitorch.nn.softmax(s, dim=2) @ tv
    tg = minitorch.tensor_from_numpy(g.copy(), backend=be)
    (o * tg).sum().backward()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), q=q, k=k, v=v, d_out=g, causal=np.int32(causal),
                        o=o.to_numpy().astype(np.float32), dq=tq.grad.to_numpy().astype(np.float32),
                        dk=np.ascontiguousarray(tkT.grad.to_numpy().transpose(0, 2, 1)).astype(np.float32),
                        dv=tv.grad.to_numpy().astype(np.float32),
                        source=np.array("reference minitorch vanilla attention on FastOps (numba CPU), fp32"))
    print("wrote", name, o.shape)


def torch_flash_case(name, T, d, seed):
    import torch
    import flash_attn_python as fap  # /root/reference/kernel_tests

    g = torch.Generator().manual_seed(seed)
    Q = (torch.rand(T, d, generator=g, dtype=torch.float64) * 2 - 1)
    K = (torch.rand(T, d, generator=g, dtype=torch.float64) * 2 - 1)
    V = (torch.rand(T, d, generator=g, dtype=torch.float64) * 2 - 1)
    dO = (torch.rand(T, d, generator=g, dtype=torch.float64) * 2 - 1)
    O_ref = fap.compute_attention(Q, K, V)
    O1, l1, m1 = fap.flash_attention(Q, K, V)
    O2, L2 = fap.flash_attention2(Q.float(), K.float(), V.float())  # its l_i/m_i temporaries are fp32
    dq1, dk1, dv1 = fap.flash_attention_backward(Q, K, V, O1, dO, l1, m1)
    dq2, dk2, dv2 = fap.flash_attention2_backward(Q, K, V, O1, dO, (m1 + torch.log(l1)))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), q=Q.numpy()[None], k=K.numpy()[None], v=V.numpy()[None],
                        d_out=dO.numpy()[None], causal=np.int32(0), o=O_ref.numpy()[None], o_fa1=O1.numpy()[None],
                        l_fa1=l1.numpy()[None], m_fa1=m1.numpy()[None], o_fa2=O2.numpy()[None].astype(np.float64),
                        L_fa2=L2.numpy()[None].astype(np.float64), dq_fa1=dq1.numpy()[None], dk_fa1=dk1.numpy()[None],
                        dv_fa1=dv1.numpy()[None], dq_fa2=dq2.numpy()[None], dk_fa2=dk2.numpy()[None], dv_fa2=dv2.numpy()[None],
                        source=np.array("reference kernel_tests/flash_attn_python.py (torch fp64; flash_attention2 in fp32)"))
    print("wrote", name)


def main():
    os.makedirs(OUT, exist_ok=True)
    minitorch, be = fastops_backend()
    # BASELINE.json configs[0] tile shape (T=128, d=64; 4 of its 16 batch*heads to keep fixtures small), reference-style grads (ones) and random grads
    minitorch_case(minitorch, be, "mt_cfg1_noncausal_ones", 4, 128, 64, False, 1001, True)
    minitorch_case(minitorch, be, "mt_cfg1_causal_ones", 4, 128, 64, True, 1001, True)
    minitorch_case(minitorch, be, "mt_cfg1_causal_rand", 4, 128, 64, True, 1001, False)
    # kernel-test shapes: test_flashattn_comb.py (T=40, d=32), ragged T / odd d (test_flashattn_2_fw.py:137-138)
    minitorch_case(minitorch, be, "mt_comb_T40_d32_noncausal", 8, 40, 32, False, 1101, False)
    minitorch_case(minitorch, be, "mt_ragged_T77_d34_causal", 3, 77, 34, True, 1201, False)
    minitorch_case(minitorch, be, "mt_T200_d128_causal", 2, 200, 128, True, 1301, False)
    minitorch_case(minitorch, be, "mt_T1_d16_causal", 2, 1, 16, True, 1401, False)
    torch_flash_case("fap_T64_d32", 64, 32, 2101)
    torch_flash_case("fap_T100_d64", 100, 64, 2201)


if __name__ == "__main__":
    sys.exit(main())
