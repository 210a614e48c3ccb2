"""Host-side logic that needs no GPU: the minitorch-compatible surface, sharding, and the N>1 path
on gloo (world_size 2)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from fa_b200 import minitorch_compat as mt
from fa_b200 import sharding


def test_tensor_bool_is_shape_unit_tensor():
    t = mt.tensor(True)
    assert t.shape == () and int(t._tensor._storage.item()) == 1
    assert int(mt.tensor(False)._tensor._storage.item()) == 0


def test_tensor_from_numpy_copies_and_views():
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    t = mt.tensor_from_numpy(a)
    a[0, 0, 0] = 99
    assert t.to_numpy()[0, 0, 0] == 0
    p = t.permute(0, 2, 1)
    assert p.shape == (2, 4, 3) and not p._tensor.is_contiguous()
    c = p.contiguous()
    assert c._tensor.is_contiguous() and np.array_equal(c.to_numpy(), np.arange(24, dtype=np.float32).reshape(2, 3, 4).transpose(0, 2, 1))
    assert c.view(8, 3).shape == (8, 3)
    with pytest.raises(AssertionError):
        p.view(24)


class RecordingOps:
    """Stands in for CudaKernelOps: checks what the Functions hand to the backend."""
    calls = []

    @staticmethod
    def _fw(name):
        def f(q, k, v, causal_mask):
            RecordingOps.calls.append((name, q.shape, int(causal_mask._tensor._storage.item())))
            return q.zeros(q.shape), q.zeros(q.shape[:-1]), q.zeros(q.shape[:-1])
        return staticmethod(f)

    @staticmethod
    def _bw(name):
        def f(q, k, v, out, out_grad, l, m, causal_mask):
            RecordingOps.calls.append((name, out_grad.shape, float(out_grad.to_numpy().mean())))
            return q.zeros(q.shape), k.zeros(k.shape), v.zeros(v.shape), causal_mask
        return staticmethod(f)


for _n in ("flash_attn", "flash_attn2", "flash_attn_causal"):
    setattr(RecordingOps, _n + "_fw", RecordingOps._fw(_n + "_fw"))
    setattr(RecordingOps, _n + "_bw", RecordingOps._bw(_n + "_bw"))


@pytest.mark.parametrize("method,prefix", [("flash_attn", "flash_attn"), ("flash_attn2", "flash_attn2"),
                                           ("flash_attn_causal", "flash_attn_causal")])
def test_functions_dispatch_like_the_reference(method, prefix):
    RecordingOps.calls.clear()
    be = mt.TensorBackend(RecordingOps)
    mk = lambda: mt.tensor_from_numpy(np.ones((1, 2, 4, 8), np.float32), backend=be, requires_grad=True)  # noqa: E731
    q, k, v = mk(), mk(), mk()
    out = getattr(q, method)(k, v, causal_mask=True)
    assert RecordingOps.calls == [(prefix + "_fw", (1, 2, 4, 8), 1)]
    out.sum().backward()  # reference style: out_grad = ones (kernel_tests/test_flashattn_bw.py:32)
    assert RecordingOps.calls[-1] == (prefix + "_bw", (1, 2, 4, 8), 1.0)
    assert q.grad is not None and k.grad is not None and v.grad is not None and q.grad.shape == (1, 2, 4, 8)


def test_backend_without_flash_ops_raises_like_reference():
    class NoFlash:
        pass
    with pytest.raises(AttributeError):
        mt.TensorBackend(NoFlash)


def test_generic_marshaller_asserts_shapes():
    be = mt.TensorBackend(mt.CudaKernelOps)
    q = mt.tensor_from_numpy(np.zeros((1, 1, 4, 8), np.float32), backend=be)
    k = mt.tensor_from_numpy(np.zeros((1, 1, 5, 8), np.float32), backend=be)
    with pytest.raises(AssertionError):
        mt.CudaKernelOps.flash_attn_fw_generic(q, k, q, mt.tensor(False), None)


@pytest.mark.parametrize("bh,world", [(512, 8), (512, 1), (10, 4), (3, 8), (128, 2)])
def test_shard_bounds_partition(bh, world):
    spans = [sharding.shard_bounds(bh, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == bh
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = sharding.shard_sizes(bh, world)
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == bh


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_sharded_attention_matches_single_rank():
    """bench.py's N>1 path with gloo on CPU: each rank owns a contiguous slice of B*H, computes it (here with
    the oracle standing in for the kernel), and the only communication is the max-over-ranks of a timing
    scalar and a gather used by this test to compare with the unsharded result."""
    worker = os.path.join(ROOT, "tests", "_gloo_worker.py")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, worker], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "sharded==full True" in outs[0]


# ---- thin NCCL C ABI (fa_b200_nccl_*) and its host side: what can be checked without a GPU ----
def test_unique_id_rendezvous_over_tcp():
    """Rank 0 serves the 128-byte id, the other ranks fetch it (fa_b200/mt/comm.py) -- three ranks as threads."""
    import socket
    import threading
    from fa_b200.mt import comm
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    payload = bytes(range(128))
    got = {}

    def run(rank):
        got[rank] = comm.exchange_unique_id(rank, 3, payload if rank == 0 else None, "127.0.0.1", port, timeout_s=30)

    threads = [threading.Thread(target=run, args=(r,)) for r in (1, 2, 0)]  # clients first: they retry until rank 0 listens
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert got == {0: payload, 1: payload, 2: payload}


def test_nccl_abi_rejects_use_before_init():
    import torch  # noqa: F401  (first: the process must not end up with two different libnccl.so.2)
    from fa_b200.mt.comm import NcclComm
    c = NcclComm()
    lib = c._bind()
    assert lib.fa_b200_nccl_world() == 0 and lib.fa_b200_nccl_rank() == -1
    assert lib.fa_b200_nccl_allreduce_f32(None, 4, 1, None) != 0
    assert b"not initialised" in lib.fa_b200_nccl_last_error()
    assert lib.fa_b200_nccl_init(None, 0, 1) != 0
    assert lib.fa_b200_nccl_destroy() == 0  # nothing to destroy: not an error


def test_flat_layout_is_clustered_per_top_level_module_and_aligned():
    """Module.flatten_parameters: every top-level child (embeddings, layers, final LayerNorm, lm_head) is one contiguous,
    128-byte aligned range of the flat buffers -- what train.gradient_buckets cuts the gradient all-reduce along."""
    import torch
    from fa_b200.mt.model import DecoderLM
    from fa_b200.mt.train import gradient_buckets
    m = DecoderLM(n_vocab=50, n_embd=16, n_head=2, n_positions=8, p_dropout=0.0, device="cpu", dtype=torch.float64, seed=3,
                  attention_fn=lambda q, k, v, causal: v)
    before = {n: p.value.detach().clone() for n, p in m.named_parameters()}
    flat, gflat = m.flatten_parameters(fuse=True)
    r = m.flat_ranges()
    assert list(r) == ["token_embeddings", "position_embeddings", "t_layer_1", "t_layer_2", "t_layer_3", "t_layer_4", "ln", "lm_head"]
    ends = 0
    for name, (a, b) in r.items():
        assert a == ends and a % 32 == 0 and b > a
        ends = b
    assert ends == flat.numel()
    esz = flat.element_size()
    for n, p in m.named_parameters():
        assert torch.equal(p.value.detach(), before[n])  # values survive the move
        off = (p.value.data_ptr() - flat.data_ptr()) // esz
        a, b = r[n.split(".")[0]]
        assert a <= off < b and (p.value.grad.data_ptr() - gflat.data_ptr()) // esz == off
    buckets = gradient_buckets(m, gflat)
    assert buckets is not None and sum(b - a for a, b in buckets) == gflat.numel() and len(buckets) == 6
    assert buckets[0] == (r["ln"][0], r["lm_head"][1]) and buckets[4] == r["t_layer_1"] and buckets[5] == (0, r["position_embeddings"][1])
    # position rows >= seq_len never get a gradient: the embedding bucket stops after row seq_len of that table
    trimmed = gradient_buckets(m, gflat, seq_len=8)
    assert trimmed[:5] == buckets[:5] and trimmed[5] == (0, r["position_embeddings"][0] + 8 * 16)
