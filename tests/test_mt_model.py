"""Decoder-LM slice (BASELINE.json configs[4]): the device-resident operator restatements against
fixtures produced by the reference's own modules, the Adam quirk, model wiring, and the data-parallel step on two
gloo ranks.  CPU only; the CUDA attention is replaced by a torch restatement injected by the tests."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from _mt_helpers import reference_forward, vanilla_attention
from fa_b200.mt import ops
from fa_b200.mt.model import DecoderLM, FeedForward, LayerNorm1d, Linear, loss_fn
from fa_b200.mt.optim import Adam
from fa_b200.mt.train import synthetic_batch, train

G = np.load(os.path.join(GOLDEN, "mod_reference_modules.npz"))
TOL = dict(rtol=2e-5, atol=2e-6)  # fp32 reference arithmetic vs fp64 restatement


def T64(a, grad=False):
    return torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=grad)


def close(a, b, **kw):
    tol = dict(TOL, **kw)
    return np.allclose(a.detach().numpy() if isinstance(a, torch.Tensor) else a, b, **tol)


def test_linear_matches_reference_module():
    x, w, b = T64(G["lin_x"], True), T64(G["lin_w"], True), T64(G["lin_b"], True)
    y = ops.linear(x, w, b)
    (y * T64(G["lin_gy"])).sum().backward()
    assert close(y, G["lin_y"]) and close(x.grad, G["lin_dx"]) and close(w.grad, G["lin_dw"]) and close(b.grad, G["lin_db"])


def test_layernorm_matches_reference_module():
    x, w, b = T64(G["ln_x"], True), T64(G["ln_w"], True), T64(G["ln_b"], True)
    y = ops.layernorm(x, w, b, float(G["ln_eps"]))
    (y * T64(G["ln_gy"])).sum().backward()
    assert close(y, G["ln_y"], atol=1e-5) and close(x.grad, G["ln_dx"], atol=1e-5)
    assert close(w.grad, G["ln_dw"], atol=1e-5) and close(b.grad, G["ln_db"], atol=1e-5)


def test_gelu_matches_reference():
    x = T64(G["gelu_x"], True)
    y = ops.gelu(x)
    (y * T64(G["gelu_gy"])).sum().backward()
    assert close(y, G["gelu_y"]) and close(x.grad, G["gelu_dx"])


def test_softmax_loss_matches_reference():
    logits = T64(G["sl_logits"], True)
    loss = ops.softmax_loss(logits, torch.tensor(G["sl_target"]))
    (loss * T64(G["sl_gw"])).sum().backward()
    assert close(loss, G["sl_loss"], atol=1e-5) and close(logits.grad, G["sl_dlogits"])


def test_embedding_matches_reference_one_hot_matmul():
    w = T64(G["emb_w"], True)
    y = ops.embedding(torch.tensor(G["emb_idx"]), w)
    (y * T64(G["emb_gy"])).sum().backward()
    assert close(y, G["emb_y"]) and close(w.grad, G["emb_dw"])


def test_feedforward_matches_reference_module():
    ff = FeedForward(16, 24, 0.0, True, "cpu", torch.float64, torch.Generator().manual_seed(0))
    for p, key in ((ff.linear_in.weights, "ff_w_in"), (ff.linear_in.bias, "ff_b_in"), (ff.linear_out.weights, "ff_w_out"),
                   (ff.linear_out.bias, "ff_b_out")):
        p.value = T64(G[key], True)
    x = T64(G["ff_x"], True)
    y = ff(x)
    (y * T64(G["ff_gy"])).sum().backward()
    assert close(y, G["ff_y"]) and close(x.grad, G["ff_dx"])
    assert close(ff.linear_in.weights.value.grad, G["ff_dw_in"]) and close(ff.linear_out.weights.value.grad, G["ff_dw_out"])


def test_adam_reproduces_reference_trajectory_including_its_beta1_quirk():
    from fa_b200.mt.model import Parameter
    p = Parameter(T64(G["adam_p0"]))
    opt = Adam([p], lr=float(G["adam_lr"]))
    for g, want in zip(G["adam_grads"], G["adam_traj"]):
        p.value.grad = T64(g)
        opt.step()
        assert close(p.value, want, rtol=1e-5, atol=1e-6)
    # a textbook Adam (second moment with 1 - beta2) would NOT follow the same trajectory
    q = torch.nn.Parameter(T64(G["adam_p0"]))
    ref = torch.optim.Adam([q], lr=float(G["adam_lr"]))
    for g in G["adam_grads"]:
        q.grad = T64(g)
        ref.step()
    assert not np.allclose(q.detach().numpy(), G["adam_traj"][-1], rtol=1e-3, atol=1e-4)


def small_model(p_dropout=0.0):
    return DecoderLM(n_vocab=40, n_embd=16, n_head=2, n_positions=8, p_dropout=p_dropout, device="cpu",
                     dtype=torch.float64, seed=3, attention_fn=vanilla_attention)


def test_model_tree_names_shapes_and_init_follow_the_reference():
    m = small_model()
    names = dict(m.named_parameters())
    # modules_transfomer.py:283-303: two embeddings (both sized n_vocab), four layers, final LN, lm_head
    assert names["token_embeddings.weights"].value.shape == (40, 16)
    assert names["position_embeddings.weights"].value.shape == (40, 16)
    for i in range(1, 5):
        for proj in ("q_projection", "k_projection", "v_projection", "out_projection"):
            assert names[f"t_layer_{i}.attention.{proj}.weights"].value.shape == (16, 16)
        assert names[f"t_layer_{i}.ff.linear_in.weights"].value.shape == (16, 256)  # FeedForward default middle_dim
        assert names[f"t_layer_{i}.ff.linear_out.weights"].value.shape == (256, 16)
        assert names[f"t_layer_{i}.ln_1.weights"].value.shape == (16,)
    assert names["lm_head.weights"].value.shape == (16, 40) and names["lm_head.bias"].value.shape == (40,)
    w = names["t_layer_1.ff.linear_out.weights"].value
    assert float(w.abs().max()) <= 1 / np.sqrt(256) + 1e-12  # RParam: U(-1/sqrt(in), 1/sqrt(in))
    assert len(m.parameters()) == 2 + 4 * (4 + 8 + 4) + 2 + 2


def test_model_is_causal_and_loss_matches_manual_cross_entropy():
    m = small_model()
    m.eval()
    gen = torch.Generator().manual_seed(1)
    batch = synthetic_batch(2, 8, 40, "cpu", gen)
    logits = m(batch["input_ids"])
    ids2 = batch["input_ids"].clone()
    ids2[:, 5:] = (ids2[:, 5:] + 7) % 40  # changing the future must not change earlier logits
    assert torch.equal(m(ids2)[:, :5], logits[:, :5])
    want = torch.nn.functional.cross_entropy(logits.view(-1, 40), batch["labels"].view(-1))
    assert abs(float(loss_fn(batch, m)) - float(want)) < 1e-12


@pytest.mark.parametrize("flatten", [None, False, True])
def test_model_forward_equals_an_operator_by_operator_restatement_of_the_reference(flatten):
    """DecoderLM.forward (residual sums handed to the next LayerNorm, optional fused q | k | v layout) against
    _mt_helpers.reference_forward, which is written straight from the reference's module formulas."""
    m = small_model()
    if flatten is not None:
        m.flatten_parameters(fuse=flatten)
    m.eval()
    idx = synthetic_batch(3, 8, 40, "cpu", torch.Generator().manual_seed(9))["input_ids"]
    got, want = m(idx), reference_forward(m, idx)
    assert float((got - want).abs().max()) <= 1e-11 * max(1.0, float(want.abs().max()))
    # a single layer through its public forward (the sum is formed here, not by a following LayerNorm)
    layer = m.t_layer_2
    x = torch.randn(2, 8, 16, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    s, f, f_bias = layer.forward_pending(x, None)
    assert f_bias is None  # CPU tensors: no bias gradient is handed on
    assert torch.allclose(layer(x), s + f, rtol=0, atol=1e-13)
    s2, f2, _ = layer.forward_pending(x * 0.25, x * 0.75)  # the layer applied to a pending sum
    assert torch.allclose(s2 + f2, layer(x), rtol=0, atol=1e-12)


def test_training_reduces_loss_on_a_fixed_batch():
    m = small_model()
    opt = Adam(m.parameters(), lr=0.02)
    gen = torch.Generator().manual_seed(2)
    batch = synthetic_batch(2, 8, 40, "cpu", gen)
    first = None
    for _ in range(12):
        opt.zero_grad()
        loss = loss_fn(batch, m)
        loss.backward()
        opt.step()
        first = first if first is not None else float(loss)
    assert float(loss) < 0.7 * first


def test_dropout_is_inverted_bernoulli_and_off_in_eval():
    x = torch.ones(200, 200, dtype=torch.float64)
    y = ops.dropout(x, 0.1, True, torch.Generator().manual_seed(0))
    kept = (y != 0).double().mean()
    assert abs(float(kept) - 0.9) < 0.01 and abs(float(y.max()) - 1 / 0.9) < 1e-12  # modules_basic.py:77-80
    assert ops.dropout(x, 0.1, False) is x


def test_flash_attention_op_refuses_cpu_tensors():
    q = torch.zeros(1, 1, 4, 8)
    with pytest.raises(RuntimeError):
        ops.flash_attention(q, q, q, True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_data_parallel_equals_single_process_on_the_union_batch():
    worker = os.path.join(ROOT, "tests", "_gloo_mt_worker.py")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, worker], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "dp==single True ranks_equal True overlap==single True" in outs[0], outs


def test_flat_parameter_adam_equals_per_parameter_adam():
    a, b = small_model(), small_model()
    oa = Adam(a.parameters(), lr=0.02)
    ob = Adam(b.parameters(), lr=0.02, flat=b.flatten_parameters())
    batch = synthetic_batch(2, 8, 40, "cpu", torch.Generator().manual_seed(1))
    for _ in range(3):
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad()
            loss_fn(batch, m).backward()
            o.step()
    fa = torch.cat([p.value.detach().reshape(-1) for p in a.parameters()])
    fb = torch.cat([p.value.detach().reshape(-1) for p in b.parameters()])
    assert torch.equal(fa, fb)
    assert b.parameters()[0].value.grad.data_ptr() == ob.flat[1].data_ptr()  # gradients live in the flat bucket


def test_fused_qkv_layout_of_flatten_parameters_equals_the_plain_model():
    """flatten_parameters(fuse=True) interleaves the q | k | v projection weights of every attention layer into one
    (in, 3E) matrix and puts their biases side by side (each parameter stays a view of it); the layer then runs ONE
    projection.  Same values, same loss and same gradients as the unflattened model, and an Adam step through the flat
    buffers moves the strided views."""
    a, b = small_model(), small_model()
    flat, gflat = b.flatten_parameters(fuse=True)
    att = b.t_layer_1.attention
    E = att.n_embd
    assert att._wqkv.shape == (E, 3 * E) and att._bqkv.shape == (3 * E,)
    assert not att.k_projection.weights.value.is_contiguous()  # a column slice of the fused matrix
    assert torch.equal(att._wqkv[:, E:2 * E], att.k_projection.weights.value)
    for (na, pa), (nb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert na == nb and torch.equal(pa.value.detach(), pb.value.detach()), na
    batch = synthetic_batch(2, 8, 40, "cpu", torch.Generator().manual_seed(4))
    la, lb = loss_fn(batch, a), loss_fn(batch, b)
    la.backward()
    lb.backward()
    assert abs(float(la) - float(lb)) <= 1e-12 * abs(float(la))
    for (name, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.allclose(pa.value.grad, pb.value.grad, rtol=1e-9, atol=1e-12), name
    assert torch.allclose(att._wqkv.grad[:, :E], att.q_projection.weights.value.grad)  # the fused gradient IS the per-projection one
    before = att.v_projection.weights.value.detach().clone()
    Adam(b.parameters(), lr=0.02, flat=(flat, gflat)).step()
    assert not torch.equal(before, att.v_projection.weights.value.detach())  # the flat step moved the strided view
