"""Decoder-LM slice on the B200 (`-m gpu`): the model running on this repo's attention kernels against the same model
running a plain torch restatement of causal attention, and a short training run."""
import pytest
import torch

from _mt_helpers import vanilla_attention
from fa_b200 import _lib
from fa_b200.mt import ops
from fa_b200.mt.model import DecoderLM, loss_fn
from fa_b200.mt.optim import Adam
from fa_b200.mt.train import synthetic_batch, train

pytestmark = pytest.mark.gpu


def build(dev, attention_fn, n_embd=64, n_head=2, p_dropout=0.0):
    return DecoderLM(n_vocab=300, n_embd=n_embd, n_head=n_head, n_positions=96, p_dropout=p_dropout, device=dev,
                     dtype=torch.float32, seed=5, attention_fn=attention_fn)


@pytest.mark.parametrize("n_embd,n_head", [(64, 2), (128, 2), (96, 2)])  # head dims 32 (the MT model's), 64, 48
def test_model_on_flash_kernels_matches_model_on_vanilla_attention(cuda_device, n_embd, n_head):
    """Forward loss and every parameter gradient: flash path (tf32 tensor-core forward where covered, exact-fp32
    backward) vs softmax(QK^T/sqrt(d) + M)V in fp32 -- the comparison of the reference's
    kernel_tests/test_flashattn_comb.py, at model level."""
    flash = build(cuda_device, ops.flash_attention, n_embd, n_head)
    plain = build(cuda_device, vanilla_attention, n_embd, n_head)
    gen = torch.Generator().manual_seed(9)
    batch = synthetic_batch(3, 96, 300, cuda_device, gen)
    before = _lib.launch_count()
    lf, lp = loss_fn(batch, flash), loss_fn(batch, plain)
    lf.backward()
    lp.backward()
    assert _lib.launch_count() - before >= 4 * 3  # 4 layers x (fwd + bwd kernels) actually ran in libfa_b200.so
    assert abs(float(lf) - float(lp)) <= 1e-3 * abs(float(lp))
    for (name, a), (_, b) in zip(flash.named_parameters(), plain.named_parameters()):
        ga, gb = a.value.grad, b.value.grad
        assert float((ga - gb).abs().max()) <= 2e-3 * float(gb.abs().max()) + 1e-6, name


def test_short_training_run_on_gpu_reduces_loss(cuda_device):
    m = build(cuda_device, ops.flash_attention, p_dropout=0.1)
    opt = Adam(m.parameters(), lr=0.01)
    gen = torch.Generator().manual_seed(2)
    batch = synthetic_batch(4, 96, 300, cuda_device, gen)
    losses = []
    for _ in range(15):
        opt.zero_grad()
        loss = loss_fn(batch, m)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.8 * losses[0], losses


def test_train_loop_reports_reference_style_timers(cuda_device):
    m = build(cuda_device, ops.flash_attention, p_dropout=0.1)
    res = train(m, Adam(m.parameters(), lr=0.02), n_iters=4, batch_size=2, seq_len=96, rank=0, world=1,
                sync=torch.cuda.synchronize)
    assert res["timed_iters"] == 2 and res["forward_s"] > 0 and res["backward_s"] > 0 and res["opt_s"] > 0
    assert res["tokens_per_s_per_rank"] > 0


def test_kernel_backed_model_matches_fp64_torch_model(cuda_device):
    """Every operator on this repo's kernels (tf32 GEMMs, fused LayerNorm / GELU / embedding / loss, flash attention)
    against the same model evaluated with torch in fp64 on the CPU: loss within 1e-3 relative, gradients within 2 % of
    each tensor's max (tf32 has 10 mantissa bits; four layers deep)."""
    dev_model = build(cuda_device, ops.flash_attention, n_embd=64, n_head=2)
    ref_model = DecoderLM(n_vocab=300, n_embd=64, n_head=2, n_positions=96, p_dropout=0.0, device="cpu", dtype=torch.float64,
                          seed=5, attention_fn=vanilla_attention)
    for (_, a), (_, b) in zip(dev_model.named_parameters(), ref_model.named_parameters()):
        assert torch.equal(a.value.detach().cpu().double(), b.value.detach())  # same seed, same init
    gen = torch.Generator().manual_seed(9)
    batch = synthetic_batch(4, 96, 300, "cpu", gen)
    dev_batch = {k: v.to(cuda_device) for k, v in batch.items()}
    dev_batch["label_token_weights"] = dev_batch["label_token_weights"].float()
    batch["label_token_weights"] = batch["label_token_weights"].double()
    ld, lr = loss_fn(dev_batch, dev_model), loss_fn(batch, ref_model)
    ld.backward()
    lr.backward()
    assert abs(float(ld) - float(lr)) <= 1e-3 * abs(float(lr))
    for (name, a), (_, b) in zip(dev_model.named_parameters(), ref_model.named_parameters()):
        ga, gb = a.value.grad.double().cpu(), b.value.grad
        # (the key-projection bias has an identically zero gradient -- softmax is shift invariant -- hence the floor)
        assert float((ga - gb).abs().max()) <= 2e-2 * float(gb.abs().max()) + 1e-5, name


@pytest.mark.parametrize("n_embd,n_head", [(64, 2), (128, 4), (256, 8)])  # 128 / 256: the register-resident LayerNorm kernels
def test_flattened_fast_path_matches_fp64_torch_model(cuda_device, n_embd, n_head):
    """The configuration the training script runs: flattened parameters (q | k | v weights interleaved, ONE head-layout
    projection GEMM), sequence length a multiple of 128 (head-layout GEMMs), head dim 32 (tf32 tensor-core attention
    forward AND backward), weight / bias / LayerNorm / embedding gradients accumulated by the kernels straight into the
    flat gradient views; at n_embd 128 / 256 also the residual sums and branch gradients inside the LayerNorm kernels and
    the out_projection / linear_out / linear_in / lm_head bias gradients inside the LayerNorm, GELU and loss backward
    passes -- against the same model in fp64 torch on the CPU, parameter by parameter."""
    T = 128  # d = 32
    dev_model = DecoderLM(n_vocab=300, n_embd=n_embd, n_head=n_head, n_positions=T, p_dropout=0.0, device=cuda_device,
                          dtype=torch.float32, seed=5, attention_fn=ops.flash_attention)
    ref_model = DecoderLM(n_vocab=300, n_embd=n_embd, n_head=n_head, n_positions=T, p_dropout=0.0, device="cpu",
                          dtype=torch.float64, seed=5, attention_fn=vanilla_attention)
    flat, gflat = dev_model.flatten_parameters()
    att = dev_model.t_layer_1.attention
    assert att._wqkv.shape == (n_embd, 3 * n_embd) and not att.q_projection.weights.value.is_contiguous()
    for (_, a), (_, b) in zip(dev_model.named_parameters(), ref_model.named_parameters()):
        assert torch.equal(a.value.detach().cpu().double(), b.value.detach())  # flattening kept every value
    gen = torch.Generator().manual_seed(19)
    batch = synthetic_batch(3, T, 300, "cpu", gen)
    dev_batch = {k: v.to(cuda_device) for k, v in batch.items()}
    dev_batch["label_token_weights"] = dev_batch["label_token_weights"].float()
    batch["label_token_weights"] = batch["label_token_weights"].double()
    from fa_b200 import _lib
    before = _lib.launch_count()
    gflat.zero_()
    ld, lr = loss_fn(dev_batch, dev_model), loss_fn(batch, ref_model)
    ld.backward()
    lr.backward()
    assert _lib.launch_count() - before > 80  # this repo's kernels ran (GEMMs, attention, LayerNorm, loss, ...; 97 with every fusion on)
    assert abs(float(ld) - float(lr)) <= 1e-3 * abs(float(lr))
    for (name, a), (_, b) in zip(dev_model.named_parameters(), ref_model.named_parameters()):
        ga, gb = a.value.grad.double().cpu(), b.value.grad
        assert float((ga - gb).abs().max()) <= 2e-2 * float(gb.abs().max()) + 1e-5, name
    # the fused views alias the per-projection gradients
    assert torch.equal(att._wqkv.grad[:, :n_embd], att.q_projection.weights.value.grad)


def test_flat_parameters_fused_adam_training_step(cuda_device):
    m = build(cuda_device, ops.flash_attention, p_dropout=0.1)
    flat = m.flatten_parameters()
    opt = Adam(m.parameters(), lr=0.01, flat=flat)
    gen = torch.Generator().manual_seed(2)
    batch = synthetic_batch(4, 96, 300, cuda_device, gen)
    losses = []
    for _ in range(15):
        opt.zero_grad()
        loss = loss_fn(batch, m)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.8 * losses[0], losses
    assert m.parameters()[0].value.grad.data_ptr() == flat[1].data_ptr()  # autograd accumulated in place


def test_cuda_graph_step_equals_eager_step_and_redraws_dropout(cuda_device):
    """One captured graph replayed per step: without dropout it reproduces the eager trajectory; with dropout every
    replay draws a fresh mask (device-side step counter), so two replays on the same batch give different losses."""
    from fa_b200.mt.train import GraphedStep
    gen = torch.Generator().manual_seed(3)
    batches = [synthetic_batch(4, 96, 300, cuda_device, gen) for _ in range(4)]
    traj = {}
    for mode in ("eager", "graph"):
        m = build(cuda_device, ops.flash_attention, p_dropout=0.0)
        opt = Adam(m.parameters(), lr=0.01, flat=m.flatten_parameters())
        step = GraphedStep(m, opt, batches[0]) if mode == "graph" else None
        losses = []
        for b in batches:
            if step is not None:
                loss = step(b)
            else:
                opt.zero_grad()
                loss = loss_fn(b, m)
                loss.backward()
            opt.step()
            losses.append(float(loss))
        traj[mode] = (losses, opt.flat[0].clone())
    # split-K weight gradients are fp32 reductions in arbitrary order, and Adam turns the sign of a rounding-noise
    # gradient (the key-projection bias has a zero true gradient) into a full +-lr step: compare the losses and the
    # bulk of the parameters, not every element
    assert max(abs(a - b) for a, b in zip(*[traj[k][0] for k in ("eager", "graph")])) <= 2e-3
    diff = (traj["eager"][1] - traj["graph"][1]).abs()
    assert float(diff.median()) <= 1e-5 and float((diff > 1e-3).float().mean()) <= 0.02
    m = build(cuda_device, ops.flash_attention, p_dropout=0.1)
    opt = Adam(m.parameters(), lr=0.0, flat=m.flatten_parameters())
    step = GraphedStep(m, opt, batches[0])
    l1, l2 = float(step(batches[0])), float(step(batches[0]))
    assert l1 != l2


@pytest.mark.parametrize("n_embd,n_head", [(64, 2), (128, 2), (256, 2)])  # head dims 32 (padded to 64), 64, 128
def test_bf16_attention_path_of_the_model_within_the_bf16_bar(cuda_device, n_embd, n_head):
    """fp32 model tensors through the bf16 tensor-core attention (forward and backward on tcgen05): gradients of every
    parameter within 2e-2 of the tensor max against the same model on vanilla fp32 attention."""
    fast = build(cuda_device, ops.flash_attention_bf16, n_embd, n_head)
    plain = build(cuda_device, vanilla_attention, n_embd, n_head)
    batch = synthetic_batch(3, 96, 300, cuda_device, torch.Generator().manual_seed(9))
    lf, lp = loss_fn(batch, fast), loss_fn(batch, plain)
    lf.backward()
    lp.backward()
    assert abs(float(lf) - float(lp)) <= 2e-3 * abs(float(lp))
    for (name, a), (_, b) in zip(fast.named_parameters(), plain.named_parameters()):
        assert float((a.value.grad - b.value.grad).abs().max()) <= 2e-2 * float(b.value.grad.abs().max()) + 1e-5, name
