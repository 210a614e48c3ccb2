"""Worker for test_mt_model.test_two_rank_gloo_data_parallel_*: two CPU ranks train the decoder LM with the gradient
all-reduce of fa_b200.mt.train (gloo standing in for NCCL, a torch restatement standing in for the CUDA attention)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-minitorch_b200"))
from _mt_helpers import vanilla_attention  # noqa: E402
from fa_b200.mt.model import DecoderLM, loss_fn  # noqa: E402
from fa_b200.mt.optim import Adam  # noqa: E402
from fa_b200.mt.train import GradAllReducer, synthetic_batch, train  # noqa: E402


def build():
    return DecoderLM(n_vocab=50, n_embd=16, n_head=2, n_positions=8, p_dropout=0.0, device="cpu", dtype=torch.float64,
                     seed=7, attention_fn=vanilla_attention)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    model = build()
    opt = Adam(model.parameters(), lr=0.02)
    res = train(model, opt, n_iters=3, batch_size=3, seq_len=8, rank=rank, world=world, sync=lambda: None, seed=5)
    # the same with flattened parameters (fused q | k | v layout) and the final-LayerNorm + lm_head gradients all-reduced
    # from a backward hook while the rest of the backward runs (train(overlap_allreduce=True))
    m2 = build()
    flat2 = m2.flatten_parameters(fuse=True)
    train(m2, Adam(m2.parameters(), lr=0.02, flat=flat2), n_iters=3, batch_size=3, seq_len=8, rank=rank, world=world,
          sync=lambda: None, seed=5, overlap_allreduce=True)
    assert m2.grad_ready_hook is not None  # the early range was found and used
    overlap_flat = torch.cat([p.value.detach().reshape(-1) for p in m2.parameters()])
    # every rank must hold identical weights after synchronised steps
    flat = torch.cat([p.value.detach().reshape(-1) for p in model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(flat, ref))
    # and the result must equal a single process that sees both ranks' batches and averages the two losses
    if rank == 0:
        single = build()
        sopt = Adam(single.parameters(), lr=0.02)
        gens = [torch.Generator().manual_seed(1000 * (r + 1) + 5) for r in range(world)]
        for _ in range(3):
            sopt.zero_grad()
            losses = [loss_fn(synthetic_batch(3, 8, 50, "cpu", g), single) for g in gens]
            (sum(losses) / world).backward()
            sopt.step()
        sflat = torch.cat([p.value.detach().reshape(-1) for p in single.parameters()])
        print("dp==single", bool(torch.allclose(flat, sflat, rtol=1e-9, atol=1e-11)), "ranks_equal", same,
              "overlap==single", bool(torch.allclose(overlap_flat, sflat, rtol=1e-9, atol=1e-11)),
              "loss", res["loss_first"], res["loss_last"])
    else:
        assert same
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
