"""Training loop of the machine-translation run on synthetic token batches, data-parallel over ranks.

Mirrors project/run_machine_translation.py: model/optimizer construction (:332-365), the per-rank batch size
int((128*40 + L) / L) (:385), loss_fn (:158-177) and the train() loop with its forward / backward / optimizer
timers accumulated over iterations 2..9 of 10 (:180-221).  The IWSLT corpus and the BPE tokenizer are not
available offline, so batches are uniform random token ids of the same shape.

Data parallelism: one process per GPU, identical initial weights (same seed), per-rank batches, and ONE
collective per step -- an all-reduce(sum)/world of the flattened gradients over NCCL (NVLink 5 / NVSwitch).
The attention path itself has no collective (heads are independent).
"""
from __future__ import annotations

import argparse
import json
import os
import time
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .model import DecoderLM, loss_fn
from .optim import Adam


def synthetic_batch(batch_size: int, seq_len: int, n_vocab: int, device, gen: torch.Generator) -> Dict[str, torch.Tensor]:
    """Random token ids; labels are the inputs shifted by one; every label position weighted 1 (collate_batch,
    run_machine_translation.py:107-155, without the source/target split of real sentence pairs)."""
    tokens = torch.randint(0, n_vocab, (batch_size, seq_len + 1), generator=gen).to(device)
    return {"input_ids": tokens[:, :-1].contiguous(), "labels": tokens[:, 1:].contiguous(),
            "label_token_weights": torch.ones(batch_size, seq_len, device=device)}


class GradAllReducer:
    """One flat bucket for all gradients: a single all-reduce per step (payload ~36 MB at the reference config,
    latency-bound on NVSwitch, so one launch beats per-parameter calls).  With flattened parameters
    (Module.flatten_parameters) the bucket IS the gradient storage: no packing copies."""

    def __init__(self, params: List, world: int, flat_grads: Optional[torch.Tensor] = None,
                 early: Optional[Tuple[int, int]] = None):
        self.params, self.world = params, world
        self.numel = sum(p.value.numel() for p in params)
        self.flat: Optional[torch.Tensor] = None
        self.flat_grads = flat_grads
        # Optional overlap (flat buffers only): the range `early` = [a, b) of the flat gradient buffer -- the final
        # LayerNorm and lm_head, 10 of the 37 MB, complete after the first ~30 % of the backward -- is all-reduced from a
        # backward hook (reduce_early, asynchronously on the collective's stream) while the rest of the backward runs;
        # __call__ then reduces what is left.  in_graph: reduce_early / join_early were captured into the step's CUDA
        # graph, so __call__ has nothing to wait for.
        self.early = early if (early is not None and flat_grads is not None and world > 1) else None
        self.in_graph = False
        self._work = None

    def _reduce(self, t: torch.Tensor, async_op: bool = False):
        if dist.get_backend() == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)  # averaged inside the collective: no extra pass
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)  # (gloo: __call__ divides afterwards)

    def reduce_early(self) -> None:
        """Backward hook: the early range is complete -- start its all-reduce."""
        if self.early is not None:
            self._work = self._reduce(self.flat_grads[self.early[0]:self.early[1]], async_op=True)

    def join_early(self) -> None:
        """Make the current stream (or the host, for gloo) wait for the early all-reduce."""
        if self._work is not None:
            self._work.wait()
            self._work = None

    def __call__(self) -> None:
        if self.world == 1:
            return
        if self.flat_grads is not None:
            if self.early is None:
                self._reduce(self.flat_grads)
            else:
                if not self.in_graph:
                    self.join_early()
                a, b = self.early
                for lo, hi in ((0, a), (b, self.flat_grads.numel())):
                    if hi > lo:
                        self._reduce(self.flat_grads[lo:hi])
            if dist.get_backend() != "nccl":
                self.flat_grads.div_(self.world)
            return
        p0 = self.params[0].value
        if self.flat is None:
            self.flat = torch.empty(self.numel, device=p0.device, dtype=p0.dtype)
        off = 0
        for p in self.params:
            n = p.value.numel()
            g = p.value.grad
            if g is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(g.reshape(-1))
            off += n
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.div_(self.world)
        off = 0
        for p in self.params:
            n = p.value.numel()
            p.value.grad = self.flat[off:off + n].view_as(p.value).clone()
            off += n


def early_gradient_range(model: DecoderLM, flat_grads: torch.Tensor) -> Optional[Tuple[int, int]]:
    """[a, b) of the flat gradient buffer holding the final LayerNorm and lm_head gradients (None if they are not one
    contiguous run of it): the part of the model whose gradients are complete first."""
    ps = [p for m in (model.ln, model.lm_head) for p in m.parameters()]
    esz = flat_grads.element_size()
    offs = sorted(((p.value.grad.data_ptr() - flat_grads.data_ptr()) // esz, p.value.numel()) for p in ps)
    a, b = offs[0][0], offs[-1][0] + offs[-1][1]
    # consecutive up to the alignment padding of Module.flatten_parameters (< 32 zero elements between two parameters)
    gaps_ok = all(0 <= nxt[0] - (cur[0] + cur[1]) < 32 for cur, nxt in zip(offs, offs[1:]))
    if a < 0 or b > flat_grads.numel() or not gaps_ok or any(not p.value.grad.is_contiguous() for p in ps):
        return None
    return a, b


class GraphedStep:
    """zero_grad + forward + backward of one training step captured ONCE into a CUDA graph and replayed per step.

    The eager step is launch-bound (a few hundred kernels of a few microseconds each behind Python dispatch); the graph
    replays them back to back.  What makes the step capturable: parameters and gradients live in two flat static
    buffers (Module.flatten_parameters), batches are copied into static input tensors, workspaces come from the graph's
    private pool, and the dropout masks take their per-step variation from a device-side counter
    (ops.dropout_state.step) instead of a host-side seed.  The gradient all-reduce (NCCL) and the fused Adam launch
    stay outside the graph (Adam's bias-corrected step size changes every step)."""

    def __init__(self, model: DecoderLM, optimizer: Adam, batch: Dict[str, torch.Tensor], warmup: int = 3,
                 reducer: Optional[GradAllReducer] = None):
        from . import ops
        assert optimizer.flat is not None, "graph capture needs flattened parameters"
        self.model, self.optimizer, self.ops = model, optimizer, ops
        self.static = {k: v.clone() for k, v in batch.items()}
        early = reducer is not None and reducer.early is not None  # the early all-reduce (a backward hook) is captured too

        def fwd_bwd():
            optimizer.zero_grad()
            loss = loss_fn(self.static, model)
            loss.backward()
            if early:
                reducer.join_early()  # the collective's stream joins the capturing stream again
            return loss

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up off the capture: lazy initialisations, autotuning, arena growth
            for _ in range(warmup):
                fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="relaxed"):
            self.loss = fwd_bwd()
        if early:
            reducer.in_graph = True

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        for k, v in batch.items():
            self.static[k].copy_(v)
        self.ops.dropout_state.begin_step()
        self.graph.replay()
        return self.loss


def train(model: DecoderLM, optimizer: Adam, n_iters: int, batch_size: int, seq_len: int, rank: int, world: int,
          sync: Callable[[], None], seed: int = 0, log: Optional[Callable[[str], None]] = None,
          cuda_graph: bool = False, overlap_allreduce: bool = False) -> Dict[str, float]:
    """Runs n_iters steps and returns the reference-style timers (iterations >= 2) plus tokens/s and losses.
    With cuda_graph the forward and backward are one graph replay, reported together as `backward_s` = 0."""
    from . import ops
    model.train()
    device = next(iter(model.parameters())).value.device
    gen = torch.Generator().manual_seed(1000 * (rank + 1) + seed)
    ops.dropout_state.base_seed = 0x5EED + 7919 * rank  # different masks on different ranks
    early = early_gradient_range(model, optimizer.flat[1]) if (overlap_allreduce and optimizer.flat is not None) else None
    reducer = GradAllReducer(model.parameters(), world, optimizer.flat[1] if optimizer.flat is not None else None, early)
    model.grad_ready_hook = reducer.reduce_early if reducer.early is not None else None
    graphed = None
    fw = bw = comm = opt = 0.0
    timed_tokens, timed_s, losses = 0, 0.0, []
    for it in range(n_iters):
        batch = synthetic_batch(batch_size, seq_len, model.n_vocab, device, gen)
        if cuda_graph and graphed is None:
            graphed = GraphedStep(model, optimizer, batch, reducer=reducer)
        sync()
        t0 = time.perf_counter()
        if graphed is not None:
            loss = graphed(batch)
            sync()
            t1 = t2 = time.perf_counter()
        else:
            ops.dropout_state.begin_step()
            optimizer.zero_grad()
            loss = loss_fn(batch, model)
            sync()
            t1 = time.perf_counter()
            loss.backward()
            sync()
            t2 = time.perf_counter()
        reducer()
        sync()
        t3 = time.perf_counter()
        optimizer.step()
        sync()
        t4 = time.perf_counter()
        losses.append(float(loss.item()))
        if it > 1:  # the reference skips its first two iterations (run_machine_translation.py:207-210)
            fw += t1 - t0
            bw += t2 - t1
            comm += t3 - t2
            opt += t4 - t3
            timed_tokens += batch_size * seq_len
            timed_s += t4 - t0
        if log:
            log(f"iter {it} loss {losses[-1]:.4f} tokens/s/rank {batch_size * seq_len / (t4 - t0):.0f}")
    return {"forward_s": fw, "backward_s": bw, "allreduce_s": comm, "opt_s": opt, "tokens_per_s_per_rank": timed_tokens / max(timed_s, 1e-9),
            "timed_iters": max(n_iters - 2, 0), "loss_first": losses[0], "loss_last": losses[-1], "losses": losses}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="decoder-only LM training with B200 flash attention (synthetic tokens)")
    ap.add_argument("--model-max-length", type=int, default=512)
    ap.add_argument("--use-flash-attention", type=lambda s: s.lower() in ("1", "true", "yes"), default=True)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--n-vocab", type=int, default=10000)
    ap.add_argument("--n-embd", type=int, default=256)
    ap.add_argument("--n-head", type=int, default=8)
    ap.add_argument("--learning-rate", type=float, default=0.02)
    ap.add_argument("--p-dropout", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=11111)
    ap.add_argument("--attention-precision", choices=["fp32", "bf16"], default="fp32",
                    help="fp32: tf32 tensor-core forward and backward on the fp32 tensors (1e-3 bar, default); bf16: forward and "
                         "backward on the bf16 tensor-core kernels (2e-2 bar)")
    ap.add_argument("--cuda-graph", type=lambda s: s.lower() in ("1", "true", "yes"), default=True,
                    help="capture forward+backward into a CUDA graph (default); False gives the reference's fw/bw split timers")
    ap.add_argument("--overlap-allreduce", type=lambda s: s.lower() in ("1", "true", "yes"), default=False,
                    help="all-reduce the final LayerNorm + lm_head gradients from a backward hook, overlapped with the rest of "
                         "the backward (inside the CUDA graph when --cuda-graph)")
    args = ap.parse_args(argv)
    if not args.use_flash_attention:
        raise SystemExit("only the flash-attention path exists here (the vanilla path is the reference's CPU oracle)")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this trainer has no CPU path")
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    L = args.model_max_length
    batch_size = int((128 * 40 + L) / L)  # run_machine_translation.py:385, per rank
    from . import ops
    model = DecoderLM(n_vocab=args.n_vocab, n_embd=args.n_embd, n_head=args.n_head, n_positions=L, p_dropout=args.p_dropout,
                      ln_eps=1e-5, bias=True, device=device, seed=args.seed,
                      attention_fn=ops.flash_attention_bf16 if args.attention_precision == "bf16" else ops.flash_attention)
    optimizer = Adam(model.parameters(), lr=args.learning_rate, flat=model.flatten_parameters())

    def sync():
        torch.cuda.synchronize()

    res = train(model, optimizer, args.iters, batch_size, L, rank, world, sync,
                log=(lambda s: print(s, flush=True)) if rank == 0 else None, cuda_graph=args.cuda_graph, overlap_allreduce=args.overlap_allreduce)
    t = torch.tensor([res["tokens_per_s_per_rank"]], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        # same line the reference prints (run_machine_translation.py:221), plus the job-level throughput
        print("N: {}, Forward Time: {}, Backward Time: {}, Opt Time: {}".format(L, res["forward_s"], res["backward_s"], res["opt_s"]))
        print(json.dumps({"metric": "decoder LM training tokens/s (whole job)", "value": float(t.item()) * world, "unit": "tokens/s",
                          "n_gpus": world, "seq_len": L, "batch_per_rank": batch_size, "iters_timed": res["timed_iters"],
                          "cuda_graph": bool(args.cuda_graph), "overlap_allreduce": bool(args.overlap_allreduce), "attention_precision": args.attention_precision,
                          "forward_s": res["forward_s"], "backward_s": res["backward_s"], "allreduce_s": res["allreduce_s"],
                          "opt_s": res["opt_s"], "loss_first": res["loss_first"], "loss_last": res["loss_last"],
                          "data": "synthetic uniform tokens", "params": sum(p.value.numel() for p in model.parameters())}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
