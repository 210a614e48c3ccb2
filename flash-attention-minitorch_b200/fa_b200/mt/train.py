"""Training loop of the machine-translation run on synthetic token batches, data-parallel over ranks.

Mirrors project/run_machine_translation.py: model/optimizer construction (:332-365), the per-rank batch size
int((128*40 + L) / L) (:385), loss_fn (:158-177) and the train() loop with its forward / backward / optimizer
timers accumulated over iterations 2..9 of 10 (:180-221).  The IWSLT corpus and the BPE tokenizer are not
available offline, so batches are uniform random token ids of the same shape.

Data parallelism: one process per GPU, identical initial weights (same seed), per-rank batches, and the gradient
all-reduce (average) of the flat gradient buffer over NCCL (NVLink 5 / NVSwitch) through the thin C ABI
fa_b200_nccl_* of libfa_b200.so (fa_b200/mt/comm.py; no torch.distributed on the GPU path) -- in five buckets that
start from backward hooks as the layers' gradients complete.  The attention path itself has no collective.
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
    """Gradient all-reduce of the data-parallel run.

    With flattened parameters (Module.flatten_parameters) the flat gradient buffer IS the bucket storage -- no packing
    copies.  Two modes:
      * one all-reduce of the whole buffer after the backward (payload ~41 MB at the reference config);
      * `buckets` = [(a, b), ...] in completion order (gradient_buckets): bucket i is all-reduced from a backward hook the
        moment its gradients are complete (reduce_bucket, on a side stream), overlapping the rest of the backward;
        __call__ joins and reduces what is left.  in_graph: hooks and join were captured into the step's CUDA graph.
    Transport: `comm` (fa_b200.mt.comm.NcclComm, the thin C ABI fa_b200_nccl_* over NCCL / NVLink) for CUDA tensors;
    torch.distributed (gloo) stands in on the CPU, where the unit tests run."""

    def __init__(self, params: List, world: int, flat_grads: Optional[torch.Tensor] = None,
                 buckets: Optional[List[Tuple[int, int]]] = None, comm=None):
        self.params, self.world = params, world
        self.numel = sum(p.value.numel() for p in params)
        self.flat: Optional[torch.Tensor] = None
        self.flat_grads = flat_grads
        self.comm = comm
        self.buckets = list(buckets) if (buckets and flat_grads is not None and world > 1) else None
        self.in_graph = False
        self._done: List[int] = []  # buckets reduced by hooks in the current step
        self._works: List = []      # pending torch.distributed work handles (gloo path)
        self._side = None           # side stream of the NCCL path
        if self.buckets is not None and comm is not None:
            self._side = torch.cuda.Stream()

    def _reduce(self, t: torch.Tensor, async_op: bool = False):
        if self.comm is not None:
            self.comm.all_reduce(t, average=True)  # averaged inside the collective: no extra pass
            return None
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)  # (gloo: __call__ divides afterwards)

    def begin_step(self) -> None:
        """Start of a training step: no bucket has been reduced yet."""
        self._done = []

    def reduce_bucket(self, i: int) -> None:
        """Backward hook: bucket i is complete -- start its all-reduce next to the rest of the backward."""
        if self.buckets is None or i >= len(self.buckets) or i in self._done:
            return
        a, b = self.buckets[i]
        self._done.append(i)
        if self.comm is not None:
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                self.comm.all_reduce(self.flat_grads[a:b], average=True)
        else:
            self._works.append(self._reduce(self.flat_grads[a:b], async_op=True))

    def join(self) -> None:
        """Make the current stream (or the host, for gloo) wait for the bucket all-reduces started by the hooks."""
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        for w in self._works:
            w.wait()
        self._works = []

    def __call__(self) -> None:
        if self.world == 1:
            return
        if self.flat_grads is not None:
            if self.buckets is None:
                self._reduce(self.flat_grads)
            else:
                if not self.in_graph:
                    self.join()
                    rest = [i for i in range(len(self.buckets)) if i not in self._done]
                else:
                    rest = self._rest_in_graph
                for i in rest:
                    a, b = self.buckets[i]
                    self._reduce(self.flat_grads[a:b])
            if self.comm is None:
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
        if self.comm is not None:
            self.comm.all_reduce(self.flat, average=True)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(self.world)
        off = 0
        for p in self.params:
            n = p.value.numel()
            p.value.grad = self.flat[off:off + n].view_as(p.value).clone()
            off += n

    def mark_captured(self) -> None:
        """The hooks ran during CUDA-graph capture: every replay re-runs their all-reduces, __call__ only adds the rest."""
        self._rest_in_graph = [i for i in range(len(self.buckets or [])) if i not in self._done]
        self._done = []
        self.in_graph = True


def gradient_buckets(model: DecoderLM, flat_grads: torch.Tensor, seq_len: Optional[int] = None) -> Optional[List[Tuple[int, int]]]:
    """Ranges of the flat gradient buffer in the order in which the backward completes them: [final LayerNorm + lm_head],
    [layer 4], [layer 3], [layer 2], [layer 1], [embeddings] -- the hook DecoderLM.forward registers on the output of
    layer k fires when everything downstream of it is done (bucket index = 3 - k; the hook on the embedding stage's output
    is bucket 4); the last bucket is complete only when the backward ends and is reduced after it.  With `seq_len` the
    embedding bucket stops after row seq_len of the position table: the reference sizes that table n_vocab x n_embd
    (modules_transfomer.py DecoderLM), but rows >= seq_len never receive a gradient -- 9.7 of the 41 MB need no all-reduce.
    None if the layout is not the clustered one of Module.flatten_parameters."""
    r = model.flat_ranges()
    need = ["token_embeddings", "position_embeddings", "t_layer_1", "t_layer_2", "t_layer_3", "t_layer_4", "ln", "lm_head"]
    if any(k not in r for k in need):
        return None

    def span(names):
        a, b = min(r[n][0] for n in names), max(r[n][1] for n in names)
        return (a, b) if b - a == sum(r[n][1] - r[n][0] for n in names) else None

    emb = span(["token_embeddings", "position_embeddings"])
    out = [span(["ln", "lm_head"]), span(["t_layer_4"]), span(["t_layer_3"]), span(["t_layer_2"]), span(["t_layer_1"]), emb]
    if any(x is None for x in out) or sum(b - a for a, b in out) != flat_grads.numel():
        return None
    if seq_len is not None and r["position_embeddings"][0] > r["token_embeddings"][0]:
        used = r["position_embeddings"][0] + min(seq_len, model.n_vocab) * model.n_embd  # rows [0, seq_len) of the position table
        out[-1] = (emb[0], min(emb[1], (used + 31) // 32 * 32))
    return out


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
        early = reducer is not None and reducer.buckets is not None  # the bucket all-reduces (backward hooks) are captured too

        def fwd_bwd():
            if reducer is not None:
                reducer.begin_step()
            optimizer.zero_grad()
            loss = loss_fn(self.static, model)
            loss.backward()
            if early:
                reducer.join()  # the collective's stream joins the capturing stream again
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
            reducer.mark_captured()

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        for k, v in batch.items():
            self.static[k].copy_(v)
        self.ops.dropout_state.begin_step()
        self.graph.replay()
        return self.loss


def train(model: DecoderLM, optimizer: Adam, n_iters: int, batch_size: int, seq_len: int, rank: int, world: int,
          sync: Callable[[], None], seed: int = 0, log: Optional[Callable[[str], None]] = None,
          cuda_graph: bool = False, overlap_allreduce: bool = False, comm=None) -> Dict[str, float]:
    """Runs n_iters steps and returns the reference-style timers (iterations >= 2) plus tokens/s and losses.
    With cuda_graph the forward and backward are one graph replay, reported together as `backward_s` = 0."""
    from . import ops
    model.train()
    device = next(iter(model.parameters())).value.device
    gen = torch.Generator().manual_seed(1000 * (rank + 1) + seed)
    ops.dropout_state.base_seed = 0x5EED + 7919 * rank  # different masks on different ranks
    buckets = gradient_buckets(model, optimizer.flat[1], seq_len) if (overlap_allreduce and optimizer.flat is not None) else None
    reducer = GradAllReducer(model.parameters(), world, optimizer.flat[1] if optimizer.flat is not None else None, buckets, comm)
    model.grad_ready_hook = reducer.reduce_bucket if reducer.buckets is not None else None
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
            reducer.begin_step()
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


def job_min(value: float, rank: int, world: int, comm, device) -> float:
    """Minimum of a per-rank scalar over the job (one sum all-reduce of a one-hot vector through the C-ABI communicator)."""
    if world == 1 or comm is None:
        return float(value)
    t = torch.zeros(world, device=device, dtype=torch.float32)
    t[rank] = float(value)
    comm.all_reduce(t, average=False)
    torch.cuda.synchronize()
    return float(t.min().item())


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
    ap.add_argument("--overlap-allreduce", type=lambda s: s.lower() in ("1", "true", "yes"), default=True,
                    help="all-reduce the gradients bucket by bucket (final LayerNorm + lm_head, layer 4, 3, 2, then embeddings + "
                         "layer 1) from backward hooks, overlapped with the rest of the backward (inside the CUDA graph when "
                         "--cuda-graph); False: one all-reduce of the whole flat buffer after the backward")
    args = ap.parse_args(argv)
    if not args.use_flash_attention:
        raise SystemExit("only the flash-attention path exists here (the vanilla path is the reference's CPU oracle)")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this trainer has no CPU path")
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    comm = None
    if world > 1:
        from .comm import NcclComm
        comm = NcclComm().init(rank, world)  # fa_b200_nccl_* (include/fa_b200.h): no torch.distributed in this loop
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
                log=(lambda s: print(s, flush=True)) if rank == 0 else None, cuda_graph=args.cuda_graph,
                overlap_allreduce=args.overlap_allreduce, comm=comm)
    slowest = job_min(res["tokens_per_s_per_rank"], rank, world, comm, device)
    if rank == 0:
        # same line the reference prints (run_machine_translation.py:221), plus the job-level throughput
        print("N: {}, Forward Time: {}, Backward Time: {}, Opt Time: {}".format(L, res["forward_s"], res["backward_s"], res["opt_s"]))
        print(json.dumps({"metric": "decoder LM training tokens/s (whole job)", "value": slowest * world, "unit": "tokens/s",
                          "n_gpus": world, "seq_len": L, "batch_per_rank": batch_size, "iters_timed": res["timed_iters"],
                          "cuda_graph": bool(args.cuda_graph), "overlap_allreduce": bool(args.overlap_allreduce), "attention_precision": args.attention_precision,
                          "forward_s": res["forward_s"], "backward_s": res["backward_s"], "allreduce_s": res["allreduce_s"],
                          "opt_s": res["opt_s"], "loss_first": res["loss_first"], "loss_last": res["loss_last"],
                          "data": "synthetic uniform tokens", "params": sum(p.value.numel() for p in model.parameters())}), flush=True)
    if comm is not None:
        comm.destroy()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
