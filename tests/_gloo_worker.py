"""Worker for test_host_logic.test_two_rank_gloo_*: world_size-2 gloo run of the sharded path on CPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-attention-minitorch_b200"))
from fa_b200 import sharding  # noqa: E402
from oracle import attention_oracle as ao  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    BH, T, d = 6, 33, 16
    q, k, v, g = ao.make_inputs(BH, T, d, seed=11)
    b, e = sharding.shard_bounds(BH, rank, world)
    dq, dk, dv, o, L = ao.attention_backward(q[b:e], k[b:e], v[b:e], g[b:e], True, np.float64)
    # the timing reduction bench.py performs: MAX over ranks of the per-rank elapsed time
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    parts = [None] * world
    dist.all_gather_object(parts, (b, e, o, dq))
    if rank == 0:
        fdq, fdk, fdv, fo, fL = ao.attention_backward(q, k, v, g, True, np.float64)
        so = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])])
        sdq = np.concatenate([p[3] for p in sorted(parts, key=lambda p: p[0])])
        print("sharded==full", bool(np.array_equal(so, fo) and np.array_equal(sdq, fdq)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
