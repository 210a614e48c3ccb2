"""Batch x head sharding.  Every (b, h) pair is an independent attention problem (reference:
blockIdx.x = b*H + h, src/flash_attn2_fw.cu:25-35; README.md:12), so the flattened B*H axis is cut
into `world` contiguous chunks and no collective touches the attention path."""
from __future__ import annotations


def shard_bounds(batch_heads: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of the B*H axis owned by `rank`; chunk sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(batch_heads, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(batch_heads: int, world: int) -> list[int]:
    return [e - b for b, e in (shard_bounds(batch_heads, r, world) for r in range(world))]


def bind_to_gpu_numa(gpu_index: int) -> int:
    """Pin this process to the CPUs NVML reports as local to `gpu_index` (best effort; returns the number of CPUs bound
    to, 0 if nothing was changed).  One process per GPU: pinned host buffers allocated afterwards are first-touched on
    the GPU's own NUMA node, so host<->device copies of different ranks do not meet on the inter-socket link."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0
