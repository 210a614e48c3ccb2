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


def split_cpus(cpus, part: int, parts: int) -> list:
    """`parts` contiguous, disjoint slices of the sorted CPU list; slice `part` (sizes differ by at most one)."""
    cpus = sorted(cpus)
    lo, hi = shard_bounds(len(cpus), part, parts)
    return cpus[lo:hi]


def bind_to_gpu_numa(gpu_index: int, local_rank: int = 0, local_world: int = 1) -> int:
    """Pin this process to CPUs local to `gpu_index` (NVML affinity), sharing them DISJOINTLY with the other ranks of the
    node: on boxes where NVML reports the same affinity for every GPU (0-31 for all eight on the round-1 B200 box) each
    rank takes its own slice, so that eight copy pipelines do not fight for the same cores.  Pinned host buffers
    allocated afterwards are first-touched by these cores.  Best effort; returns the number of CPUs bound to (0 = nothing
    changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if local_world > 1 and len(cpus) >= local_world:
            # ranks whose GPUs report the same CPU set share it; count this rank's position among them
            same = [g for g in range(pynvml.nvmlDeviceGetCount())
                    if g < local_world and list(pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(g), words)) == list(mask)]
            if local_rank in same and len(cpus) >= len(same):
                cpus = set(split_cpus(cpus, same.index(local_rank), len(same)))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0
