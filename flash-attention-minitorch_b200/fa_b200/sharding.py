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
