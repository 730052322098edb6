"""Clip sharding across ranks (one process per GPU).  The transforms have no exchange step, so a
shard is just a contiguous clip range; results are bitwise independent of the sharding."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int):
    """Rank r of R owns items [floor(r*B/R), floor((r+1)*B/R))."""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    return (rank * n_items) // world, ((rank + 1) * n_items) // world
