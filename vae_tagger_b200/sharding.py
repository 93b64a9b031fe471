"""Batch sharding of an image stream over the GPUs of one box: contiguous index
ranges, no collective on the inference path."""


def shard_range(num_items: int, rank: int, world: int):
    """Half-open index range of ``rank``: sizes differ by at most one, ranges tile [0, num_items)."""
    base, extra = divmod(num_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
