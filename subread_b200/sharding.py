"""Fragment sharding across GPUs: fragments are independent inside the voting
step, so a chunk's fragment range [0, n) is cut into contiguous slices, one per GPU; each GPU writes
its own slice of the host bigtable and there is no collective on the data path."""


def shard_range(n_frags: int, world: int, rank: int):
    """[first, last) of rank's slice; slice sizes differ by at most one fragment."""
    base, extra = divmod(int(n_frags), int(world))
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)
