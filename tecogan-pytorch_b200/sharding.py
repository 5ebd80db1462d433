"""Clip-level sharding across ranks (reference codes/main.py:169: `for idx in range(rank,
num_seq, world_size)`).  The per-frame recurrence keeps a clip on one device, so there is no
data-path collective: ranks own disjoint clips."""


def clips_for_rank(num_clips, rank, world_size):
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank/world_size: {rank}/{world_size}')
    return list(range(rank, num_clips, world_size))
