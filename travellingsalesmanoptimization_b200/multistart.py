"""Host-side plumbing of multi-GPU multi-start work (one process per GPU, torch.distributed).

The path shards by START NODE: starts are independent units (heuristics.c:42,82 loop over them), so
rank r of W takes starts r, r+W, r+2W, ... (block-cyclic keeps descent lengths balanced) and there
is NO data-path collective.  The only exchange is the incumbent: one all-reduce(min) of a packed
64-bit key and one all-reduce(sum) that carries the winning successor array (zeros elsewhere).

Key = cost << 20 | start.  Costs are integers < 2^42, starts < 2^20 (n <= 131072): the integer order
of the key is (cost, start) lexicographic, i.e. the reference's strict-'<' incumbent rule evaluated in
ascending start order (tsp.c:671, heuristics.c:58/82): the LOWEST start wins cost ties.
"""

START_BITS = 20


def shard_starts(first_start, nstarts, rank, world):
    """(first, count, stride) of the starts owned by `rank`: first_start + (rank + k*world)."""
    count = (nstarts - rank + world - 1) // world if nstarts > rank else 0
    return first_start + rank, count, world


def pack_key(cost, start):
    cost = int(cost)
    assert 0 <= cost < (1 << 42) and 0 <= start < (1 << START_BITS)
    return (cost << START_BITS) | int(start)


def unpack_key(key):
    return int(key) >> START_BITS, int(key) & ((1 << START_BITS) - 1)


NO_RESULT = (1 << 63) - 1


def reduce_best(dist, device, local_cost, local_start, local_path, first_start=0):
    """All ranks call this with their local best (or cost=None when they own no start).
    Returns (best_cost, best_start, best_path tensor on `device`) on every rank."""
    import torch
    world = dist.get_world_size()
    key = torch.tensor([NO_RESULT if local_cost is None else pack_key(local_cost, local_start)],
                       dtype=torch.int64, device=device)
    dist.all_reduce(key, op=dist.ReduceOp.MIN)
    cost, start = unpack_key(key.item())
    owner = (start - first_start) % world
    # The winner's successor array travels as a second all-reduce (sum of the winner's array and
    # zeros from everybody else): same communicator, same ring/NVLS channels as the min-reduce, so no
    # lazily-built per-root broadcast connections appear inside a timed region.
    path = local_path.to(device) if dist.get_rank() == owner else torch.zeros_like(local_path, device=device)
    dist.all_reduce(path, op=dist.ReduceOp.SUM)
    return cost, start, path
