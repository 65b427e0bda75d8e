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


def tabu_tenures(n, k):
    """The reference's tenure schedule: tabu_init (metaheuristic.c:65-84: MIN_FRACTION 0.125, MAX_FRACTION 0.25) followed by
    k steps of tabu_linear_policy (:40-59) -- pure integer bookkeeping that depends on n and k only."""
    lo, hi = int(0.125 * n), int(0.25 * n)
    tenure, inc, out = int(0.125 * n + 1), True, []
    for _ in range(k):
        if tenure == hi or tenure == lo:
            inc = not inc
        tenure += 1 if inc else -1
        out.append(tenure)
    return out


def tabu_restarts(inst, first_start, count, stride, k):
    """Config 5 "restarts" (a framework addition -- the reference has one trajectory): restart r = start node
    first_start + r*stride: NN tour -> 2-opt local optimum -> k tabu iterations (linear tenure), all device-resident.
    Returns this shard's best (cost, start, path) under the incumbent rule (strict '<', ascending starts), or
    (None, -1, None) for an empty shard; reduce_best() merges the shards."""
    import numpy as np
    best_cost, best_start, best_path = None, -1, None
    tenures = tabu_tenures(inst.n, k)
    for r in range(count):
        start = first_start + r * stride
        code, path, cost = inst.nn(start)
        code, moves, path, cost, _ = inst.two_opt(path)
        tl = np.full(inst.n, -1, dtype=np.int32)
        run_best, run_path = cost, path.copy()
        final, trace, bi, bp, bc = inst.tabu_search(path, cost, tl, tenures, incumbent_cost=cost)
        if bi >= 0:
            run_best, run_path = bc, bp
        if best_cost is None or run_best < best_cost:
            best_cost, best_start, best_path = run_best, start, run_path
    return best_cost, best_start, best_path


def vns_restarts(inst, first_start, count, stride, k, seed=0):
    """Config 5 VNS "restarts" (a framework addition, as tabu_restarts): restart r = start node first_start + r*stride with
    the glibc stream re-seeded to srand(seed + start) (the reference seeds only random instances, tsp.c:469): NN tour ->
    k iterations of the loop of mh_VNS (metaheuristic.c:277-321: 2-opt descent, incumbent test, rand() % 9 - 2 kicks),
    tour resident on the device.  A restart depends only on (instance, start, seed), never on the shard it runs in.
    Returns this shard's best (cost, start, path) under the incumbent rule (strict '<', ascending starts), or
    (None, -1, None) for an empty shard; reduce_best() merges the shards."""
    import ctypes
    libc = ctypes.CDLL(None)
    best_cost, best_start, best_path = None, -1, None
    for r in range(count):
        start = first_start + r * stride
        libc.srand(ctypes.c_uint(seed + start))
        code, path, cost = inst.nn(start)
        # the cost to beat is +inf: the first local optimum always enters (mh_VNS starts from the incumbent's cost, :270,
        # because its start tour already is the incumbent; a restart has no incumbent yet)
        code, path, cost, trace, bi, bp, bc, done = inst.vns_search(path, k, float("inf"))
        if bi >= 0 and (best_cost is None or bc < best_cost):
            best_cost, best_start, best_path = bc, start, bp
    return best_cost, best_start, best_path
