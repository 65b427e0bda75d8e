"""CPU, world_size 2, gloo: the N>1 host logic of multi-start work (sharding by start node + the
packed-key all-reduce(min) + winner all-reduce) reproduces the reference's incumbent rule."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from travellingsalesmanoptimization_b200 import multistart as MS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, per_start, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = len(per_start)
    first, count, stride = MS.shard_starts(0, n, rank, world)
    best = None
    for k in range(count):
        s = first + k * stride
        c = per_start[s][2]
        if best is None or c < best[0]:          # strict '<' in ascending start order, tsp.c:671
            best = (c, s)
    path = torch.full((8,), best[1] if best else -1, dtype=torch.int32)      # stand-in tour tagged with its start
    cost, start, bpath = MS.reduce_best(dist, torch.device("cpu"), best[0] if best else None, best[1] if best else 0, path)
    out_q.put((rank, cost, start, bpath.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["berlin52", "kroA100"])
def test_two_rank_reduction_matches_reference_incumbent(name):
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    g = golden[name]["greedy_2opt"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, g["per_start"], q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, cost, start, bpath in res:
        assert (cost, start) == (g["cost"], g["first_start"])
        assert bpath == [g["first_start"]] * 8        # every rank received the winner's tour


def test_shard_and_key():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            first, count, stride = MS.shard_starts(3, 21, r, world)
            seen += [first + k * stride for k in range(count)]
        assert sorted(seen) == list(range(3, 24))
    # lexicographic (cost, start): lower start wins a cost tie, lower cost wins otherwise
    assert MS.pack_key(100, 7) < MS.pack_key(100, 9) < MS.pack_key(101, 0)
    assert MS.unpack_key(MS.pack_key(148543747, 85899)) == (148543747, 85899)
    # a rank that owns no start cannot win
    assert MS.NO_RESULT > MS.pack_key((1 << 42) - 1, (1 << 20) - 1)


def test_tabu_tenure_schedule_matches_reference_policy(oracle):
    """multistart.tabu_tenures (host logic of the sharded tabu restarts) against the oracle's restatement of
    tabu_init + tabu_linear_policy (metaheuristic.c:40-84, pinned on the reference objects elsewhere)."""
    from travellingsalesmanoptimization_b200 import multistart as MS
    for n in (8, 52, 100, 2392, 15112, 20000):
        tenure, lo, hi = oracle.tabu_init(n)
        inc, ref = True, []
        for _ in range(3 * max(1, hi - lo) + 7):
            tenure, inc = oracle.tabu_linear_step(tenure, lo, hi, inc)
            ref.append(tenure)
        assert MS.tabu_tenures(n, len(ref)) == ref, n
