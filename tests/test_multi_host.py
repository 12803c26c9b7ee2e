"""Host-side logic of the multi-GPU path (paramugsy_b200/multi.py) on CPU: pair assignment, index
ownership, and the two collectives over gloo at world_size 2 with CPU tensors standing in for HBM."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from paramugsy_b200 import multi
from paramugsy_b200.mugsy_nucmer import searches


def all_pairs(n):
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


@pytest.mark.parametrize("n,world", [(8, 1), (8, 2), (8, 4), (8, 8), (57, 8), (3, 8), (2, 2)])
def test_assignment_covers_every_pair_once_and_is_balanced(n, world):
    pairs = all_pairs(n)
    a = multi.assign_pairs(pairs, world)
    flat = sorted(k for r in a for k in r)
    assert flat == list(range(len(pairs)))
    sizes = [len(r) for r in a]
    assert max(sizes) - min(sizes) <= 1 or len(pairs) < world
    # contiguous in reference order: a rank's references form an interval
    for r in a:
        refs = [pairs[k][0] for k in r]
        assert refs == sorted(refs)
    assert a == multi.assign_pairs(pairs, world)          # deterministic


def test_assignment_follows_cost():
    pairs = all_pairs(6)
    cost = [10 if i == 0 else 1 for i, _ in pairs]         # genome 0 is huge
    a = multi.assign_pairs(pairs, 3, cost)
    load = [sum(cost[k] for k in r) for r in a]
    assert max(load) <= sum(cost) / 3 + 10


def test_index_plan_builds_each_reference_once():
    pairs = all_pairs(8)
    for world in (1, 2, 4, 8):
        a = multi.assign_pairs(pairs, world)
        plan = multi.index_plan(pairs, a)
        assert sorted(plan) == sorted({p[0] for p in pairs})
        for ref, (owner, ranks) in plan.items():
            assert 0 <= owner < world
            for r, ks in enumerate(a):
                assert (r in ranks) == any(pairs[k][0] == ref for k in ks)
        builds = [sum(1 for o, _ in plan.values() if o == r) for r in range(world)]
        assert max(builds) <= -(-len(plan) // world)


def test_pair_enumeration_matches_pm_job():
    # lib/base/pm_job.ml:43-51: ordered pairs, the earlier genome is the reference
    assert searches(["a", "b", "c"]) == [("a", "b"), ("a", "c"), ("b", "c")]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- index replication: images are byte tensors; the owner's content must arrive at every consumer
        pairs = all_pairs(5)
        a = multi.assign_pairs(pairs, world)
        plan = multi.index_plan(pairs, a)
        size = {ref: 1000 + 37 * ref for ref in plan}
        truth = {ref: (torch.arange(size[ref]) * (ref + 3) % 251).to(torch.uint8) for ref in plan}
        mine = {ref: truth[ref].clone() for ref, (o, _) in plan.items() if o == rank}
        recv = {ref: torch.zeros(size[ref], dtype=torch.uint8) for ref, (o, rs) in plan.items() if o != rank and rank in rs}
        sunk = []
        def sink(ref):
            sunk.append(ref); return torch.empty(size[ref], dtype=torch.uint8)
        got = multi.replicate_images(plan, rank, lambda r: mine[r], lambda r: recv[r], sink, dist)
        assert sorted(got) == sorted(recv)
        for ref, t in recv.items():
            assert torch.equal(t, truth[ref])
        for ref, (o, rs) in plan.items():
            assert (ref in sunk) == (rs != [o] and rank not in rs and rank != o)
        # ---- anchors of a sharded seeding: ragged all-gather, rank order, an empty contribution
        n = [7, 0][rank] if world == 2 else rank
        local = torch.arange(n * 4, dtype=torch.int32).view(n, 4) + 1000 * rank
        allr = multi.gather_concat(local, dist)
        want = torch.cat([torch.arange(k * 4, dtype=torch.int32).view(k, 4) + 1000 * r for r, k in enumerate([7, 0] if world == 2 else range(world))])
        assert torch.equal(allr, want)
        empty = multi.gather_concat(torch.empty((0, 4), dtype=torch.int32), dist)
        assert empty.shape == (0, 4)
        q.put((rank, "ok"))
    except Exception as e:          # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_collectives_over_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
