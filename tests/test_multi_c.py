"""The in-process multi-device scheduler of the C ABI (pmn_multi, include/pmnucmer.h): one process, several GPUs — the
reference's `-cores N` (lib/base/paramugsy.ml:54-57, lib/base/queued_task_server.ml:57-64) over devices.

CPU: the pair plan (pure host arithmetic).  -m gpu: the same device named several times stands in for several GPUs on a
one-GPU box, so the whole path — per-device schedulers, peer copies of the anchor lists, index copy — runs on it."""
import os

import pytest

from paramugsy_b200 import lib, multi, synth


def all_pairs(n):
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


@pytest.mark.parametrize("n,nd", [(8, 1), (8, 2), (8, 4), (8, 8), (57, 8), (3, 8), (2, 2)])
def test_plan_covers_every_pair_and_keeps_references_together(n, nd):
    pairs = all_pairs(n)
    dev = lib.multi_plan(nd, pairs)
    assert len(dev) == len(pairs) and all(0 <= d < nd for d in dev)
    sizes = [dev.count(d) for d in range(nd)]
    assert max(sizes) - min(sizes) <= 1 or len(pairs) < nd
    # pairs in reference order are cut into contiguous runs: device numbers never decrease along that order
    order = sorted(range(len(pairs)), key=lambda k: (pairs[k][0], k))
    assert [dev[k] for k in order] == sorted(dev[k] for k in order)
    # the same rule as the torch.distributed form of the scheduler
    a = multi.assign_pairs(pairs, nd)
    assert all(dev[k] == r for r, ks in enumerate(a) for k in ks)
    # distinct indexes a device needs: few (C2 on 8 GPUs: at most 3, the last device takes the short tails of three references)
    if (n, nd) == (8, 8):
        assert max(len({pairs[k][0] for k in range(len(pairs)) if dev[k] == d}) for d in range(nd)) <= 3


def test_plan_follows_genome_sizes():
    pairs = all_pairs(6)
    nb = [10_000_000] + [1_000_000] * 5                # genome 0 is ten times the others
    dev = lib.multi_plan(3, pairs, nb)
    load = [sum(nb[a] + nb[b] for k, (a, b) in enumerate(pairs) if dev[k] == d) for d in range(3)]
    assert max(load) <= sum(load) / 3 + 11_000_000
    with pytest.raises(lib.PmnError):
        lib.multi_plan(2, [(0, 9)], [1, 1])


@pytest.mark.gpu
def test_all_vs_all_over_several_schedulers_equals_one(oracle):
    gs = [synth.random_genome(120_000, 300)]
    gs += [synth.invert(synth.mutate(gs[0], 0.03, 301 + k), 1, 5_000, 401 + k) for k in range(5)]
    fa = [synth.fasta(f"m{k}.1", g) for k, g in enumerate(gs)]
    names = [f"m{k}.fa" for k in range(6)]
    pairs = all_pairs(6)
    with lib.Scheduler(0, 4) as s:
        one = [(r.delta, r.filtered, r.maf) for r in s.align_fasta(fa, pairs, names=names, post=1)]
    for devices in ([0], [0, 0], [0, 0, 0, 0]):
        with lib.Multi(devices, workers=2) as m:
            got = [(r.delta, r.filtered, r.maf) for r in m.align_fasta(fa, pairs, names=names, post=1)]
        assert got == one, f"results depend on the number of devices ({len(devices)})"
    assert one[0][0] == oracle.nucmer(fa[0], fa[1], "m0.fa", "m1.fa")
    assert one[-1][0] == oracle.nucmer(fa[4], fa[5], "m4.fa", "m5.fa")


@pytest.mark.gpu
def test_files_over_several_schedulers(tmp_path, oracle):
    g0 = synth.random_genome(60_000, 500)
    fa = [synth.fasta("f0.1", g0)] + [synth.fasta(f"f{k}.1", synth.mutate(g0, 0.02, 501 + k)) for k in (1, 2, 3)]
    paths = []
    for k, f in enumerate(fa):
        p = tmp_path / f"f{k}"; p.write_bytes(f); paths.append(str(p))
    pairs = all_pairs(4)
    refs, qrys = [paths[a] for a, _ in pairs], [paths[b] for _, b in pairs]
    outs = [str(tmp_path / f"{a}-{b}.delta") for a, b in pairs]; mafs = [str(tmp_path / f"{a}-{b}.maf") for a, b in pairs]
    with lib.Multi([0, 0], workers=2) as m:
        m.align_files(refs, qrys, outs, mafs, post=1)
        for (a, b), o, mf in zip(pairs, outs, mafs):
            d = oracle.nucmer(fa[a], fa[b], paths[a], paths[b])
            f = oracle.delta_filter(d, 1)
            assert open(o, "rb").read() == f and open(mf, "rb").read() == oracle.delta2maf(f, fa[a], fa[b])
        with pytest.raises(lib.PmnError):                 # any failing pair fails the batch (lib/base/job_processor.ml:72-73)
            m.align_files(refs + [str(tmp_path / "missing")], qrys + [paths[0]], outs + [str(tmp_path / "x.delta")])


@pytest.mark.gpu
@pytest.mark.parametrize("index", ["build", "copy"])
def test_large_pair_over_several_devices_is_byte_identical(oracle, index, monkeypatch):
    monkeypatch.setenv("PMN_MULTI_INDEX", index)
    g0 = synth.random_genome(400_000, 600)
    g1 = synth.invert(synth.mutate(g0, 0.02, 601), 3, 20_000, 602)
    ref, qry = synth.fasta("big0.1", g0), synth.fasta("big1.1", g1)
    want = oracle.nucmer(ref, qry, "big0.fa", "big1.fa")
    for devices in ([0], [0, 0], [0, 0, 0, 0, 0]):
        with lib.Multi(devices, workers=1) as m:
            res, ms = m.align_large(ref, qry, "big0.fa", "big1.fa")
            assert res.delta == want, f"{len(devices)} devices, index {index}"
            assert len(ms) == 4 and all(x >= 0 for x in ms)
            res.close()


def test_the_c_plan_and_the_torch_distributed_plan_are_the_same_rule():
    """bench.py shards by pmn_multi_plan while multi.AllVsAll (index broadcast) uses assign_pairs: the same cut for any sizes."""
    import random
    rnd = random.Random(7)
    for _ in range(100):
        n, world = rnd.randint(2, 12), rnd.randint(1, 8)
        nb = [rnd.randint(1_000_000, 6_000_000) for _ in range(n)]
        pairs = all_pairs(n)
        dev = lib.multi_plan(world, pairs, nb)
        a = multi.assign_pairs(pairs, world, [nb[x] + nb[y] for x, y in pairs])
        assert all(dev[k] == r for r, ks in enumerate(a) for k in ks), (n, world)
