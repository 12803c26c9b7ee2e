"""Several GPUs of one box, one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The path is embarrassingly parallel by genome pair (the reference runs one
`mugsy_nucmer` process per pair, lib/base/job_processor.ml:128-154), so there is no data-path
collective except the two that north_star names:

  * all-vs-all: pairs are dealt to the ranks; every reference index is built ONCE, by its owner, and
    replicated to the ranks that hold pairs of that reference with a broadcast of the index image
    (one contiguous range of HBM, sent and received in place);
  * one large pair: index and query are replicated, the query positions are sharded for seeding, the
    ranks' anchor lists are all-gathered (their concatenation in rank order is the ordered anchor list
    of the undivided run), and the rest runs replicated, so the .delta does not depend on the GPU count.

Everything here is plumbing over torch tensors; it runs unchanged over gloo with CPU tensors, which is
how tests/test_multi_host.py covers it at world_size 2.
"""
from collections import defaultdict

from . import lib


def assign_pairs(pairs, world, cost=None):
    """Deal `pairs` [(ref, qry)] to `world` ranks: pairs are kept in reference order and cut into `world`
    contiguous runs of (nearly) equal cost, so that a rank needs as few distinct indexes as possible.
    Returns a list of pair-index lists, one per rank.  Deterministic."""
    order = sorted(range(len(pairs)), key=lambda k: (pairs[k][0], k))
    w = [float(cost[k]) if cost is not None else 1.0 for k in order]
    total = sum(w)
    out = [[] for _ in range(world)]
    acc = 0.0
    for k, wk in zip(order, w):
        # the rank whose interval [r*total/world, (r+1)*total/world) holds the pair's midpoint
        r = min(world - 1, int((acc + wk / 2) * world / total)) if total > 0 else 0
        out[r].append(k)
        acc += wk
    return out


def index_plan(pairs, assignment):
    """For every reference: (owner rank, sorted consumer ranks).  Any rank may own (build) an index; the
    owner is the rank with the fewest indexes to build so far, consumers first among equals, then the
    lowest rank — so the builds are spread evenly even though the last references have few pairs."""
    world = len(assignment)
    users = defaultdict(set)
    for r, ks in enumerate(assignment):
        for k in ks:
            users[pairs[k][0]].add(r)
    builds = [0] * world
    plan = {}
    for ref in sorted(users):
        owner = min(range(world), key=lambda r: (builds[r], 0 if r in users[ref] else 1, r))
        builds[owner] += 1
        plan[ref] = (owner, sorted(users[ref]))
    return plan


def replicate_images(plan, rank, my_image, recv_image, sink, dist, group=None):
    """Broadcast every reference's index image from its owner.
    my_image(ref) -> tensor to send (owner), recv_image(ref) -> tensor to receive into (consumer),
    sink(nbytes) -> scratch tensor for ranks that hold no pair of that reference.  All ranks call this
    with the same plan; returns the list of references this rank received."""
    got, works = [], []
    for ref in sorted(plan):
        owner, ranks = plan[ref]
        if ranks == [owner]:
            continue                                    # nobody else needs it
        if rank == owner:
            t = my_image(ref)
        elif rank in ranks:
            t = recv_image(ref); got.append(ref)
        else:
            t = sink(ref)
        works.append(dist.broadcast(t, src=owner, group=group, async_op=True))
    for w in works:
        w.wait()
    return got


def gather_concat(local, dist, group=None):
    """All-gather of row blocks of different lengths: every rank contributes an (n_r, C) tensor and gets
    the (sum n_r, C) concatenation in rank order."""
    import torch
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts + [1])
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


class AllVsAll:
    """All-vs-all over the ranks of a process group: one step = every pair of the batch aligned once,
    every reference index built once in the whole job."""

    def __init__(self, sched, seqs, names, pairs, rank, world, dist=None, group=None, cost=None):
        self.sched, self.seqs, self.names, self.pairs = sched, seqs, names, pairs
        self.rank, self.world, self.dist, self.group = rank, world, dist, group
        self.assignment = assign_pairs(pairs, world, cost)
        self.plan = index_plan(pairs, self.assignment)
        self.mine = self.assignment[rank]
        self._sink = None

    def step(self):
        """Returns {pair index: Result} for this rank's pairs."""
        import torch
        ctx = self.sched.context(0)
        held = {}
        for ref, (owner, ranks) in sorted(self.plan.items()):
            if owner == self.rank:
                held[ref] = self.seqs[ref].index()                 # built here, sent below if anybody else needs it
            elif self.rank in ranks:
                held[ref] = self.seqs[ref].index(empty=True)       # received below
        if self.world > 1:
            def sink(ref):
                n = lib.index_image_bytes(self.seqs[ref].bases)      # a function of the reference length alone
                if self._sink is None or self._sink.numel() < n:
                    self._sink = torch.empty(n, dtype=torch.uint8, device=torch.device("cuda", ctx.device))
                return self._sink[:n]
            got = replicate_images(self.plan, self.rank, lambda r: held[r].image_tensor(), lambda r: held[r].image_tensor(), sink, self.dist, self.group)
            torch.cuda.synchronize()        # the workers' streams are not ordered after torch's NCCL stream
            for ref in got:
                held[ref].adopt()
        idx = [held.get(g) for g in range(len(self.seqs))]
        my_pairs = [self.pairs[k] for k in self.mine]
        res = self.sched.align_seqs(self.seqs, my_pairs, names=self.names, indexes=idx) if my_pairs else []
        for ix in held.values():
            ix.close()
        return dict(zip(self.mine, res))


def align_large_pair(ref_index, qry, rank, world, dist, group=None, ref_path="ref.fa", qry_path="qry.fa", **opts):
    """One large pair on `world` GPUs: every rank holds the index and the packed query; rank r seeds its
    range of query positions; anchors are all-gathered; clustering and extension run replicated.  Every
    rank returns the same Result (byte-identical .delta for any world size)."""
    local = ref_index.seed_part_tensor(qry, rank, world, **opts)
    anchors = gather_concat(local, dist, group) if world > 1 else local
    if anchors.is_cuda:
        import torch
        torch.cuda.synchronize()            # the context's stream is not ordered after torch's streams
    return ref_index.align_anchors(qry, anchors, ref_path=ref_path, qry_path=qry_path, **opts)
