#!/usr/bin/env python
"""bench.py — the pairwise nucmer stage on N B200s, one JSON line.

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...          # the CPU path (oracle port) on the host cores
    python bench.py --config c4 ...               # the single 100 Mbp pair, query-chunk partitioned over the ranks

Workload (BASELINE.json configs[1]): 8 synthetic 5 Mbp genomes (2 % divergence from a common ancestor, two 50 kbp
inversions each), all-vs-all = 28 (reference, query) pairs, the earlier genome being the reference
(lib/base/pm_job.ml:43-51).  One STEP = one pass of the hot path over that batch: every reference index built, 28 x
(seeding, clustering, extension, .delta text).

  value   pairs/s with the packed genomes already resident in HBM (index builds are inside).
  e2e     the same through the C ABI from FASTA bytes in pinned HOST memory to .delta bytes in HOST memory: parse + H2D +
          pack of the genomes and D2H of every result are inside.
  N > 1   the 28 pairs are SHARDED over the ranks (pmn_multi_plan: the pair list in reference order cut into N runs), every
          rank packs the genomes and builds the indexes its pairs name: no collective in the data path ("scaling": "strong",
          the batch is fixed).  `extra_weak` in the same line is every rank running the whole batch (N replicas).
          --replicate broadcast builds every index once in the job and sends its image to the consumers over NCCL instead.
  parity  before anything is timed, the .delta of all 28 pairs is compared with the oracle's committed digests
          (tests/golden/golden_configs.json) and, in the cpu_baseline leg, byte for byte with the oracle run on the host.
"""
import argparse
import hashlib
import gc
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
# one hardware work queue per stream of the scheduler (default 8: the 16+ streams of the workers alias and serialise), read by
# the driver when the CUDA context is created: before anything touches the GPU
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# stdout carries the one JSON line and nothing else: libraries that print to file descriptor 1 (NCCL's version banner
# is a plain printf) are sent to stderr, the JSON line goes to a private copy of the original stdout
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    print(json.dumps(obj), file=_REAL_STDOUT, flush=True)

from paramugsy_b200 import synth  # noqa: E402

METRIC = "nucmer_pair_alignments_per_s"
UNIT = "pairs/s"
PARITY_NOTE = ("bit-exact against the in-repo CPU oracle (oracle/pmn_oracle.c, a restatement of MUMmer 3.20's nucmer); the oracle "
               "itself is UNPINNED against MUMmer, which the reference does not vendor")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload(args):
    genomes = synth.config_c2(n=args.genome_bp, count=args.genomes, inv_len=max(1000, args.genome_bp // 100))
    fastas = [(name, synth.fasta(name, seq)) for name, seq in genomes]
    pairs = [(i, j) for i in range(len(genomes)) for j in range(i + 1, len(genomes))]
    return genomes, fastas, pairs


def golden(cfg):
    p = os.path.join(ROOT, "tests", "golden", "golden_configs.json")
    return json.load(open(p)).get(cfg, {}) if os.path.exists(p) else {}


def sha(b):
    return hashlib.sha256(b).hexdigest()


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    # no power.draw: in tools/outlier_probe.py the only long e2e steps (35-39 ms instead of 13 / 19, one in 150) came with it in the
    # query; a 10-step timed region with one such step is 10-30 % off.  (It is not the only source: see DESIGN.md §6.)
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "400", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def wait_first_sample(self, timeout=15.0):
        """nvidia-smi takes up to a second to come up and holds driver locks while it does: the
        warm-up and the timed region start only after its first sample has arrived."""
        t = time.time()
        while self.proc and not self.rows and time.time() - t < timeout and self.proc.poll() is None:
            time.sleep(0.05)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Ranks:
    """torch.distributed plumbing shared by the two workloads: barrier, device-timed regions, max over ranks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, counters=None, allocs=None):
        """K calls of fn bracketed by barrier + synchronize, timed with CUDA events on the (idle) current stream: every call
        returns with all worker streams drained, so the two events bracket exactly the device work.  -> (ms max over ranks, info)"""
        torch, dist = self.torch, self.dist
        self.barrier()
        c0 = counters() if counters else {}; a0 = allocs() if allocs else 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # no cyclic garbage collection inside the timed region (what timeit does): the path measured is a C library, and a
        # full collection of this process (torch imported, every .delta of the parity pass alive) is tens of milliseconds —
        # one such pause in a region of ten 12 ms steps is a fifth of the number
        gc.collect(); gc.disable()
        e0.record()
        walls = []
        for _ in range(steps):
            t = time.perf_counter(); fn(); walls.append((time.perf_counter() - t) * 1e3)
        e1.record()
        gc.enable()
        self.barrier()
        ms = e0.elapsed_time(e1)
        ranks_ms = [round(ms / steps, 3)]
        if self.world > 1:
            allms = torch.zeros(self.world, device="cuda", dtype=torch.float64); allms[self.rank] = ms
            dist.all_reduce(allms)                              # every rank's own time (diagnostic: stragglers)
            ranks_ms = [round(float(x) / steps, 3) for x in allms.tolist()]
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        c1 = counters() if counters else {}
        d = {k: c1[k] - c0[k] for k in c0}
        d["walls"] = [round(w, 2) for w in walls]; d["allocs"] = (allocs() - a0) if allocs else 0; d["ranks_ms"] = ranks_ms
        return ms, d

    def sum(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def spread(walls):
    return {"min": min(walls), "median": statistics.median(walls), "max": max(walls), "n": len(walls)} if walls else None


def ncu_traffic():
    """DRAM bytes per launch of the kernels, from this round's committed `ncu --set full` capture of one C2 pair."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p))["kernels"], f"profiles/{name} (ncu --set full --clock-control none, one C2 pair)"
    return {}, None


# ------------------------------------------------------------------------------------------ GPU arm, C2

def run_gpu(args):
    from paramugsy_b200 import lib
    R = Ranks(args)
    torch, rank, world, local = R.torch, R.rank, R.world, R.local
    if args.workers_auto:
        # a worker holds 3-4 GB of scratch, traceback slots and arena: no more workers than the free memory takes
        free_b, _ = torch.cuda.mem_get_info()
        args.workers = max(4, min(args.workers, int((free_b / 2**30 - 24) // 4.5)))
    sched = lib.Scheduler(local, args.workers)
    ctx = sched.context(0)

    genomes, fastas, pairs = workload(args)
    names = [g[0] + ".fa" for g in genomes]
    nbytes = [len(f[1]) for f in fastas]
    # end-to-end arm: the FASTA text of every genome sits in PINNED host memory and is copied to the device every step
    pinned = [torch.frombuffer(bytearray(f[1]), dtype=torch.uint8).pin_memory() for f in fastas]
    fasta_bytes = [(t.data_ptr(), t.numel()) for t in pinned]
    mode = args.mode if args.mode != "auto" else ("strong" if world > 1 else "single")
    sharded = mode == "strong" and world > 1
    dev_of = lib.multi_plan(world, pairs, nbytes) if sharded else [rank] * len(pairs)
    my_pairs = [p for p, d in zip(pairs, dev_of) if d == rank]
    needed = sorted({g for p in my_pairs for g in p})
    refs = sorted({i for i, _ in my_pairs})
    total_bp_in = sum(len(genomes[i][1]) + len(genomes[j][1]) for i, j in my_pairs)
    broadcast = sharded and args.replicate == "broadcast"
    if broadcast:
        from paramugsy_b200 import multi
        needed = list(range(len(genomes)))       # an owner may build an index none of its own pairs names

    resident = {g: ctx.sequence(fastas[g][1]) for g in needed}
    seq_list = [resident.get(g) for g in range(len(genomes))]
    all_pairs_weak = pairs

    def run_batch(plist, from_fasta, keep=None, post=0, deltas=False):
        if broadcast and not from_fasta and plist is my_pairs:
            out = multi.AllVsAll(sched, seq_list, names, pairs, rank, world, R.dist, cost=[nbytes[a] + nbytes[b] for a, b in pairs]).step()
            res = [out[k] for k in sorted(out)]
        elif not plist:
            res = []
        elif from_fasta:
            res = sched.align_fasta(fasta_bytes, plist, names=names, post=post)
        else:
            res = sched.align_seqs(seq_list, plist, names=names)
        for r in res:
            if keep is not None:
                keep.append((r.delta if deltas else None, r.stats_raw))      # the raw struct: turned into numbers after the timed region
            r.close()

    # ---- parity before timing: every pair of this rank against the oracle's committed digests (tests/golden)
    gold = golden("c2") if (args.genome_bp, args.genomes) == (5_000_000, 8) else {}
    first = []
    run_batch(my_pairs, False, keep=first, deltas=True)
    checked, bad = 0, []
    my_delta = {}
    for (i, j), (d, st) in zip(my_pairs, first):
        my_delta[(i, j)] = d
        g = gold.get(f"{genomes[i][0]}-{genomes[j][0]}")
        if g:
            checked += 1
            if sha(d) != g["delta_sha256"]:
                bad.append(f"{genomes[i][0]}-{genomes[j][0]}")
    if bad:
        raise SystemExit(f"bench.py: PARITY FAILURE on rank {rank}: .delta of {bad} differs from the oracle's digest")
    (checked_all,) = R.sum(checked)

    def c_sched():
        return sched.counters()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # before the warm-up: nvidia-smi's own start-up must not land in the timed region
        sampler.wait_first_sample()
    R.barrier()
    # ---- resident arm
    for _ in range(args.warmup):
        run_batch(my_pairs, False)
    timed_stats = []
    ms_res, cnt_res = R.timed(lambda: run_batch(my_pairs, False, keep=timed_stats), args.steps, c_sched, lib.alloc_count)
    # ---- end-to-end arm (host FASTA bytes -> host .delta bytes)
    for _ in range(args.warmup):
        run_batch(my_pairs, True)
    ms_e2e, cnt_e2e = R.timed(lambda: run_batch(my_pairs, True), args.steps, c_sched, lib.alloc_count)
    # ---- the whole reference worker per pair: nucmer + delta-filter -1 + delta2maf in one call (pmn_opts.post)
    for _ in range(args.warmup):
        run_batch(my_pairs, True, post=1)
    ms_wrk, cnt_wrk = R.timed(lambda: run_batch(my_pairs, True, post=1), args.steps, c_sched, lib.alloc_count)
    clocks = sampler.stop() if rank == 0 else None
    free_b, total_b = torch.cuda.mem_get_info()
    mem_used_gb = (total_b - free_b) / 2**30
    # ---- N > 1: the replica mode next to the sharded one (every rank runs the whole batch)
    weak = None
    if sharded:
        for g in range(len(genomes)):
            if g not in resident:
                resident[g] = ctx.sequence(fastas[g][1])
        seq_list = [resident.get(g) for g in range(len(genomes))]
        for _ in range(args.warmup):
            run_batch(all_pairs_weak, False)
        ms_w, _ = R.timed(lambda: run_batch(all_pairs_weak, False), args.steps)
        for _ in range(args.warmup):
            run_batch(all_pairs_weak, True)
        ms_we, _ = R.timed(lambda: run_batch(all_pairs_weak, True), args.steps)
        weak = {"what": "every rank runs the whole 28-pair batch on its own GPU (N independent replicas, no sharding)", "scaling": "weak",
                "value": len(pairs) * world * args.steps / (ms_w * 1e-3), "e2e": len(pairs) * world * args.steps / (ms_we * 1e-3), "unit": UNIT,
                "ms_per_step": ms_w / args.steps}
    # ---- one instrumented pass, one pair at a time (kernels alone on the GPU), and the INT32 roof
    alone = []
    for i in refs:
        ix = resident[i].index()
        for (a, b) in my_pairs:
            if a == i:
                res = ix.align(resident[b], ref_path=names[a], qry_path=names[b]); alone.append((a, b, res.stats, len(res.delta))); res.close()
        ix.close()
    int32_gops, _ = ctx.int32_peak()

    npairs_rank = len(my_pairs)
    npairs_all, bp_all, aligned_all = R.sum(npairs_rank, total_bp_in, sum(d[2]["aligned_ref_bases"] for d in alone))
    if rank == 0:
        hbm_peak, peak_src = peaks()
        stats_in = [st.as_dict() for _, st in timed_stats]      # every pair of every timed step, CUDA-event times taken under load
        n_in = max(1, len(stats_in))
        S_in = lambda k: sum(st[k] for st in stats_in)
        S = lambda k: sum(d[2][k] for d in alone)
        aligned_bp = S("aligned_ref_bases")
        step_ms = ms_res / args.steps
        seen = {}
        for a, b, st, _ in alone:
            seen.setdefault(a, st["ms_index"])
        stage_alone = {"index": sum(seen.values()), "seed": S("ms_seed"), "cluster": S("ms_cluster"), "extend": S("ms_extend")}
        stage_in = {"seed": S_in("ms_seed") / args.steps, "cluster": S_in("ms_cluster") / args.steps, "extend": S_in("ms_extend") / args.steps}
        traffic, traffic_src = ncu_traffic()
        if args.genome_bp != 5_000_000:
            traffic, traffic_src = {}, None
        # Roofline objects.  `achieved` / `frac` use the kernel's duration ALONE on the GPU (CUDA events on its stream in the one-pair-at-a-time
        # pass of this same run): a roofline assumes the kernel has the machine.  `under_load` are the same events taken inside the timed
        # steps, where up to 32 pairs share the GPU and an interval on one stream includes waiting for SMs other pairs hold — reported, but
        # not a kernel duration.  `whole_step` is the aggregate: all launches of the kernel in a step over the step's duration.
        n_pairs_alone = max(1, len(alone))
        # ---- seeding: bytes the kernel's own algorithm needs, per launch
        #   per looked-up position: K-mer table entry pair 8 B + suffix-array entry 4 B + one 8-byte reference word + 1 B of the skip table;
        #   per block probe: one 4-byte word of the presence bitmap;
        #   per position (looked up, settled by a probe or stepped over): 0.375 B of packed query (text + mask bits); 16 B per anchor written.
        b_own = sum(st["seed_lookups"] * 21.0 + st["seed_probes"] * 4.0 + 2 * st["qry_bases"] * 0.375 + 16 * st["anchors"] for _, _, st, _ in alone) / n_pairs_alone
        b_model = sum(2 * st["qry_bases"] * (12 * math.ceil(math.log2(max(2, st["ref_bases"]))) + 8.25) + 12 * st["anchors"] for _, _, st, _ in alone) / n_pairs_alone
        seed_ms_alone, seed_ms_in = S("ms_seed_kernel") / n_pairs_alone, S_in("ms_seed_kernel") / n_in
        lookups = sum(st["seed_lookups"] for _, _, st, _ in alone); positions = sum(2 * st["qry_bases"] for _, _, st, _ in alone)
        gbs = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9 if ms else None
        frac_of = lambda x, peak: x / peak if x else None
        seed_traffic = (traffic.get("k_seed") or {}).get("dram_bytes_per_launch")
        roof_seed = {"kernel": "k_seed", "bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": b_own, "launch_ms": seed_ms_alone,
                     "achieved": gbs(b_own, seed_ms_alone), "frac": frac_of(gbs(b_own, seed_ms_alone), hbm_peak),
                     "under_load": {"launch_ms": seed_ms_in, "achieved": gbs(b_own, seed_ms_in), "frac": frac_of(gbs(b_own, seed_ms_in), hbm_peak)},
                     "whole_step": {"achieved": gbs(b_own * npairs_rank, step_ms), "frac": frac_of(gbs(b_own * npairs_rank, step_ms), hbm_peak)},
                     "traffic": seed_traffic, "traffic_source": traffic_src,
                     "dram_gbs_measured": gbs(seed_traffic, seed_ms_alone) if seed_traffic else None,
                     "dram_frac_measured": frac_of(gbs(seed_traffic, seed_ms_alone), hbm_peak) if seed_traffic else None,
                     "positions_looked_up": lookups / max(1, positions), "block_probes_per_position": sum(st["seed_probes"] for _, _, st, _ in alone) / max(1, positions),
                     "survey_model": {"what": "SURVEY.md §8d counts a full binary search per position (12 B x log2 n + 8.25): work the kernel no longer does; "
                                              "equivalent-work rate only, NOT a roofline fraction", "bytes_per_launch": b_model, "equivalent_gbs": gbs(b_model, seed_ms_alone)},
                     "note": "random 32-byte sectors of a 64 MB table, a 20 MB suffix array and a 32 MB bitmap: a latency-bound gather, not a stream; "
                             "north_star's 50 % of HBM does not apply to an algorithm that avoids the traffic (DESIGN.md §4.3)"}
        # ---- extension wave 1
        w1_in, w1_alone = S_in("ms_wave1") / n_in, S("ms_wave1") / n_pairs_alone
        cells_pair = S("wave1_cells") / n_pairs_alone
        gops = lambda cells, ms: cells * 16 / (ms * 1e-3) / 1e9 if ms else None
        gcups_in = cells_pair / (w1_in * 1e-3) / 1e9 if w1_in else None
        gcups_alone = cells_pair / (w1_alone * 1e-3) / 1e9 if w1_alone else None
        roof_ext = {"kernel": "k_ex_wave1_tpj + k_ex_wave1_big (side by side on two streams)", "bound": "int32", "peak": int32_gops, "unit": "Gop/s",
                    "peak_source": "measured (pmn_measure_int32_peak, add+max chains)", "ops_per_cell": 16, "cells_per_launch": cells_pair,
                    "launch_ms": w1_alone, "gcups": gcups_alone, "achieved": gops(cells_pair, w1_alone), "frac": frac_of(gops(cells_pair, w1_alone), int32_gops),
                    "under_load": {"launch_ms": w1_in, "gcups": gcups_in, "achieved": gops(cells_pair, w1_in), "frac": frac_of(gops(cells_pair, w1_in), int32_gops)},
                    "whole_step": {"gcups": cells_pair * npairs_rank / (step_ms * 1e-3) / 1e9, "achieved": gops(cells_pair * npairs_rank, step_ms),
                                   "frac": frac_of(gops(cells_pair * npairs_rank, step_ms), int32_gops)},
                    "traffic": sum((traffic.get(k) or {}).get("dram_bytes_per_launch", 0) for k in ("k_ex_wave1_tpj", "k_ex_wave1_big")) or None,
                    "traffic_source": traffic_src,
                    "note": "the window is as long as its longest alignment (cluster-end searches one warp runs each, the largest windows of the thread-per-job kernel), "
                            "not INT32 issue: the thread-per-job kernel keeps the SMs busy for 15 % of it (DESIGN.md §4.5)"}
        b_index = sum(len(genomes[a][1]) * (4.25 + 16 * max(1, rounds) + 12.5 + 6) for a, rounds in {a: st["sa_rounds"] for a, _, st, _ in alone}.items())
        roof_idx = {"kernel": "index build (sort + doubling + lcp + table + skip table + presence map)", "bound": "hbm", "peak": hbm_peak, "unit": "GB/s",
                    "achieved": b_index / (stage_alone["index"] * 1e-3) / 1e9 if stage_alone["index"] else None, "launch_ms": stage_alone["index"] / max(1, len(seen))}
        roof_idx["frac"] = roof_idx["achieved"] / hbm_peak if roof_idx["achieved"] else None
        # the dominant kernel: the largest share of the time a pair spends on the device when it has the GPU (the ncu launch list of this
        # command, profiles/r02_launches_bench_summary.txt, gives the same order)
        tot_alone = stage_alone["seed"] + stage_alone["cluster"] + stage_alone["extend"]
        roof_seed["share_of_pair_time"] = S("ms_seed_kernel") / tot_alone if tot_alone else None
        roof_ext["share_of_pair_time"] = S("ms_wave1") / tot_alone if tot_alone else None
        dominant = dict(roof_ext if S("ms_wave1") >= S("ms_seed_kernel") else roof_seed)
        launches_pair = cnt_res["launches"] / max(1, npairs_rank * args.steps)
        out = {
            "metric": METRIC, "value": npairs_all * args.steps / (ms_res * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak" if mode == "weak" else "strong", "vs_baseline": None,
            "dtype": "int32 (2-bit packed text, u8 traceback)", "data": "synthetic",
            "config": {"workload": f"{args.genomes} synthetic {args.genome_bp / 1e6:g} Mbp genomes, all-vs-all {len(pairs)} pairs "
                                   f"(BASELINE.json configs[1]); rank 0: {npairs_rank} pairs, {len(refs)} index builds per step",
                       "mode": ("sharded by pair over the ranks, indexes " + ("built once and broadcast over NCCL" if broadcast else "rebuilt by every rank that needs them (no collective)"))
                               if sharded else ("every rank runs the whole batch" if world > 1 else "one GPU"),
                       "pairs_per_step_all_ranks": npairs_all, "workers_per_gpu": args.workers,
                       "hw_queues": int(os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "8")),
                       "l2": "no explicit flush: one step streams > 1 GB of index, staging and score data per pair through a 126 MB L2",
                       "parity": PARITY_NOTE},
            "parity_checked_pairs": int(checked_all),
            "parity_how": "sha256 of every pair's .delta == the oracle's committed digest (tests/golden/golden_configs.json), checked before the timed region on every rank",
            "aligned_mbp_per_s": aligned_all / 1e6 / (step_ms * 1e-3),
            "input_mbp_per_s": bp_all / 1e6 / (step_ms * 1e-3),
            "extension_gcups": gcups_alone, "extension_gcups_under_load": gcups_in,
            "e2e": {"value": npairs_all * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": cnt_e2e["h2d_bytes"] // args.steps, "d2h_bytes_per_step": cnt_e2e["d2h_bytes"] // args.steps,
                    "delta_bytes_per_step": sum(d[3] for d in alone)},
            "e2e_worker": {"what": "nucmer + delta-filter -1 + delta2maf per pair in the same call (pmn_opts.post = 1): .delta, filtered .delta and MAF in host memory",
                           "value": npairs_all * args.steps / (ms_wrk * 1e-3), "unit": UNIT, "ms_per_step": ms_wrk / args.steps,
                           "h2d_bytes_per_step": cnt_wrk["h2d_bytes"] // args.steps, "d2h_bytes_per_step": cnt_wrk["d2h_bytes"] // args.steps},
            "extra_weak": weak,
            "gpu_launches": cnt_res["launches"], "launches_per_pair": launches_pair,
            "host_syncs_per_pair": cnt_res["syncs"] / max(1, npairs_rank * args.steps),
            "step_wall_ms": {"resident": spread(cnt_res["walls"]), "e2e": spread(cnt_e2e["walls"]), "e2e_worker": spread(cnt_wrk["walls"]),
                             "every_step": {"resident": cnt_res["walls"], "e2e": cnt_e2e["walls"], "e2e_worker": cnt_wrk["walls"]}},
            "device_allocations_in_timed_region": {"resident": cnt_res["allocs"], "e2e": cnt_e2e["allocs"]},
            "ms_per_step_by_rank": {"resident": cnt_res["ranks_ms"], "e2e": cnt_e2e["ranks_ms"]}, "host_cores": os.cpu_count(),
            "clocks": clocks, "device_memory_used_gb": round(mem_used_gb, 1),
            "stage_ms_per_step": {"what": "sum over the pairs of a step of the CUDA-event time of each stage; under_load = inside the timed steps (16 pairs share the GPU), alone = one pair at a time",
                                  "under_load": stage_in, "alone": stage_alone},
            "roofline": dominant, "roofline_seed": roof_seed, "roofline_extend": roof_ext, "roofline_index": roof_idx,
            "counts_per_step": {"anchors": S("anchors"), "clusters": S("clusters"), "alignments": S("alignments"), "dp_cells": S("dp_cells"),
                                "dp_jobs": S("dp_jobs"), "aligned_ref_bases": aligned_bp, "seed_lookups": lookups, "arena_bytes_max": max(d[2]["arena_bytes"] for d in alone) if alone else 0},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, fastas, pairs, names, my_delta, budget_s=args.cpu_budget)
            out["parity_checked_pairs_live_oracle"] = out["cpu_baseline"].pop("parity_checked_pairs")
        emit(out)
    for s in resident.values():
        s.close()
    sched.close()
    R.close()


# ------------------------------------------------------------------------------------------ GPU arm, C4

def run_gpu_c4(args):
    """BASELINE.json configs[3]: one synthetic 100 Mbp pair, the query positions partitioned over the ranks.  Every rank holds
    both packed genomes and builds the index itself (all ranks at once: faster than one build plus a broadcast the others wait
    for), rank r seeds part r, the anchor lists are all-gathered, clustering and extension run on every rank (the .delta does
    not depend on N).  One step = index build + the pair."""
    from paramugsy_b200 import lib, multi
    R = Ranks(args)
    torch, rank, world, local = R.torch, R.rank, R.world, R.local
    n = args.genome_bp if args.genome_bp != 5_000_000 else 100_000_000
    gs = synth.config_c4(n=n, inv_len=max(1000, n // 100))
    ref_fa, qry_fa = synth.fasta(*gs[0]), synth.fasta(*gs[1])
    names = [gs[0][0] + ".fa", gs[1][0] + ".fa"]
    pin = [torch.frombuffer(bytearray(f), dtype=torch.uint8).pin_memory() for f in (ref_fa, qry_fa)]
    host = [(t.data_ptr(), t.numel()) for t in pin]
    ctx = lib.Context(local)
    rs, qs = ctx.sequence(ref_fa), ctx.sequence(qry_fa)
    dist = R.dist if world > 1 else None
    phases = []

    def step(rseq, qseq, keep=None):
        t0 = time.perf_counter()
        if args.replicate == "broadcast" and world > 1:
            ix = rseq.index() if rank == 0 else rseq.index(empty=True)
            dist.broadcast(ix.image_tensor(), src=0); torch.cuda.synchronize()
            if rank != 0:
                ix.adopt()
        else:
            ix = rseq.index()
        t1 = time.perf_counter()
        res = multi.align_large_pair(ix, qseq, rank, world, dist, ref_path=names[0], qry_path=names[1])
        t2 = time.perf_counter()
        phases.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
        if keep is not None:
            keep.append((res.delta, res.stats))
        res.close(); ix.close()

    def step_e2e():
        a = lib.Sequence.from_address(ctx, *host[0]); b = lib.Sequence.from_address(ctx, *host[1])
        step(a, b)
        b.close(); a.close()

    first = []
    step(rs, qs, keep=first)
    g = golden("c4").get("c0.1-c1.1") if n == 100_000_000 else None
    if g and sha(first[0][0]) != g["delta_sha256"]:
        raise SystemExit(f"bench.py: PARITY FAILURE on rank {rank}: the 100 Mbp .delta differs from the oracle's digest")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); sampler.wait_first_sample()
    for _ in range(args.warmup):
        step(rs, qs)
    del phases[:]
    ms_res, cnt = R.timed(lambda: step(rs, qs), args.steps, ctx.counters, lib.alloc_count)
    ph = list(phases)
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e, cnt_e = R.timed(step_e2e, args.steps, ctx.counters, lib.alloc_count)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        st = first[0][1]
        step_ms = ms_res / args.steps
        hbm_peak, peak_src = peaks()
        out = {"metric": METRIC, "value": args.steps / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "int32 (2-bit packed text, u8 traceback)", "data": "synthetic",
               "config": {"workload": f"one synthetic {n / 1e6:g} Mbp pair (BASELINE.json configs[3]), query positions partitioned over {world} rank(s); "
                                      "index " + ("built on rank 0 and broadcast over NCCL" if args.replicate == "broadcast" and world > 1 else "built by every rank") +
                                      ", anchors all-gathered, clustering and extension on every rank", "parity": PARITY_NOTE},
               "parity_checked_pairs": 1 if g else 0,
               "input_mbp_per_s": 2 * n / 1e6 / (step_ms * 1e-3), "aligned_mbp_per_s": st["aligned_ref_bases"] / 1e6 / (step_ms * 1e-3),
               "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                       "h2d_bytes_per_step": cnt_e["h2d_bytes"] // args.steps, "d2h_bytes_per_step": cnt_e["d2h_bytes"] // args.steps},
               "gpu_launches": cnt["launches"], "phase_ms_rank0": {"index": statistics.mean(p[0] for p in ph), "seed_gather_cluster_extend_text": statistics.mean(p[1] for p in ph)},
               "pair_stats": {k: st[k] for k in ("anchors", "clusters", "alignments", "aligned_ref_bases", "dp_cells", "seed_lookups", "arena_bytes", "sa_rounds", "kmer_bits")},
               "ms_per_step_by_rank": cnt["ranks_ms"], "clocks": clocks, "host_cores": os.cpu_count(),
               "roofline": {"kernel": "index build (radix-sort passes over 100 M suffixes)", "bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                            "achieved": n * (4.25 + 16 * max(1, st["sa_rounds"]) + 12.5 + 6) / (statistics.mean(p[0] for p in ph) * 1e-3) / 1e9, "traffic": None}}
        out["roofline"]["frac"] = out["roofline"]["achieved"] / hbm_peak
        emit(out)
    qs.close(); rs.close(); ctx.close()
    R.close()


# ------------------------------------------------------------------------------------------ CPU arm

def _cpu_one(job):
    ref, qry, rname, qname = job
    from oracle import pmn_oracle
    t = time.time()
    r = pmn_oracle.Run(ref, qry, fast_chain=1)
    d = r.delta(rname, qname)
    rows = r.alignments()[0]
    aligned = int((rows[:, 4] - rows[:, 3] + 1).sum()) if len(rows) else 0
    cells = r.dp_cells()
    r.close()
    return time.time() - t, d, aligned, cells


def cpu_sample(fastas, pairs, names, nproc, sample_bp):
    """`nproc` pairs of the workload, each cut to its first sample_bp bases (0 = as they are), one process per pair."""
    import multiprocessing as mp
    jobs = []
    for (i, j) in pairs[:nproc]:
        def cut(name, fa):
            if not sample_bp:
                return fa
            seq = b"".join(fa.split(b"\n")[1:])[:sample_bp]
            return synth.fasta(name, seq)
        jobs.append((cut(*fastas[i]), cut(*fastas[j]), names[i], names[j]))
    t = time.time()
    with mp.get_context("fork").Pool(nproc) as pool:
        res = pool.map(_cpu_one, jobs)
    wall = time.time() - t
    return wall, res, len(jobs)


def cpu_baseline(args, fastas, pairs, names, gpu_delta, budget_s=30.0):
    """The oracle on FULL-SIZE pairs of the workload, one process per pair on the host cores; every .delta it produces is
    compared byte for byte with what the GPU produced for the same pair."""
    from oracle import pmn_oracle
    pmn_oracle.build()
    ncores = os.cpu_count() or 1
    nproc = max(1, min(ncores, len(pairs), 8))
    # the oracle needs ~4 s per Mbp of pair length on one core; a full-size pair is taken whenever it fits the budget
    full = args.genome_bp * 4.5e-6 <= budget_s
    sample_bp = 0 if full else int(max(100_000, budget_s / 4.5 * 1e6))
    wall, res, n = cpu_sample(fastas, pairs, names, nproc, sample_bp)
    checked, bad = 0, []
    if full:
        for (i, j), r in zip(pairs[:nproc], res):
            if (i, j) in gpu_delta:
                checked += 1
                if gpu_delta[(i, j)] != r[1]:
                    bad.append((i, j))
    if bad:
        raise SystemExit(f"bench.py: PARITY FAILURE: the GPU .delta of pairs {bad} differs from the oracle run in this process")
    scale = 1.0 if full else sample_bp / args.genome_bp
    return {"value": n / wall * scale, "unit": UNIT, "cores": nproc, "kind": "port",
            "sample": (f"{n} full-size pairs of the workload ({args.genome_bp} bp each)" if full else f"{n} pairs cut to {sample_bp} bp, scaled linearly to {args.genome_bp} bp") +
                      f", {nproc} processes, one pair per process, wall {wall:.1f} s (oracle with fast_chain=1, identical output)",
            "aligned_mbp_per_s": sum(r[2] for r in res) / 1e6 / wall, "what": "CPU restatement oracle/pmn_oracle.c, not MUMmer",
            "parity_checked_pairs": checked}


def run_reference(args):
    """The reference's own CPU implementation of the path is MUMmer 3.20, which the reference does not vendor: the oracle port is
    timed instead, all host cores, one process per pair (`paramugsy local -cores N`: lib/base/paramugsy.ml:54-57).  A step is
    `cores` pairs of the workload; they are full size (no scaling, same config) whenever K + W steps of ~22 s fit four
    minutes, and cut to a stated prefix and scaled linearly otherwise."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pmn_oracle
    pmn_oracle.build()
    genomes, fastas, pairs = workload(args)
    names = [g[0] + ".fa" for g in genomes]
    ncores = os.cpu_count() or 1
    nproc = max(1, min(ncores, len(pairs)))
    sec_per_bp = 4.5e-6
    budget = args.ref_budget
    full_s = args.genome_bp * sec_per_bp * (args.steps + 0.1 * args.warmup)
    if args.ref_sample_bp:
        sample_bp = int(min(args.genome_bp, args.ref_sample_bp))
    elif full_s <= budget:
        sample_bp = 0
    else:
        sample_bp = int(max(200_000, args.genome_bp * budget / full_s))
    for _ in range(args.warmup):
        cpu_sample(fastas, pairs, names, nproc, min(sample_bp or args.genome_bp, 100_000))
    t0 = time.time(); n = 0
    for _ in range(args.steps):
        wall, res, k = cpu_sample(fastas, pairs, names, nproc, sample_bp)
        n += k
    tot = time.time() - t0
    scale = 1.0 if not sample_bp else sample_bp / args.genome_bp
    v = n / tot * scale
    sample = (f"each step: {nproc} full-size pairs ({args.genome_bp} bp), one process per pair, no scaling" if not sample_bp else
              f"each step: {nproc} pairs cut to their first {sample_bp} bp, one process per pair; scaled linearly to {args.genome_bp} bp pairs "
              f"(full-size steps would take {full_s:.0f} s for this K; the full-size rate is the cpu_baseline of the B200 arm's line)")
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64 (CPU)",
           "data": "synthetic",
           "config": {"workload": f"{args.genomes} synthetic {args.genome_bp / 1e6:g} Mbp genomes, all-vs-all {len(pairs)} pairs (BASELINE.json configs[1])",
                      "same_config": not sample_bp},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": nproc, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "the reference's own CPU path is MUMmer 3.20 (not vendored, absent here); this times the in-repo CPU restatement"}
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c4"])
    ap.add_argument("--mode", default="auto", choices=["auto", "weak", "strong"], help="N > 1: strong = the 28 pairs sharded over the ranks (default), weak = every rank runs the whole batch")
    ap.add_argument("--replicate", default="rebuild", choices=["rebuild", "broadcast"], help="sharded runs: every rank builds the indexes it needs (default) or one build per index and an NCCL broadcast")
    ap.add_argument("--workers", type=int, default=0, help="pairs in flight per GPU (pmn_sched worker threads); 0 = twice the host cores per local rank, between 8 and 32")
    ap.add_argument("--genomes", type=int, default=8)
    ap.add_argument("--genome-bp", type=int, default=5_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=30.0)
    ap.add_argument("--ref-sample-bp", type=int, default=0, help="reference arm: cut every pair to this many bases (0 = decide from the step count)")
    ap.add_argument("--ref-budget", type=float, default=240.0, help="reference arm: seconds the timed steps may take")
    args = ap.parse_args()
    args.workers_auto = args.workers <= 0
    if args.workers <= 0:
        local_ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
        args.workers = max(8, min(PMN_DEFAULT_WORKERS, 2 * (os.cpu_count() or 16) // local_ranks))
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c4":
        run_gpu_c4(args)
    else:
        run_gpu(args)


PMN_DEFAULT_WORKERS = 32      # pairs in flight per GPU when the host has the cores for them (16 -> 1955, 28 -> 2030, 32 -> 2040 pairs/s)


if __name__ == "__main__":
    main()
