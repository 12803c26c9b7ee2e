#!/usr/bin/env python
"""bench.py — the pairwise nucmer stage on N B200s, one JSON line.

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...          # the CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): 8 synthetic 5 Mbp genomes (2 % divergence from a common
ancestor, two 50 kbp inversions each), all-vs-all = 28 (reference, query) pairs, the earlier
genome being the reference (lib/base/pm_job.ml:43-51).  One STEP = one pass of the hot path
over that batch: 7 index builds + 28 x (seeding, clustering, extension, .delta text).

  value   pairs/s with the packed genomes already resident in HBM (index build is inside).
  e2e     the same through the C ABI from FASTA bytes in HOST memory to .delta bytes in HOST
          memory: parse + H2D + pack of all 8 genomes and D2H of every result are inside.
  N > 1   weak scaling: every rank runs the same 28-pair batch on its own GPU (the path is
          embarrassingly parallel by pair; `--mode strong` shards the 28 pairs over the ranks
          with the reference indexes built once and broadcast over NCCL instead).
"""
import argparse
import json
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


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.device)],
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
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ GPU arm

def run_gpu(args):
    import torch
    import torch.distributed as dist
    from paramugsy_b200 import lib

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if getattr(args, "workers_auto", False):
        # a worker holds about 9 GB of scratch and traceback arenas: no more workers than the free memory takes
        free_b, _ = torch.cuda.mem_get_info()
        args.workers = max(4, min(args.workers, int((free_b / 2**30 - 20) // 9.5)))
    sched = lib.Scheduler(local, args.workers)
    ctx = sched.context(0)

    genomes, fastas, pairs = workload(args)
    names = [g[0] for g in genomes]
    # end-to-end arm: the FASTA text of every genome sits in PINNED host memory and is copied to the device every step
    pinned = [torch.frombuffer(bytearray(f[1]), dtype=torch.uint8).pin_memory() for f in fastas]
    fasta_bytes = [(t.data_ptr(), t.numel()) for t in pinned]
    strong = args.mode == "strong" and world > 1
    if strong:
        # the 28 pairs are dealt to the ranks; every reference index is built once in the whole job and
        # replicated to the ranks that need it by an NCCL broadcast of its image (paramugsy_b200/multi.py)
        from paramugsy_b200 import multi
        assignment = multi.assign_pairs(pairs, world)
        plan = multi.index_plan(pairs, assignment)
        my_pairs = [pairs[k] for k in assignment[rank]]
        needed = sorted({g for p in my_pairs for g in p} | {ref for ref, (o, rs) in plan.items() if o == rank or rank in rs})
    else:
        my_pairs = pairs
        needed = list(range(len(genomes)))
    refs = sorted({i for i, _ in my_pairs})
    total_bp_in = sum(len(genomes[i][1]) + len(genomes[j][1]) for i, j in my_pairs)

    def step_resident(seqs, collect=None, one_worker=None):
        """One pass of the hot path over the batch, genomes already packed in HBM: every reference
        index is built once, all pairs are aligned, every .delta ends up in host memory."""
        if one_worker is not None:          # the instrumented pass: one pair at a time, kernels timed alone
            for i in refs:
                ix = seqs[i].index()
                for (a, b) in my_pairs:
                    if a == i:
                        res = ix.align(seqs[b], ref_path=names[a], qry_path=names[b])
                        collect.append((a, b, res.stats, len(res.delta)))
                        # the two post-steps every pair goes through in the reference (mugsy_nucmer.ml:102-105,118-124)
                        t0 = time.perf_counter(); filt = ctx.delta_filter(res.delta, 1)
                        t1 = time.perf_counter(); maf = ctx.delta2maf(filt, seqs[a], seqs[b]); t2 = time.perf_counter()
                        post.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, len(filt), len(maf)))
                        res.close()
                ix.close()
            return
        if strong:
            out = multi.AllVsAll(sched, [seqs.get(g) for g in range(len(genomes))], names, pairs, rank, world, dist).step()
            for res in out.values():
                res.close()
            return
        for res in sched.align_seqs([seqs[g] for g in range(len(genomes))], my_pairs, names=names):
            res.close()

    def step_e2e():
        """The same from FASTA bytes in host memory (parse, H2D, pack inside)."""
        if strong:
            seqs = {g: ctx.sequence(fastas[g][1]) for g in needed}
            step_resident(seqs)
            for q in seqs.values():
                q.close()
            return
        for res in sched.align_fasta(fasta_bytes, my_pairs, names=names):
            res.close()

    def step_worker():
        """The whole reference worker per pair (lib/nucmer/mugsy_nucmer.ml:127-131): nucmer, delta-filter -1 and delta2maf,
        from FASTA text in pinned host memory to the three texts in host memory, in one scheduler call (pmn_opts.post)."""
        for res in sched.align_fasta(fasta_bytes, my_pairs, names=names, post=1):
            res.close()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        if os.environ.get("PMN_ALLOC_LOG"):
            print(f"[bench] timed region starts, {lib.alloc_count()} allocations so far", file=sys.stderr, flush=True)
        # the workers launch on their own streams and every call returns with all of them drained, so
        # events recorded on the (idle) current stream around the calls bracket exactly the device work
        barrier()
        c0 = sched.counters(); a0 = lib.alloc_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        walls = []
        for _ in range(steps):
            t = time.perf_counter(); fn(); walls.append((time.perf_counter() - t) * 1e3)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        ranks_ms = [round(ms / steps, 2)]
        if world > 1:
            allms = torch.zeros(world, device="cuda", dtype=torch.float64); allms[rank] = ms
            dist.all_reduce(allms)                              # every rank's own time (diagnostic: stragglers)
            ranks_ms = [round(float(x) / steps, 2) for x in allms.tolist()]
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        c1 = sched.counters()
        d = {k: c1[k] - c0[k] for k in c0}; d["walls"] = [round(w, 2) for w in walls]; d["allocs"] = lib.alloc_count() - a0
        d["ranks_ms"] = ranks_ms
        return ms, d

    # ---- resident arm
    resident = {g: ctx.sequence(fastas[g][1]) for g in needed}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # before the warm-up: nvidia-smi's own start-up must not land in the timed region
        sampler.wait_first_sample()
    barrier()
    for _ in range(args.warmup):
        step_resident(resident)
    ms_res, cnt_res = timed(lambda: step_resident(resident), args.steps)
    # ---- end-to-end arm (host FASTA bytes -> host .delta bytes)
    for _ in range(args.warmup):
        step_e2e()
    ms_e2e, cnt_e2e = timed(step_e2e, args.steps)
    ms_wrk, cnt_wrk = (None, None)
    if not strong:
        for _ in range(args.warmup):
            step_worker()
        ms_wrk, cnt_wrk = timed(step_worker, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # ---- one instrumented pass for the per-kernel figures
    detail = []; post = []
    step_resident(resident, detail, one_worker=True)
    int32_gops, _ = ctx.int32_peak()

    npairs_rank = len(my_pairs)
    if world > 1:
        t = torch.tensor([npairs_rank, total_bp_in], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        npairs_all, bp_all = float(t[0].item()), float(t[1].item())
    else:
        npairs_all, bp_all = float(npairs_rank), float(total_bp_in)

    if rank == 0:
        hbm_peak, peak_src = peaks()
        S = lambda k: sum(d[2][k] for d in detail)
        aligned_bp = S("aligned_ref_bases")
        stage_ms = {"index": 0.0, "seed": S("ms_seed"), "cluster": S("ms_cluster"), "extend": S("ms_extend")}
        # every reference index is built once per step: take ms_index once per distinct reference
        seen = {}
        for a, b, st, _ in detail:
            seen.setdefault(a, st["ms_index"])
        stage_ms["index"] = sum(seen.values())
        step_ms = ms_res / args.steps
        # algorithmic bytes (SURVEY.md §8d / DESIGN.md §6)
        import math
        b_seed = sum(2 * st["qry_bases"] * (12 * math.ceil(math.log2(max(2, st["ref_bases"]))) + 8.25) + 12 * st["anchors"] for _, _, st, _ in detail)
        b_index = sum(len(genomes[a][1]) * (4.25 + 16 * max(1, rounds) + 12.5)
                      for a, rounds in {a: st["sa_rounds"] for a, _, st, _ in detail}.items())
        seed_kernel_ms = S("ms_seed_kernel")
        wave1_ms, wave1_cells = S("ms_wave1"), S("wave1_cells")
        roof_seed = {"kernel": "k_seed", "bound": "hbm", "achieved": b_seed / (seed_kernel_ms * 1e-3) / 1e9 if seed_kernel_ms else None,
                     "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
                     "algorithmic_bytes_per_launch": b_seed / max(1, len(detail)), "launch_ms": seed_kernel_ms / max(1, len(detail)),
                     "share_of_step": seed_kernel_ms / (sum(stage_ms.values()) or 1)}
        roof_seed["frac"] = roof_seed["achieved"] / hbm_peak if roof_seed["achieved"] else None
        # SURVEY §8d counts a full binary search per position (12 B x log2 n); the K-mer table replaces it, so the kernel
        # moves far fewer bytes than that model (hence frac > 1 at 5 Mbp, where the index is also L2-resident).  The bytes the
        # table path itself needs per (position, strand): table 8 + suffix 4 + reference word 8 + query 0.375.
        b_seed_table = sum(2 * st["qry_bases"] * 20.375 + 16 * st["anchors"] for _, _, st, _ in detail)
        roof_seed["table_path"] = {"bytes_per_launch": b_seed_table / max(1, len(detail)),
                                   "achieved": b_seed_table / (seed_kernel_ms * 1e-3) / 1e9 if seed_kernel_ms else None,
                                   "frac": b_seed_table / (seed_kernel_ms * 1e-3) / 1e9 / hbm_peak if seed_kernel_ms else None,
                                   "what": "bytes the K-mer-table path needs (20.375 B per position and strand), same launch time"}
        gcups = wave1_cells / (wave1_ms * 1e-3) / 1e9 if wave1_ms else None
        roof_ext = {"kernel": "k_ex_wave1_tpj + k_ex_wave1_big (side by side on two streams)", "bound": "int32", "traffic": None, "achieved": gcups * 16 if gcups else None, "peak": int32_gops, "unit": "Gop/s",
                    "peak_source": "measured (pmn_measure_int32_peak, add+max chains)", "gcups": gcups, "ops_per_cell": 16,
                    "cells_per_step": wave1_cells, "launch_ms": wave1_ms / max(1, len(detail)),
                    "share_of_step": wave1_ms / (sum(stage_ms.values()) or 1)}
        roof_ext["frac"] = roof_ext["achieved"] / int32_gops if roof_ext["achieved"] else None
        roof_idx = {"kernel": "index build (sort + doubling + lcp + table)", "bound": "hbm", "achieved": b_index / (stage_ms["index"] * 1e-3) / 1e9 if stage_ms["index"] else None,
                    "peak": hbm_peak, "unit": "GB/s"}
        roof_idx["frac"] = roof_idx["achieved"] / hbm_peak if roof_idx["achieved"] else None
        # DRAM bytes per launch from the committed ncu --set full capture of one 5 Mbp pair (profiles/): static, not measured in this run
        tr = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
        if os.path.exists(tr) and args.genome_bp == 5_000_000:
            kt = json.load(open(tr))["kernels"]
            if "k_seed" in kt: roof_seed["traffic"] = kt["k_seed"]["dram_bytes_per_launch"]
            roof_ext["traffic"] = sum(kt[k]["dram_bytes_per_launch"] for k in ("k_ex_wave1_tpj", "k_ex_wave1_big") if k in kt) or None
            roof_ext["traffic_source"] = roof_seed["traffic_source"] = "profiles/r01_ncu_traffic.json (ncu --set full, one pair)"
        dominant = roof_ext if wave1_ms >= seed_kernel_ms else roof_seed
        out = {
            "metric": METRIC, "value": npairs_all * args.steps / (ms_res * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak" if args.mode == "weak" else "strong", "vs_baseline": None,
            "dtype": "int32 (2-bit packed text, u8 traceback)", "data": "synthetic",
            "config": {"workload": f"{args.genomes} synthetic {args.genome_bp / 1e6:g} Mbp genomes, all-vs-all {len(pairs)} pairs "
                                   f"(BASELINE.json configs[1]); per rank: {npairs_rank} pairs, {len(refs)} index builds per step",
                       "mode": args.mode, "pairs_per_step_all_ranks": npairs_all, "workers_per_gpu": args.workers,
                       "hw_queues": int(os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "8")),
                       "l2": "no explicit flush: one step streams > 1 GB of index, staging and score data per pair through a 126 MB L2"},
            "aligned_mbp_per_s": aligned_bp * world / 1e6 / (step_ms * 1e-3) if args.mode == "weak" else aligned_bp / 1e6 / (step_ms * 1e-3),
            "input_mbp_per_s": bp_all / 1e6 / (step_ms * 1e-3),
            "extension_gcups": gcups,
            "e2e": {"value": npairs_all * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": cnt_e2e["h2d_bytes"] // args.steps, "d2h_bytes_per_step": cnt_e2e["d2h_bytes"] // args.steps,
                    "delta_bytes_per_step": sum(d[3] for d in detail)},
            "e2e_worker": {"what": "nucmer + delta-filter -1 + delta2maf per pair in the same call (pmn_opts.post = 1): .delta, filtered .delta and MAF in host memory",
                           "value": npairs_all * args.steps / (ms_wrk * 1e-3), "unit": UNIT, "ms_per_step": ms_wrk / args.steps,
                           "h2d_bytes_per_step": cnt_wrk["h2d_bytes"] // args.steps, "d2h_bytes_per_step": cnt_wrk["d2h_bytes"] // args.steps} if ms_wrk else None,
            "gpu_launches": cnt_res["launches"], "step_wall_ms": {"resident": cnt_res["walls"], "e2e": cnt_e2e["walls"], "e2e_worker": cnt_wrk["walls"] if cnt_wrk else None},
            "device_allocations_in_timed_region": {"resident": cnt_res["allocs"], "e2e": cnt_e2e["allocs"]},
            "ms_per_step_by_rank": {"resident": cnt_res["ranks_ms"], "e2e": cnt_e2e["ranks_ms"]}, "host_cores": os.cpu_count(),
            "clocks": clocks,
            "stage_ms_per_step": stage_ms,
            "host_wall_ms_per_step": {"index_build": sum({a: st["wall_ms_index"] for a, _, st, _ in detail}.values()),
                                      "align_calls": S("wall_ms_align"), "of_which_delta_text": S("wall_ms_text")},
            "post_steps_per_pair": {"what": "delta-filter -1 then delta2maf on each pair's .delta (pmn_delta_filter, pmn_delta2maf: host text in, host text out), one pair at a time, wall clock",
                                    "delta_filter_ms": statistics.mean(p[0] for p in post), "delta2maf_ms": statistics.mean(p[1] for p in post),
                                    "maf_bytes": statistics.mean(p[3] for p in post), "maf_gb_per_s": sum(p[3] for p in post) / 1e9 / (sum(p[1] for p in post) * 1e-3)} if post else None,
            "roofline": dominant, "roofline_seed": roof_seed, "roofline_extend": roof_ext, "roofline_index": roof_idx,
            "counts_per_step": {"anchors": S("anchors"), "clusters": S("clusters"), "alignments": S("alignments"), "dp_cells": S("dp_cells"),
                                "dp_jobs": S("dp_jobs"), "aligned_ref_bases": aligned_bp},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, genomes, fastas, pairs, budget_s=args.cpu_budget)
        emit(out)
    for s in resident.values():
        s.close()
    sched.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ CPU arm

def _cpu_one(job):
    ref, qry = job
    from oracle import pmn_oracle
    t = time.time()
    r = pmn_oracle.Run(ref, qry, fast_chain=1)
    d = r.delta("ref", "qry")
    rows = r.alignments()[0]
    aligned = int((rows[:, 4] - rows[:, 3] + 1).sum()) if len(rows) else 0
    cells = r.dp_cells()
    r.close()
    return time.time() - t, len(d), aligned, cells


def cpu_sample(args, fastas, pairs, nproc, sample_bp):
    """`nproc` pairs of the workload, each truncated to sample_bp bases, one process per pair."""
    import multiprocessing as mp
    jobs = []
    for (i, j) in pairs[:nproc]:
        def cut(name, fa):
            seq = b"".join(fa.split(b"\n")[1:])[:sample_bp]
            return synth.fasta(name, seq)
        jobs.append((cut(*fastas[i]), cut(*fastas[j])))
    t = time.time()
    with mp.get_context("fork").Pool(nproc) as pool:
        res = pool.map(_cpu_one, jobs)
    wall = time.time() - t
    return wall, res, len(jobs)


def cpu_baseline(args, genomes, fastas, pairs, budget_s=25.0):
    from oracle import pmn_oracle
    pmn_oracle.build()
    ncores = os.cpu_count() or 1
    nproc = max(1, min(ncores, len(pairs), 8))
    # the oracle needs ~3.5 s per Mbp of pair length on one core; bound the sample to the budget
    sample_bp = int(min(args.genome_bp, max(100_000, budget_s / 4.0 * 1e6)))
    wall, res, n = cpu_sample(args, fastas, pairs, nproc, sample_bp)
    # pairs/s at FULL pair size, scaled linearly in sequence length from the sample
    scale = sample_bp / args.genome_bp
    return {"value": n / wall * scale, "unit": UNIT, "cores": nproc, "kind": "port",
            "sample": f"{n} pairs of the workload truncated to {sample_bp} bp each, {nproc} processes, one pair per process, "
                      f"wall {wall:.1f} s; scaled linearly to {args.genome_bp} bp pairs (oracle with fast_chain=1, identical output)",
            "aligned_mbp_per_s": sum(r[2] for r in res) / 1e6 / wall, "what": "CPU restatement oracle/pmn_oracle.c, not MUMmer"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pmn_oracle
    pmn_oracle.build()
    genomes, fastas, pairs = workload(args)
    ncores = os.cpu_count() or 1
    nproc = max(1, min(ncores, len(pairs)))
    sample_bp = int(min(args.genome_bp, args.ref_sample_bp))
    for _ in range(args.warmup):
        cpu_sample(args, fastas, pairs, nproc, min(sample_bp, 100_000))
    t0 = time.time(); n = 0
    for _ in range(args.steps):
        wall, res, k = cpu_sample(args, fastas, pairs, nproc, sample_bp)
        n += k
    tot = time.time() - t0
    scale = sample_bp / args.genome_bp
    v = n / tot * scale
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64 (CPU)",
           "data": "synthetic",
           "config": {"workload": f"{args.genomes} synthetic {args.genome_bp / 1e6:g} Mbp genomes, all-vs-all {len(pairs)} pairs (BASELINE.json configs[1])"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": nproc, "kind": "port",
                            "sample": f"each step: {nproc} pairs truncated to {sample_bp} bp, one process per pair; scaled linearly to full-size pairs"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "the reference's own CPU path is MUMmer 3.20 (not vendored, absent here); this times the in-repo CPU restatement"}
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="weak", choices=["weak", "strong"])
    ap.add_argument("--workers", type=int, default=0, help="pairs in flight per GPU (pmn_sched worker threads); 0 = twice the host cores per local rank, between 8 and 16")
    ap.add_argument("--genomes", type=int, default=8)
    ap.add_argument("--genome-bp", type=int, default=5_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--ref-sample-bp", type=int, default=1_000_000)
    args = ap.parse_args()
    args.workers_auto = args.workers <= 0
    if args.workers <= 0:
        # 16 pairs in flight fill one B200 (20 no longer fit its memory); a host with few cores per GPU (8 ranks on 32 cores)
        # is better off with 8 worker threads per rank
        local_ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
        args.workers = max(8, min(16, 2 * (os.cpu_count() or 16) // local_ranks))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
