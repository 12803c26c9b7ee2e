#!/bin/bash
# One GPU visit (run under gpurun): tools/gpu_visit.sh <tag> [steps...]
#   steps: tests smoke bench bench_ref ncu_launches ncu_pair ncu_full sweep multi multic trace c4 configs stress stitchprof diag  (default: tests smoke bench)
# Everything lands in gpurun_out/<tag>_*.  Each step has its own timeout, a failing step does not stop the others.
tag=${1:-visit}; shift
steps=${@:-tests smoke bench}
mkdir -p gpurun_out
for s in $steps; do
  case $s in
    tests)   timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -5 gpurun_out/${tag}_pytest.log ;;
    tests_all) timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -15 gpurun_out/${tag}_pytest.log ;;
    smoke)   timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log ;;
    bench)   timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/${tag}_bench.json; echo ;;
    bench_ref) timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/${tag}_bench_ref.json; echo ;;
    ncu_launches) timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "ncu launches rc=$?" ;;
    ncu_pair) timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/launches_pair.csv python tools/profile_pair.py 5000000 2 > gpurun_out/${tag}_ncu_pair.log 2>&1; echo "ncu pair rc=$?"; cp gpurun_out/launches_pair.csv gpurun_out/${tag}_launches_pair.csv ;;
    ncu_full) timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:k_ex_wave1|k_ex_targets|k_ex_jobdesc|k_ex_stitch|k_seed$|k_cl_chains|pmn_rs_scatter|k_bucket_fill|k_skip_fill|k_present_fill" -c 24 -o gpurun_out/pair_full -f python tools/profile_pair.py 5000000 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?" ;;
    sweep)   # tools/sweep.txt: one run per line, "<name> [VAR=value ...] -- <bench.py arguments>"
             while read -r name rest; do
               [ -z "$name" ] && continue; case $name in \#*) continue;; esac
               envs=${rest%%--*}; args=${rest#*--}
               env $envs timeout 600 python bench.py --no-cpu-baseline $args > gpurun_out/${tag}_sweep_${name}.json 2> gpurun_out/${tag}_sweep_${name}.err
               python - "$name" gpurun_out/${tag}_sweep_${name}.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:24s} value {d['value']:8.1f}  e2e {d['e2e']['value']:8.1f}  worker {(d.get('e2e_worker') or {}).get('value', 0):8.1f}  ms/step {d['ms_per_step']:.2f}  W {d['config']['workers_per_gpu']}")
except Exception as e:
    print(f"{sys.argv[1]:24s} FAILED {e}")
P
             done < tools/sweep.txt ;;
    multi)   # GPUS=N tools/gpu_visit.sh <tag> multi : the sharded C2 batch (rebuild and broadcast) and the C4 pair on N ranks
             N=${GPUS:-2}; run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
             timeout 600 $run bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_c2_${N}gpu.json 2> gpurun_out/${tag}_c2_${N}gpu.err; echo "c2 x$N rc=$?"
             if [ -z "$NO_BROADCAST" ]; then timeout 600 $run bench.py --gpus $N --steps 20 --warmup 5 --replicate broadcast > gpurun_out/${tag}_c2_${N}gpu_broadcast.json 2> gpurun_out/${tag}_c2_${N}gpu_broadcast.err; echo "c2 broadcast x$N rc=$?"; fi
             timeout 900 $run bench.py --gpus $N --config c4 --steps 6 --warmup 2 > gpurun_out/${tag}_c4_${N}gpu.json 2> gpurun_out/${tag}_c4_${N}gpu.err; echo "c4 x$N rc=$?"
             if [ -z "$NO_BROADCAST" ]; then timeout 900 $run bench.py --gpus $N --config c4 --steps 6 --warmup 2 --replicate broadcast > gpurun_out/${tag}_c4_${N}gpu_broadcast.json 2> gpurun_out/${tag}_c4_${N}gpu_broadcast.err; echo "c4 broadcast x$N rc=$?"; fi
             python - gpurun_out/${tag}_c2_${N}gpu.json gpurun_out/${tag}_c2_${N}gpu_broadcast.json gpurun_out/${tag}_c4_${N}gpu.json gpurun_out/${tag}_c4_${N}gpu_broadcast.json <<'P'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        w = d.get("extra_weak") or {}
        print(f"{f}: value {d['value']:.1f} e2e {d['e2e']['value']:.1f} ms/step {d['ms_per_step']:.2f} weak {w.get('value', 0):.1f} by rank {d.get('ms_per_step_by_rank')}")
    except Exception as e:
        print(f, "FAILED", e)
P
             ;;
    multic)  # GPUS=N: pmn_multi (one process, N real devices through the C ABI): C2 digests + timed steps, a 20 Mbp pair cut over the devices
             N=${GPUS:-2}; timeout 600 python tools/multi_c_check.py $N 10 20000000 > gpurun_out/${tag}_multi_c_${N}gpu.json 2> gpurun_out/${tag}_multi_c_${N}gpu.err; echo "multi_c x$N rc=$?"; tail -1 gpurun_out/${tag}_multi_c_${N}gpu.json | cut -c1-900 ;;
    trace)   timeout 300 python tools/trace_pair.py > gpurun_out/${tag}_trace_pair.txt 2>&1; head -3 gpurun_out/${tag}_trace_pair.txt | tail -1 | cut -c1-300
             timeout 300 python tools/trace_pair.py 5000000 index > gpurun_out/${tag}_trace_index.txt 2>&1
             timeout 300 python tools/trace_step.py 32 > gpurun_out/${tag}_trace_w32.log 2>&1; sed -n 3,4p gpurun_out/${tag}_trace_w32.log ;;
    c4)      timeout 900 python bench.py --config c4 --steps 6 --warmup 2 > gpurun_out/${tag}_c4_1gpu.json 2> gpurun_out/${tag}_c4_1gpu.err; echo "c4 rc=$?"; head -c 600 gpurun_out/${tag}_c4_1gpu.json; echo ;;
    configs) timeout 600 python tools/run_configs.py c5 > gpurun_out/${tag}_c5.json 2> gpurun_out/${tag}_c5.err; echo "c5 rc=$?"
             timeout 900 python tools/run_configs.py c3 > gpurun_out/${tag}_c3.json 2> gpurun_out/${tag}_c3.err; echo "c3 rc=$?"; tail -c 400 gpurun_out/${tag}_c3.json; echo ;;
    stress)  timeout 600 python tools/stress.py 5000000 8 24 12 > gpurun_out/${tag}_stress.log 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/${tag}_stress.log
             timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/stress.py 300000 5 6 1 > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/${tag}_memcheck.log; tail -4 gpurun_out/${tag}_memcheck.log ;;
    stitchprof) PMN_STITCH_TIMING=1 timeout 300 python tools/profile_div.py 0.10 > gpurun_out/${tag}_stitch_q10.log 2>&1; echo "rc=$?"; grep "stitch cycles" gpurun_out/${tag}_stitch_q10.log | tail -1
             PMN_STITCH_TIMING=1 timeout 300 python tools/profile_div.py 0.02 > gpurun_out/${tag}_stitch_q02.log 2>&1; grep "stitch cycles" gpurun_out/${tag}_stitch_q02.log | tail -1 ;;
    diag)    # where a single pair and a small share spend their time: stitch sections (instrumented build in _lib_timing), engine-call log, step timeline, shares of an N-GPU plan
             PMN_LIB_DIR=_lib_timing PMN_STITCH_TIMING=1 timeout 300 python tools/probe_pairs.py c2 3 > gpurun_out/${tag}_stitch_c2.log 2>&1; grep "stitch cycles" gpurun_out/${tag}_stitch_c2.log | tail -1
             PMN_JOBLOG_LAST=gpurun_out/${tag}_joblog_c2.txt timeout 300 python tools/probe_pairs.py c2 3 > gpurun_out/${tag}_probe_c2.log 2>&1; tail -1 gpurun_out/${tag}_probe_c2.log | cut -c1-900
             python tools/joblog_summary.py gpurun_out/${tag}_joblog_c2.txt > gpurun_out/${tag}_joblog_c2_summary.txt 2>&1; rm -f gpurun_out/${tag}_joblog_c2.txt
             timeout 300 python tools/trace_step.py 32 > gpurun_out/${tag}_trace_w32.log 2>&1; head -4 gpurun_out/${tag}_trace_w32.log
             timeout 600 python tools/share_latency.py 32 > gpurun_out/${tag}_shares.log 2>&1; cat gpurun_out/${tag}_shares.log ;;
    *) echo "unknown step $s" ;;
  esac
done
