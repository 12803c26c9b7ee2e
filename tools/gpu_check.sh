#!/bin/bash
# usage: tools/gpu_check.sh <stages> [case ...]   — each case in its own process with a timeout
stages=${1:-index,seed}; shift
cases=${@:-tiny 200k multi repeat C1}
mkdir -p gpurun_out
: > gpurun_out/check.log
rc=0
for c in $cases; do
  timeout 150 python -u tools/gpu_stage_check.py "$stages" "$c" >> gpurun_out/check.log 2>&1 || { echo "case $c rc=$?" >> gpurun_out/check.log; rc=1; }
done
tail -60 gpurun_out/check.log
exit $rc
