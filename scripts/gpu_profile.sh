#!/bin/bash
# ncu evidence for the scoring kernel: launch list + one full capture per variant.
# Usage (under gpurun): bash scripts/gpu_profile.sh [sites] [tag]
set -u
SITES=${1:-200000}
TAG=${2:-r01}
mkdir -p gpurun_out
B="python bench.py --sites $SITES --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/launches_$TAG.out 2>&1
echo "launch list rc=$?"
for v in ${VARIANTS:-0 1}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-svgt_} -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_v$v $B --variant $v > gpurun_out/prof_${TAG}_v$v.out 2>&1
  echo "full capture v$v rc=$?"
done
ls -la gpurun_out
