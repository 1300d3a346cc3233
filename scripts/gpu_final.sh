#!/bin/bash
# Round evidence: full default bench (1M sites, cpu baseline, parity), reference arm, ncu launch list +
# full capture of the tally kernel at the bench workload.  Usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
echo "== bench (default)"; timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo rc=$?; tail -c 2500 gpurun_out/bench_$TAG.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo rc=$?; cat gpurun_out/bench_ref_$TAG.json | cut -c1-600
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/launches_$TAG.out 2>&1; echo rc=$?
echo "== ncu full (tally)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"svgt_(tally|lean)" -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_tally $B > gpurun_out/prof_${TAG}_tally.out 2>&1; echo rc=$?
echo "== ncu full (call)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgt_call -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_call $B > gpurun_out/prof_${TAG}_call.out 2>&1; echo rc=$?
ls -la gpurun_out | tail -12
