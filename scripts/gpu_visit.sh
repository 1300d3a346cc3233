#!/bin/bash
# One GPU-box visit: GPU parity tests, quick benches of chosen variants, default bench, ncu full capture
# of the tally kernel (200k sites).  Usage (under gpurun): bash scripts/gpu_visit.sh <tag> ; VARIANTS="2 5"
TAG=${1:-cur}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
VARIANTS="${VARIANTS:-}" bash scripts/gpu_bench_quick.sh 200000
if [ -z "$NOBENCH" ]; then
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo rc=$?; cut -c1-400 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
fi
B="python bench.py --sites 200000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
echo "== ncu full (tally)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"svgt_(tally|lean)" -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_tally $B > gpurun_out/prof_${TAG}_tally.out 2>&1; echo rc=$?
ls -la gpurun_out | tail -8
