#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, short benches of both kernel variants, ncu launch list.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [sites]
set -u
SITES=${1:-200000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-0 1 2}; do
  echo "== bench variant $v"
  timeout 900 python bench.py --sites $SITES --steps 20 --warmup 3 --variant $v --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  echo "rc=$?"; tail -c 1500 gpurun_out/bench_v$v.json; tail -3 gpurun_out/bench_v$v.err
done
