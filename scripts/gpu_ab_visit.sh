#!/bin/bash
# GPU parity tests (optional), then A/B benches of every build under svtyper_b200/ab/ and of chosen variants.
# Usage (under gpurun): [NOTEST=1] [VARIANTS="3 4"] bash scripts/gpu_ab_visit.sh [sites] [variant]
mkdir -p gpurun_out
if [ -z "$NOTEST" ]; then
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_ab.log
fi
bash scripts/gpu_ab.sh ${1:-200000} ${2:-5}
if [ -n "$VARIANTS" ]; then bash scripts/gpu_bench_quick.sh ${1:-200000}; fi
