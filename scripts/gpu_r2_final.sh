#!/bin/bash
# round 2: final check of the tree as committed -- -m gpu suite, smoke, default bench + reference arm, whole-function drop-in
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
echo "bench rc=$?"; head -c 400 gpurun_out/bench_1gpu.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
echo "reference arm rc=$?"; head -c 300 gpurun_out/bench_reference_arm.json; echo
timeout 900 python scripts/bench_dropin.py --reps 20 > gpurun_out/dropin_whole_function.json 2> gpurun_out/dropin.err
echo "dropin rc=$?"; head -c 600 gpurun_out/dropin_whole_function.json; echo
