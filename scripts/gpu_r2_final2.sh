#!/bin/bash
# round 2, final tree: -m gpu suite, smoke, default bench + reference arm, whole-function drop-in, kernel time of every
# config shape, ncu --set full of the tally kernel (benchmark shape) and of both call kernel variants, launch list,
# fp64 add latency
set -u
mkdir -p gpurun_out
C=/dev/shm/svgt_cache
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
echo "bench rc=$?"; head -c 400 gpurun_out/bench_1gpu.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
echo "reference arm rc=$?"; head -c 300 gpurun_out/bench_reference_arm.json; echo
timeout 900 python scripts/bench_dropin.py --reps 20 > gpurun_out/dropin_whole_function.json 2> gpurun_out/dropin.err
echo "dropin rc=$?"; head -c 400 gpurun_out/dropin_whole_function.json; echo
./scripts/ub/ub_fp64_latency > gpurun_out/ub_fp64_latency.txt 2>&1; head -6 gpurun_out/ub_fp64_latency.txt
for cs in del10k:10000 mixed100k:100000 stress1m:125000 stress1m:200000 del1m4lib:125000 del1m4lib:1000000; do
  cfg=${cs%%:*}; n=${cs#*:}
  timeout 900 python scripts/gpu_compact_check.py --config $cfg --sites $n --cache $C --skip-parity --no-e2e --no-wide --steps 30 --pieces off --tag cfg_${cfg}_$n > gpurun_out/cfg_${cfg}_$n.log 2>&1
  echo "$cfg $n rc=$?"; tail -1 gpurun_out/cfg_${cfg}_$n.log | cut -c1-330
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgt_compact_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02f_del1m4lib \
  python scripts/gpu_compact_check.py --config del1m4lib --sites 1000000 --cache $C --skip-parity --no-e2e --no-wide --steps 2 --pieces off --tag ncu_tally > gpurun_out/prof_r02f_del1m4lib.out 2>&1
echo "ncu tally rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgt_call_compact_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02f_call_1m \
  python scripts/gpu_compact_check.py --config del1m4lib --sites 1000000 --cache $C --skip-parity --no-e2e --no-wide --steps 2 --pieces off --tag ncu_call > gpurun_out/prof_r02f_call_1m.out 2>&1
echo "ncu call 1M rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgt_call_compact_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02f_call_stress125k \
  python scripts/gpu_compact_check.py --config stress1m --sites 125000 --cache $C --skip-parity --no-e2e --no-wide --steps 2 --pieces off --tag ncu_call_s > gpurun_out/prof_r02f_call_stress125k.out 2>&1
echo "ncu call stress rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_1m.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_bench.out 2>&1
echo "launch list rc=$?"
ls -la gpurun_out | tail -12
