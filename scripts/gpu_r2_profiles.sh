#!/bin/bash
# round 2 evidence: kernel timing of every config shape, ncu --set full of the tally kernel on each, the call kernel
# on the benchmark shape, and the ncu launch list of the default bench command
set -u
mkdir -p gpurun_out
C=/dev/shm/svgt_cache
for cs in del10k:10000 mixed100k:100000 stress1m:200000 del1m4lib:1000000; do
  cfg=${cs%%:*}; n=${cs#*:}
  timeout 900 python scripts/gpu_compact_check.py --config $cfg --sites $n --cache $C --skip-parity --no-e2e --no-wide --steps 30 --tag cfg_$cfg > gpurun_out/cfg_$cfg.log 2>&1
  echo "$cfg rc=$?"; tail -1 gpurun_out/cfg_$cfg.log | cut -c1-400
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgt_compact_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_$cfg \
    python scripts/gpu_compact_check.py --config $cfg --sites $n --cache $C --skip-parity --no-e2e --no-wide --steps 2 --tag ncu_$cfg > gpurun_out/prof_r02_$cfg.out 2>&1
  echo "ncu $cfg rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgt_call_compact_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_call \
  python scripts/gpu_compact_check.py --config del1m4lib --sites 1000000 --cache $C --skip-parity --no-e2e --no-wide --steps 2 --tag ncu_call > gpurun_out/prof_r02_call.out 2>&1
echo "ncu call rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_1m.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_bench.out 2>&1
echo "launch list rc=$?"
ls -la gpurun_out | tail -15
