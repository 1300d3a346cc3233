#!/bin/bash
# round 2: per-kernel durations (ncu launch list) of a config shape scored with a piece plan
set -u
mkdir -p gpurun_out
C=/dev/shm/svgt_cache
for cs in ${SHAPES:-del10k:10000 stress1m:125000}; do
  cfg=${cs%%:*}; n=${cs#*:}
  timeout 600 python scripts/gpu_compact_check.py --config $cfg --sites $n --cache $C --skip-parity --no-e2e --no-wide --steps 20 --pieces ${PIECES:-off,auto} --tag pl_${cfg}_${n} > gpurun_out/pl_${cfg}_${n}.log 2>&1
  echo "$cfg $n rc=$?"; tail -1 gpurun_out/pl_${cfg}_${n}.log | cut -c1-1500
  for pc in ${NCU_PIECES:-auto}; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svgt -c 24 --csv --log-file gpurun_out/ll_${cfg}_${n}_$pc.csv \
    python scripts/gpu_compact_check.py --config $cfg --sites $n --cache $C --skip-parity --no-e2e --no-wide --steps 4 --pieces $pc --tag ncu_ll > /dev/null 2>&1
  echo "ncu $cfg $pc rc=$?"; tail -9 gpurun_out/ll_${cfg}_${n}_$pc.csv | cut -d, -f5,12- | cut -c1-200
  done
done
