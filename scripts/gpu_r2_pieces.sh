#!/bin/bash
# round 2: the piece plan (long sites scored in pieces + ordered replay): parity tests, then kernel time of every
# config shape with and without a plan
set -u
mkdir -p gpurun_out
C=/dev/shm/svgt_cache
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${K:-piece or plan or fixture or synthetic or hazard or classic or idempotent or unsafe}" > gpurun_out/pytest_pieces.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_pieces.log
fi
for cs in ${SHAPES:-del10k:10000 mixed100k:100000 stress1m:125000 del1m4lib:1000000}; do
  cfg=${cs%%:*}; n=${cs#*:}
  timeout 900 python scripts/gpu_compact_check.py --config $cfg --sites $n --cache $C --skip-parity --no-e2e --no-wide --steps 30 --pieces ${PIECES:-off,auto} --tag pc_${cfg}_${n} > gpurun_out/pc_${cfg}_${n}.log 2>&1
  echo "$cfg $n rc=$?"; tail -1 gpurun_out/pc_${cfg}_${n}.log | cut -c1-900
done
