#!/bin/bash
# round 2: small-batch regime -- kernel time of each shape with the unit ramp applied below N sites per resident warp
set -u
mkdir -p gpurun_out
C=/dev/shm/svgt_cache
for cs in del1m4lib:125000 del1m4lib:250000 del1m4lib:500000 stress1m:200000 mixed100k:100000 del10k:10000; do
  cfg=${cs%%:*}; n=${cs#*:}
  for r in 256 32 8 0; do
    SVGT_C_RAMP_PER_WARP=$r timeout 600 python scripts/gpu_compact_check.py --config $cfg --sites $n --cache $C --skip-parity --no-e2e --no-wide --steps 30 --tag ramp_${cfg}_${n}_$r > gpurun_out/ramp_${cfg}_${n}_$r.log 2>&1
    echo "$cfg $n ramp<$r/warp: $(tail -1 gpurun_out/ramp_${cfg}_${n}_$r.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms' % d['compact']['ms_avg'], d['rows_md5'][:8])" 2>&1)"
  done
done
