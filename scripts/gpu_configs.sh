#!/bin/bash
# Throughput of the other BASELINE.json config shapes (parity-test cases, not bench lines), default kernel.
# Usage (under gpurun): bash scripts/gpu_configs.sh
mkdir -p gpurun_out
for cs in "del10k 10000" "mixed100k 100000" "stress1m 200000"; do
  set -- $cs
  timeout 900 python bench.py --config $1 --sites $2 --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/cfg_$1.json 2> gpurun_out/cfg_$1.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/cfg_$1.json'))
    print('$1 ($2 sites): %.1fM/s'%(d['value']/1e6), 'kernel %.4f ms'%d['roofline']['kernel_ms_avg'], 'frac %.3f'%d['roofline']['frac'], 'rows', d['config']['fragment_rows_per_gpu'], d['config']['split_rows_per_gpu'])
except Exception as e: print('$1 ERR', e); print(open('gpurun_out/cfg_$1.err').read()[-600:])
PY
done
