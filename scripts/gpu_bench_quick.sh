#!/bin/bash
# quick bench of selected variants: bash scripts/gpu_bench_quick.sh [sites] ; VARIANTS="2 3"
SITES=${1:-200000}
mkdir -p gpurun_out
for v in ${VARIANTS:-2}; do
  timeout 900 python bench.py --sites $SITES --steps 30 --warmup 3 --variant $v --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_v$v.json'))
    print('variant $v: value %.1fM/s'%(d['value']/1e6), 'kernel %.3f ms'%d['roofline']['kernel_ms_avg'], 'frac %.3f'%d['roofline']['frac'])
except Exception as e: print('variant $v ERR', e); print(open('gpurun_out/bench_v$v.err').read()[-800:])
PY
done
