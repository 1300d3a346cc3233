#!/bin/bash
# bench every A/B build under svtyper_b200/ab/: bash scripts/gpu_ab.sh [sites] [variant]
SITES=${1:-200000}; V=${2:-2}
mkdir -p gpurun_out
for lib in svtyper_b200/ab/libsvgt_*.so; do
  name=$(basename $lib .so)
  SVGT_LIB=$PWD/$lib timeout 600 python bench.py --sites $SITES --steps 30 --warmup 3 --variant $V --no-cpu-baseline --e2e-steps 1 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_$name.json'))
    print('$name: value %.1fM/s'%(d['value']/1e6), 'kernel %.4f ms'%d['roofline']['kernel_ms_avg'], 'frac %.3f'%d['roofline']['frac'])
except Exception as e: print('$name ERR', e); print(open('gpurun_out/ab_$name.err').read()[-500:])
PY
done
