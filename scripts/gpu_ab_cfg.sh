#!/bin/bash
# A/B builds under svtyper_b200/ab/ on a chosen config: bash scripts/gpu_ab_cfg.sh <config> <sites>
mkdir -p gpurun_out
for lib in svtyper_b200/ab/libsvgt_*.so; do
  name=$(basename $lib .so)
  SVGT_LIB=$PWD/$lib timeout 900 python bench.py --config $1 --sites $2 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/abc_$name.json 2> gpurun_out/abc_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/abc_$name.json'))
    print('$1 $name: %.1fM/s'%(d['value']/1e6), 'kernel %.4f ms'%d['roofline']['kernel_ms_avg'], 'frac %.3f'%d['roofline']['frac'])
except Exception as e: print('$name ERR', e); print(open('gpurun_out/abc_$name.err').read()[-400:])
PY
done
