#!/bin/bash
# round 2, first GPU visit: parity of the compact kernel, timing vs the wide-row kernel, one ncu capture
set -u
mkdir -p gpurun_out
timeout 1500 python scripts/gpu_compact_check.py --sites ${SITES:-1000000} > gpurun_out/compact_check.log 2>&1
echo "check rc=$?"
tail -5 gpurun_out/compact_check.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgt_compact_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_compact \
  python scripts/gpu_compact_check.py --sites 1000000 --skip-parity --steps 2 > gpurun_out/prof_r02_compact.out 2>&1
echo "ncu rc=$?"
ls -la gpurun_out | tail -5
