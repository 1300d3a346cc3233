#!/bin/bash
# round 2: parity of the default build, then kernel timing of every A/B build on one cached 1M-site batch
set -u
mkdir -p gpurun_out
C=/dev/shm/svgt_cache
timeout 1500 python scripts/gpu_compact_check.py --sites ${SITES:-1000000} --cache $C --no-wide --tag base > gpurun_out/compact_check.log 2>&1
echo "check rc=$?"; tail -3 gpurun_out/compact_check.log
for lib in svtyper_b200/ab/libsvgt_*.so; do
  name=$(basename $lib .so)
  SVGT_LIB=$PWD/$lib timeout 600 python scripts/gpu_compact_check.py --sites ${SITES:-1000000} --cache $C --skip-parity --no-e2e --tag $name > gpurun_out/ab_$name.log 2>&1
  echo "$name rc=$?"; tail -1 gpurun_out/ab_$name.log
done
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgt_compact_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_compact \
  python scripts/gpu_compact_check.py --sites 1000000 --cache $C --skip-parity --no-e2e --steps 2 --tag ncu > gpurun_out/prof_r02_compact.out 2>&1
echo "ncu rc=$?"
fi
