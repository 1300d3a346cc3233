#!/bin/bash
# ncu full capture of the tally kernel for given A/B libs: bash scripts/gpu_prof_lib.sh name1 name2 ...
mkdir -p gpurun_out
for name in "$@"; do
  SVGT_LIB=$PWD/svtyper_b200/ab/libsvgt_$name.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:svgt_tally -s 3 -c 1 -f -o gpurun_out/prof_ab_$name python bench.py --sites 200000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/prof_ab_$name.out 2>&1
  echo "$name rc=$?"
done
