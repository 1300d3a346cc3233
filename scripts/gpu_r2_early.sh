#!/bin/bash
# round 2: fp64 add latency, parity of the early call / piece plan, kernel time of the config shapes with the
# genotype call of long sites made early (call_early_rows) or not
set -u
mkdir -p gpurun_out
C=/dev/shm/svgt_cache
./scripts/ub/ub_fp64_latency > gpurun_out/ub_fp64_latency.txt 2>&1; cat gpurun_out/ub_fp64_latency.txt
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${K:-early or piece or plan or hazard or classic or error_flags}" > gpurun_out/pytest_early.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_early.log
fi
for cs in ${SHAPES:-del10k:10000 mixed100k:100000 stress1m:125000 del1m4lib:125000 del1m4lib:1000000}; do
  cfg=${cs%%:*}; n=${cs#*:}
  timeout 900 python scripts/gpu_compact_check.py --config $cfg --sites $n --cache $C --skip-parity --no-e2e --no-wide --steps 30 --pieces ${VARIANTS:-off/0,off/auto} --tag early_${cfg}_${n} > gpurun_out/early_${cfg}_${n}.log 2>&1
  echo "$cfg $n rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/compact_check_early_${cfg}_${n}.json"))
for k, v in d.items():
    if isinstance(v, dict) and "ms_avg" in v:
        print("  %-28s %.4f ms (min %.4f) early=%s unit_mode=%s same=%s" % (k, v["ms_avg"], v["ms_min"], v.get("call_early_rows"), v.get("unit_mode"), v.get("identical_to_first")))
PY
done
