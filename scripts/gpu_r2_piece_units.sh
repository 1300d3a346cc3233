#!/bin/bash
# round 2: parity subset, then kernel time of config shapes with piece plans scored in full 6-entry units
set -u
mkdir -p gpurun_out
C=/dev/shm/svgt_cache
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${K:-fixture or synthetic or hazard or piece or plan or classic or idempotent}" > gpurun_out/pytest_subset.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_subset.log
fi
for cs in ${SHAPES:-del10k:10000 mixed100k:100000 stress1m:125000 del1m4lib:125000 del1m4lib:1000000}; do
  cfg=${cs%%:*}; n=${cs#*:}
  v=${VARIANTS:-off,4/1,8/1,16/1,auto/1}
  if [ "$n" = "1000000" ]; then v=off; fi
  timeout 900 python scripts/gpu_compact_check.py --config $cfg --sites $n --cache $C --skip-parity --no-e2e --no-wide --steps 30 --pieces $v --tag v_${cfg}_${n} > gpurun_out/v_${cfg}_${n}.log 2>&1
  echo "$cfg $n rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/compact_check_v_${cfg}_${n}.json"))
for k, v in d.items():
    if isinstance(v, dict) and "ms_avg" in v:
        print("  %-28s %.4f ms (min %.4f) unit_mode=%s plan=%s same=%s" % (k, v["ms_avg"], v["ms_min"], v.get("unit_mode"), v.get("plan"), v.get("identical_to_first")))
PY
done
