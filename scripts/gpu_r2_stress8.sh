#!/bin/bash
# round 2: BASELINE.json configs[4] -- 1M sites x max_reads=10000 ragged-evidence stress, site-sharded over N GPUs by rows
set -u
N=${N:-8}
mkdir -p gpurun_out
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --config stress1m --sites ${SITES:-1000000} --strong-only --steps 10 --warmup 3 --gather dma \
  > gpurun_out/bench_stress1m_${N}gpu.json 2> gpurun_out/bench_stress1m_${N}gpu.err
echo "stress rc=$?"; tail -c 2500 gpurun_out/bench_stress1m_${N}gpu.json; tail -5 gpurun_out/bench_stress1m_${N}gpu.err
free -g | head -2
