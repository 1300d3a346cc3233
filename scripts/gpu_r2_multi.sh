#!/bin/bash
# round 2, N GPUs: the 2-GPU shard test and bench.py under torchrun (weak + strong in one run)
set -u
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_shard_nccl.py -m gpu -q -s > gpurun_out/pytest_nccl.log 2>&1
echo "nccl test rc=$?"; tail -4 gpurun_out/pytest_nccl.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 ${EXTRA:-} > gpurun_out/bench_${N}gpu${TAG:-}.json 2> gpurun_out/bench_${N}gpu${TAG:-}.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_${N}gpu${TAG:-}.json; tail -5 gpurun_out/bench_${N}gpu${TAG:-}.err
