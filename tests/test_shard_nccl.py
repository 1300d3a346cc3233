"""GPU, world_size 2 over NCCL (needs two GPUs; skipped otherwise): the 2-GPU twin of test_shard_gloo.py.
Each rank scores its row-balanced shard of one global batch with the CUDA engine; the rows reach rank 0 (i) by
the NCCL gather of shard.gather_rows and (ii) by the call kernels storing straight into rank 0's peer-mapped
buffer (bench.RowGather, the path bench.py times).  Both must equal the one-GPU rows and the oracle."""
import os
import subprocess
import sys

import pytest

from util import REPO

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank = int(sys.argv[3])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2,
                        device_id=torch.device("cuda", rank))
import bench
from svtyper_b200 import compact as cp, engine, shard, synth, evidence as ev
from oracle import oracle
batch = synth.generate("stress1m", n_sites=4000, seed=4)
cb = cp.compact_from_wide(batch)
local, bounds = shard.local_shard(cb, rank, 2)
eng = engine.Engine(rank)
dev = eng.upload(local)
eng.score(dev)
rows = eng.rows(dev)
full = shard.gather_rows(rows, bounds, rank, 2, device=torch.device("cuda", rank))
want = oracle.score(batch, n_threads=4)
one = eng.upload(cb)                               # the whole batch on this rank's GPU: what the shards must add up to
eng.score(one)
exp = eng.rows(one)
for k in ("GT", "GQ", "DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP", "GL"):
    assert np.array_equal(exp[k], want[k]), k      # one GPU == oracle (SQ: device pow / log, compared to 1e-9 elsewhere)
if rank == 0:
    got = shard.rows_from_tensor(full)
    assert got.tobytes() == exp.tobytes(), "nccl gather"
# the collective-free routes: rows forwarded by the copy engine under the next step ("dma"), or stored by the call
# kernel itself ("peer"), into rank 0's IPC-mapped buffer, one flag per rank
counts = [bounds[1] - bounds[0], bounds[2] - bounds[1]]
stream = torch.cuda.current_stream()
for mode in ("dma", "peer"):
    g = bench.RowGather(mode, rank, 2, counts, torch.device("cuda", rank))
    for step in range(3):
        g.before_score(stream, step & 1)
        g.arm(dev.desc, step & 1)
        eng.score(dev, stream, out=g.out_tensor(dev.out, step & 1))
        g.after_score(stream, step & 1)
    g.drain(stream)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        got = g.gathered_rows(0)
        assert got.tobytes() == exp.tobytes(), "%s gather (%s)" % (mode, g.mode)
        print("gather mode", mode, "->", g.mode)
    dist.barrier()
    g.close()
    dev.desc.out_final = None
    dev.desc.done_flag = None
dist.destroy_process_group()
'''


def test_two_gpu_shards_match_one_gpu_and_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), REPO, port, str(r)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
