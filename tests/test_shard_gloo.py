"""CPU, world_size 2 over gloo: the multi-rank host logic (row-balanced contiguous site shards,
one gather of the 80-byte output rows to rank 0).  The per-shard scorer here is the oracle
(checker) because there is no GPU in this container; the N>1 GPU path runs the same shard /
gather code with the CUDA engine: bench.py's strong-scaling mode partitions with shard.shard_bounds, and
tests/test_shard_nccl.py is the 2-GPU twin of this test."""
import os
import subprocess
import sys

import numpy as np

from svtyper_b200 import shard, synth
from util import REPO

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from svtyper_b200 import compact as cp, shard, synth, evidence as ev
from oracle import oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
batch = synth.generate("stress1m", n_sites=600, seed=4)
cb = cp.compact_from_wide(batch)                    # the product path shards compact batches
local, bounds = shard.local_shard(cb, rank, 2)
assert local.order is not None and local.n_rows == int(local.work().sum())
rows = oracle.score(cp.wide_from_compact(local))
full = shard.gather_rows(rows, bounds, rank, 2)
if rank == 0:
    got = shard.rows_from_tensor(full)
    exp = oracle.score(batch)
    assert got.tobytes() == exp.tobytes()
    np.save(sys.argv[4], np.array(bounds))
dist.barrier()
dist.destroy_process_group()
'''


def test_bounds_balance_rows_not_sites():
    from svtyper_b200 import compact as cp
    b = synth.generate("stress1m", n_sites=2000, seed=4)
    bounds = shard.shard_bounds(b, 4)
    cbounds = shard.shard_bounds(cp.compact_from_wide(b), 8)
    cw = cp.compact_from_wide(b).work()
    cper = [int(cw[cbounds[i]:cbounds[i + 1]].sum()) for i in range(8)]
    assert cbounds[0] == 0 and cbounds[-1] == b.n_sites and max(cper) < 1.6 * (sum(cper) / 8.0)
    assert bounds[0] == 0 and bounds[-1] == b.n_sites and bounds == sorted(bounds)
    work = b.sites[:, 12].astype(np.int64) + b.sites[:, 15]
    per = [int(work[bounds[i]:bounds[i + 1]].sum()) for i in range(4)]
    assert max(per) < 1.5 * (sum(per) / 4.0)
    assert shard.shard_bounds(synth.generate("del10k", n_sites=0), 3) == [0, 0, 0, 0]


def test_two_rank_gather_matches_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "bounds.npy"
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), REPO, port, str(r), str(out)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    bounds = np.load(out)
    assert bounds[0] == 0 and bounds[-1] == 600 and 0 < bounds[1] < 600
