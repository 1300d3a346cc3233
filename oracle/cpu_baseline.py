"""CPU baseline leg of bench.py: the reference's own scoring segment on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never on the product path).

Times `tally_variant_read_fragments` + `bayesian_genotype` of the (py3-patched, otherwise
unmodified) reference under oracle/_ref, fed the same evidence arrays as the GPU kernel
through oracle/ref_adapter.py.  Sites are split into contiguous batches, one per worker
process, like the reference's `genotype_parallel` fan-out (svtyper/singlesample.py:710-751).
The adapter's rebuilding of reference `SamFragment` objects from the arrays is cached by an
untimed warm-up pass: a timed pass runs only the reference's own two functions.
When oracle/_ref is absent, falls back to the C restatement (kind "port").
"""
from __future__ import annotations

import os
import time

import numpy as np


def _worker(conn, sites, frags, splits, lib_sources, lo, hi):
    from svtyper_b200 import evidence as ev
    from oracle import ref_adapter, ref_loader
    batch = ev.EvidenceBatch(sites, frags, splits, ev.LibraryTable(lib_sources))
    ref = ref_loader.load()
    cache = {}
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        rows = ref_adapter.reference_score(ref, batch, range(lo, hi), cache=cache)
        conn.send((cache["seconds"], rows.tobytes()))        # the reference's two functions only
    conn.close()


class ReferencePool(object):
    """Persistent worker processes holding one sample batch; each `step()` scores all of it."""

    def __init__(self, sample, cores=None):
        import multiprocessing as mp
        from oracle import ref_loader
        self.sample = sample
        self.cores = int(cores or os.cpu_count() or 1)
        self.kind = "reference" if ref_loader.ensure() else "port"
        self.workers = []
        if self.kind != "reference":
            return
        repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.environ["PYTHONPATH"] = repo + os.pathsep + os.environ.get("PYTHONPATH", "")
        ctx = mp.get_context("forkserver")
        n = sample.n_sites
        nw = max(1, min(self.cores, n))
        bounds = [n * i // nw for i in range(nw + 1)]
        for w in range(nw):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(child, sample.sites, sample.frags, sample.splits,
                                                  sample.libs.sources, bounds[w], bounds[w + 1]), daemon=True)
            p.start()
            self.workers.append((p, parent))
        for _, c in self.workers:
            assert c.recv() == "ready"

    def step(self):
        """Score the whole sample once; returns (seconds, OUT_DTYPE rows).  Reference: the seconds are the slowest
        worker's time inside tally_variant_read_fragments + bayesian_genotype (the workers run side by side; the
        adapter's conversions either side of the two functions and the pipes are not the reference's work)."""
        from svtyper_b200 import evidence as ev
        t0 = time.perf_counter()
        if self.kind == "reference":
            for _, c in self.workers:
                c.send("go")
            got = [c.recv() for _, c in self.workers]
            rows = np.frombuffer(b"".join(g[1] for g in got), dtype=ev.OUT_DTYPE)
            return max(g[0] for g in got), rows
        from oracle import oracle
        rows = oracle.score(self.sample, n_threads=self.cores)
        return time.perf_counter() - t0, rows

    def close(self):
        for p, c in self.workers:
            try:
                c.send("stop")
            except Exception:
                pass
        for p, c in self.workers:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
        self.workers = []
