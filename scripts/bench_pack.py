#!/usr/bin/env python
"""Host-side measurement of SURVEY.md 8f row 1: evidence extraction (gather + split QC + row packing)
of the reference fixture's 211 breakpoints, replicated, through the native packer (libsvgt_pack.so) at
1..N threads and through the Python gather path it replaces.  Prints one JSON line; rows are asserted
identical first.  CPU only."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from svtyper_b200 import gather, genotype, packer          # noqa: E402
from svtyper_b200.sample import SampleInfo                 # noqa: E402
import test_pack_native as t                               # noqa: E402


def main(reps=20):
    plan = t.make_plan()
    s = SampleInfo.open(t.BAM, t.LIB, None, 1000000)
    t0 = time.perf_counter()
    want = genotype.pack_sample_python(s, plan, lambda smp, bp: gather.gather_sso(smp, bp, genotype.Z, 1000), 20)
    t_py = time.perf_counter() - t0
    got = packer.pack_sample(s, plan, packer.MODE_SSO, 1000, genotype.Z)
    assert np.array_equal(got.sites, want.sites) and np.array_equal(got.frags, want.frags)
    assert np.array_equal(got.splits, want.splits)

    class Plan(object):
        pass
    big = Plan()
    big.breakpoints = plan.breakpoints * reps
    native = {}
    for th in sorted(set([1, 2, 4, 8, os.cpu_count() or 1])):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            b = packer.pack_sample(s, big, packer.MODE_SSO, 1000, genotype.Z, threads=th)
            best = min(best, time.perf_counter() - t0)
        native[str(th)] = {"sites_per_s": b.n_sites / best, "rows_per_s": (b.n_frag + b.n_split) / best}
    print(json.dumps({
        "metric": "breakpoints_gathered_and_packed_per_sec", "unit": "breakpoints/s", "mode": "sso, max_reads=1000",
        "fixture": "tests/data NA12878.target_loci.sorted.bam, %d breakpoints x %d" % (len(plan.breakpoints), reps),
        "rows_per_breakpoint": (want.n_frag + want.n_split) / float(want.n_sites),
        "python_gather": {"sites_per_s": want.n_sites / t_py, "threads": 1},
        "native_packer_by_threads": native, "host_cores": os.cpu_count(),
        "rows_identical_to_python_gather": True}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20)
