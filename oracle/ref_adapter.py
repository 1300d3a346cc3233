"""Feed evidence *arrays* through the REFERENCE's own scoring functions.

TEST INFRASTRUCTURE ONLY.  Rebuilds, from an `EvidenceBatch`, duck-typed reads and
genuine reference `SamFragment` / `SplitRead` instances (constructed without their
pysam-dependent `__init__`), then calls the reference's
`tally_variant_read_fragments` + `bayesian_genotype` (reference
singlesample.py:355, :406) on them.  This is how synthetic batches -- which have no
BAM behind them -- are scored by the reference itself, both to validate the C
restatement and as the `bench.py --impl reference` CPU arm.
"""
from __future__ import annotations

from collections import Counter

import numpy as np

from svtyper_b200 import evidence as ev

_SVNAME = {ev.SV_DEL: "DEL", ev.SV_DUP: "DUP", ev.SV_INV: "INV", ev.SV_BND: "BND"}


class FakeRead(object):
    __slots__ = ("reference_name", "reference_start", "reference_end", "is_reverse",
                 "mapping_quality", "intervals")

    def __init__(self, tid, start, end, is_reverse, mapq, intervals=None):
        self.reference_name = "t%d" % tid
        self.reference_start = int(start)
        self.reference_end = int(end)
        self.is_reverse = bool(is_reverse)
        self.mapping_quality = int(mapq)
        self.intervals = intervals or [(int(start), int(end))]

    def get_overlap(self, start, end):
        tot = 0
        for s, e in self.intervals:
            o = min(e, end) - max(s, start)
            if o > 0:
                tot += o
        return tot


class FakePiece(object):
    __slots__ = ("chrom", "reference_start", "reference_end", "mapping_quality")

    def __init__(self, tid, start, end, mapq):
        self.chrom = None if tid == ev.TID_NONE else "t%d" % tid
        self.reference_start, self.reference_end, self.mapping_quality = int(start), int(end), int(mapq)


class FakeLib(object):
    def __init__(self, name, mean, sd, hist):
        self.name, self.mean, self.sd = name, float(mean), float(sd)
        total = sum(hist.values())
        self.dens = Counter()
        for k in list(hist):
            self.dens[k] = float(hist[k]) / total      # parsers.py:579-583


def make_libs(lib_table):
    return [FakeLib("lib%d" % i, m, s, h) for i, (m, s, h) in enumerate(lib_table.sources)]


def site_inputs(ref, batch, i, libs):
    """(breakpoint dict, {qname: SamFragment}) for site i of the batch."""
    s = batch.sites[i]
    meta = int(s[9])
    svtype = _SVNAME[meta & 3]
    bp = {"id": "site%d" % i, "svtype": svtype,
          "A": {"chrom": "t%d" % s[6], "pos": int(s[0]), "ci": [int(s[2]), int(s[3])],
                "is_reverse": bool(meta & ev.SITE_O1_REV)},
          "B": {"chrom": "t%d" % s[7], "pos": int(s[1]), "ci": [int(s[4]), int(s[5])],
                "is_reverse": bool(meta & ev.SITE_O2_REV)}}
    if svtype == "DEL":
        bp["var_length"] = int(s[8])
    foff = int(s[10:12].copy().view(np.int64)[0])
    soff = int(s[13:15].copy().view(np.int64)[0])
    SamFragment, SplitRead = ref.parsers.SamFragment, ref.parsers.SplitRead
    frags = []
    pend = {0: [], 1: []}
    for j in range(int(s[12])):
        f = batch.frags[foff + j]
        fl = int(f[7])
        if fl & ev.F_EXTRA:
            if fl & ev.F_HAS_A:
                pend[0].append((int(f[0]), int(f[1])))
            if fl & ev.F_HAS_B:
                pend[1].append((int(f[2]), int(f[3])))
            continue
        reads = []
        if fl & ev.F_HAS_A:
            reads.append(FakeRead(f[4], f[0], f[1], fl & ev.F_REV_A, f[6] & 0xFF,
                                  pend[0] if fl & ev.F_MULTI_A else None))
        if fl & ev.F_HAS_B:
            reads.append(FakeRead(f[5], f[2], f[3], fl & ev.F_REV_B, (f[6] >> 8) & 0xFF,
                                  pend[1] if fl & ev.F_MULTI_B else None))
        pend = {0: [], 1: []}
        if fl & ev.F_CONT and frags:
            fr = frags[-1]
            fr.primary_reads.extend(reads)
            fr.num_primary += len(reads)
            continue
        fr = SamFragment.__new__(SamFragment)
        fr.lib = libs[(int(f[6]) >> 16) & 0xFFFF]
        fr.primary_reads, fr.split_reads = reads, []
        fr.num_primary = len(reads)
        fr.query_name = "f%08d" % len(frags)
        fr.readA = fr.readB = None
        if fr.num_primary == 2:
            fr.readA, fr.readB = reads
        frags.append(fr)
    group = -1
    for j in range(int(s[15])):
        q = batch.splits[soff + j]
        sfl = (int(q[6]) >> 16) & 0xFFFF
        if sfl & ev.S_FIRST or group < 0:
            group += 1
        sp = SplitRead.__new__(SplitRead)
        sp.query_left = FakePiece(q[0], q[1], q[2], q[6] & 0xFF)
        sp.query_right = FakePiece(q[3], q[4], q[5], (q[6] >> 8) & 0xFF)
        sp.is_soft_clip = bool(sfl & ev.S_SOFT_CLIP)
        frags[group].split_reads.append(sp)
    return bp, {fr.query_name: fr for fr in frags}


def reference_score_site(ref, bp, fragments, split_weight=1, disc_weight=1, min_aligned=20,
                         split_slop=3):
    """The reference's scoring segment for one breakpoint (singlesample.py:523-536)."""
    ss = ref.singlesample
    counts = ss.tally_variant_read_fragments(split_slop, min_aligned, bp, fragments, False)
    if sum(counts[k] for k in counts) == 0:
        return counts, ss.blank_genotype_result()
    return counts, ss.bayesian_genotype(bp, counts, split_weight, disc_weight, False)


def result_to_row(ref, bp, counts, result):
    """Numeric OUT_DTYPE row from a reference result dict (GL re-derived via bayes_gt)."""
    row = np.zeros((), dtype=ev.OUT_DTYPE)
    fm = result["formats"]
    if fm["GL"] == ".":
        row["GT"], row["GQ"] = ev.GT_BLANK, -1
        return row
    for k in ("DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP"):
        row[k] = fm[k]
    row["GL"] = ref.statistics.bayes_gt(fm["QR"], fm["QA"], bp["svtype"] == "DUP")
    if fm["GT"] == "./.":
        row["GT"], row["GQ"] = ev.GT_UNDERFLOW, -1
    else:
        row["GT"] = {"0/0": 0, "0/1": 1, "1/1": 2}[fm["GT"]]
        row["GQ"], row["SQ"] = fm["GQ"], fm["SQ"]
    return row


def reference_score(ref, batch, sites=None, cache=None, **kw):
    """Score (a subset of) a batch with the reference; returns OUT_DTYPE rows.

    `cache` (dict) keeps the rebuilt (breakpoint, fragments) objects per site, so a timed
    second pass measures only the reference's own scoring functions, not this adapter:
    cache["seconds"] is the time spent inside tally_variant_read_fragments + bayesian_genotype
    alone -- building the inputs and turning the result dicts into rows (which calls bayes_gt a
    second time for the numeric GL) are outside it.
    """
    import time
    libs = make_libs(batch.libs) if cache is None else cache.setdefault("libs", make_libs(batch.libs))
    idx = range(batch.n_sites) if sites is None else sites
    out = np.zeros(len(idx), dtype=ev.OUT_DTYPE)
    todo = []
    for k, i in enumerate(idx):
        if int(batch.sites[i, 9]) & ev.SITE_SKIP:
            out[k]["GT"], out[k]["GQ"] = ev.GT_SKIPPED, -1
            continue
        if cache is not None and i in cache:
            bp, frags = cache[i]
        else:
            bp, frags = site_inputs(ref, batch, i, libs)
            if cache is not None:
                cache[i] = (bp, frags)
        todo.append((k, bp, frags))
    t0 = time.perf_counter()
    done = [reference_score_site(ref, bp, frags, **kw) for _, bp, frags in todo]
    seconds = time.perf_counter() - t0
    if cache is not None:
        cache["seconds"] = seconds
    for (k, bp, _), (counts, res) in zip(todo, done):
        out[k] = result_to_row(ref, bp, counts, res)
    return out
