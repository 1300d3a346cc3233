"""Shared machinery of the two drop-in entry points (classic.sv_genotype, singlesample.sso_genotype).

The reference walks the VCF and, per breakpoint, gathers reads, tallies evidence and calls the
genotype in Python (svtyper/classic.py:212-521, svtyper/singlesample.py:577-652).  Here the walk
and the read gathering stay on the host, but the tally + call of ALL breakpoints of a sample are
one batch for the CUDA engine: gather -> `evidence.BatchPacker` -> `Engine.score_host` -> rows ->
FORMAT fields.  There is no CPU scoring path in this package; without a CUDA device the engine
raises.
"""
from __future__ import annotations

import sys

from . import evidence as ev
from .evidence import GT_BLANK, GT_SKIPPED, GT_UNDERFLOW

Z = 3               # fetch flank in standard deviations (reference classic.py:183)
SPLIT_SLOP = 3      # reference classic.py:184, singlesample.py:793
GT_TEXT = {0: "0/0", 1: "0/1", 2: "1/1"}
COUNT_FIELDS = ("DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP")

_scorer_override = None     # tests may install a checker-backed scorer to exercise the plumbing on CPU
_engine = None


def set_scorer(fn):
    """Install `fn(batch, **params) -> OUT_DTYPE rows` in place of the CUDA engine (tests only)."""
    global _scorer_override
    _scorer_override = fn


def score(batch, **params):
    """Score one EvidenceBatch on the GPU (or through the installed test scorer)."""
    global _engine
    if _scorer_override is not None:
        return _scorer_override(batch, **params)
    if _engine is None:
        from .engine import Engine
        _engine = Engine()              # raises without libsvgt.so / a CUDA device
    return _engine.score_host(batch, **params)


def _count_items(row):
    items = [(key, int(row[key])) for key in COUNT_FIELDS]
    total = int(row["QR"]) + int(row["QA"])
    items.append(("AB", "%.2g" % (int(row["QA"]) / float(total)) if total else "."))
    return items


_BLANK_ITEMS = ([("GT", "./."), ("GQ", "."), ("SQ", "."), ("GL", ".")] +
                [(key, 0) for key in ("DP", "AO", "RO", "AS", "ASC", "RS", "AP", "RP", "QR", "QA")] + [("AB", ".")])


def apply_row(rec, sample_name, row, classic):
    """Write one scored row into `rec`'s FORMAT fields for `sample_name`.

    classic=True follows classic.py:437-513 (a too-many-reads site only gets GT './.'; a site
    with no evidence resets QUAL to 0); classic=False follows bayesian_genotype +
    assign_genotype_to_variant (singlesample.py:406-473, :544-575), where every non-called
    outcome is the blank row.  All fields of the row go in through one SampleCall.set_many().
    """
    call = rec.call(sample_name)
    gt = int(row["GT"])
    if gt == GT_SKIPPED and classic:
        call.set("GT", "./.")
        return
    if gt in (GT_BLANK, GT_SKIPPED):
        if classic:
            rec.qual = 0
        call.set_many(_BLANK_ITEMS)
        return
    items = [("GL", ",".join("%.0f" % x for x in row["GL"]))] + _count_items(row)
    if gt == GT_UNDERFLOW:
        items += [("GQ", "."), ("SQ", "."), ("GT", "./.")]
        call.set_many(items)
        return
    sq = float(row["SQ"])
    items += [("GQ", int(row["GQ"])), ("SQ", sq), ("GT", GT_TEXT[gt])]
    call.set_many(items)
    rec.qual += sq


class SitePlan(object):
    """Output order of a VCF: pass-through records and genotyped sites (one or two records)."""

    def __init__(self):
        self.entries = []       # ("raw", rec) | ("site", rec, mate_or_None, site_index)
        self.breakpoints = []

    def passthrough(self, rec):
        self.entries.append(("raw", rec, None, -1))

    def site(self, rec, mate, breakpoint):
        self.entries.append(("site", rec, mate, len(self.breakpoints)))
        self.breakpoints.append(breakpoint)


def warn(msg):
    sys.stderr.write(msg)


def pack_sample_python(sample, plan, gather, min_aligned):
    """Gather + pack every planned site of one sample into an EvidenceBatch (Python path: the parity
    checker of the native packer, and the route for inputs the native reader does not open)."""
    packer = ev.BatchPacker(sample.bam.gettid, sample.library_table())
    for bp in plan.breakpoints:
        fragments, too_many = gather(sample, bp)
        packer.add_site(bp, fragments, skip=too_many)
    return packer.finish()


def pack_sample(sample, plan, gather, min_aligned, mode=None, max_reads=None):
    """Gather + pack every planned site of one sample.  With `mode` (packer.MODE_SSO / MODE_CLASSIC)
    and an indexed .bam on disk the native packer (libsvgt_pack.so) does it in one call; otherwise the
    Python gather `gather(sample, breakpoint)` runs per site."""
    if mode is not None:
        from . import packer as native_packer
        if native_packer.usable(sample):
            return native_packer.pack_sample(sample, plan, mode, max_reads, Z)
    return pack_sample_python(sample, plan, gather, min_aligned)
