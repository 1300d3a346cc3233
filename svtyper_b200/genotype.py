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


class RowFormatter(object):
    """Fast text path for the single-sample entry point: the sample column of every
    scored row straight from the output arrays, bypassing the per-field SampleCall machinery.  Same text as
    apply_row(..., classic=False) + VcfRecord.render(); records it cannot take (other samples in the
    header, FORMAT values already present) go through the generic path."""

    def __init__(self, header, sample_name, rows):
        self.ok = header.samples == [sample_name]
        self.sample = sample_name
        keys = set(["GT", "GQ", "SQ", "GL", "AB"] + list(COUNT_FIELDS))
        self.order = [k for k in header.format_ids() if k in keys]
        if len(self.order) != len(keys):
            self.ok = False
        self.fmt = ":".join(self.order)
        if not self.ok:
            return
        self.gt = rows["GT"].tolist()
        self.gq = rows["GQ"].tolist()
        self.sq = rows["SQ"].tolist()
        self.gl = rows["GL"].tolist()
        self.counts = {k: rows[k].tolist() for k in COUNT_FIELDS}
        blank = dict(_BLANK_ITEMS)
        self.blank_call = ":".join(str(blank[k]) for k in self.order)

    def eligible(self, rec):
        call = rec.calls.get(self.sample)
        return (self.ok and rec.active_formats == ["GT"] and call is not None and len(call.values) == 1
                and len(rec.calls) == 1)

    def columns(self, rec, idx):
        """(QUAL after the call, FORMAT column, sample column) of scored row `idx` for record `rec`."""
        gt = self.gt[idx]
        if gt == GT_BLANK or gt == GT_SKIPPED:
            return rec.qual, self.fmt, self.blank_call
        vals = {k: v[idx] for k, v in self.counts.items()}
        total = vals["QR"] + vals["QA"]
        vals["AB"] = ("%.2g" % (vals["QA"] / float(total))) if total else "."
        vals["GL"] = ",".join("%.0f" % x for x in self.gl[idx])
        qual = rec.qual
        if gt == GT_UNDERFLOW:
            vals["GQ"] = "."; vals["SQ"] = "."; vals["GT"] = "./."
        else:
            sq = self.sq[idx]
            vals["GQ"] = self.gq[idx]; vals["SQ"] = "%0.2f" % sq; vals["GT"] = GT_TEXT[gt]
            qual = qual + sq
        return qual, self.fmt, ":".join(str(vals[k]) for k in self.order)

    @staticmethod
    def line(rec, qual, fmt, call):
        return "\t".join((rec.chrom, str(rec.pos), rec.var_id, rec.ref, rec.alt, "%0.2f" % qual, rec.filter,
                          rec.info_string(), fmt, call))


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
