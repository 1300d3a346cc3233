"""Shared machinery of the two drop-in entry points (classic.sv_genotype, singlesample.sso_genotype).

The reference walks the VCF and, per breakpoint, gathers reads, tallies evidence and calls the
genotype in Python (svtyper/classic.py:212-521, svtyper/singlesample.py:577-652).  Here the walk
and the read gathering stay on the host, but the tally + call are batched for the CUDA engine:

    VCF records --walk--> chunks of `batch_size` breakpoints
        chunk --pack (native packer, `cores` threads; one evidence batch per sample)
              --merge (site x sample: ONE compact batch, library tables concatenated)
              --Engine.score_host (pinned host rows -> H2D -> kernels -> D2H)
              --format (native FORMAT text of all rows at once) --> VCF lines

Chunks are pipelined: while chunk k is scored and written, chunk k + 1 is being packed on a worker
thread (the native packer releases the GIL), so host memory is bounded by two chunks whatever the
size of the VCF.  There is no CPU scoring path in this package; without a CUDA device the engine
raises.
"""
from __future__ import annotations

import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import compact as cp
from . import evidence as ev
from .evidence import GT_BLANK, GT_SKIPPED, GT_UNDERFLOW

Z = 3               # fetch flank in standard deviations (reference classic.py:183)
SPLIT_SLOP = 3      # reference classic.py:184, singlesample.py:793
GT_TEXT = {0: "0/0", 1: "0/1", 2: "1/1"}
COUNT_FIELDS = ("DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP")
CALL_KEYS = ("GT", "GQ", "SQ", "GL", "AB") + COUNT_FIELDS
DEFAULT_BATCH = 1000    # reference --batch_size default (singlesample.py:39)

_engine = None
_arena = None


def engine():
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine()              # raises without libsvgt.so / a CUDA device
    return _engine


def score(batch, **params):
    """Score one CompactBatch on the GPU: OUT_DTYPE rows in site order.  (The CPU tests of the host plumbing
    replace this function with the parity oracle; the product has no other scorer.)"""
    return engine().score_host(batch, **params)


def pinned_allocs():
    """Two alloc(name, shape, dtype) callables handing out views of two reusable pinned host arenas (chunk k
    is packed into one while chunk k - 1 is copied to the GPU from the other); (None, None) when torch / CUDA
    is not there (the CPU tests of the host plumbing)."""
    global _arena
    try:
        from .engine import PinnedArena
        if _arena is None:
            _arena = (PinnedArena(), PinnedArena())
        return _arena[0].alloc, _arena[1].alloc
    except Exception:
        return None, None


def _count_items(row):
    items = [(key, int(row[key])) for key in COUNT_FIELDS]
    total = int(row["QR"]) + int(row["QA"])
    items.append(("AB", "%.2g" % (int(row["QA"]) / float(total)) if total else "."))
    return items


_BLANK_ITEMS = ([("GT", "./."), ("GQ", "."), ("SQ", "."), ("GL", ".")] +
                [(key, 0) for key in ("DP", "AO", "RO", "AS", "ASC", "RS", "AP", "RP", "QR", "QA")] + [("AB", ".")])


def apply_row(rec, sample_name, row, classic):
    """Write one scored row into `rec`'s FORMAT fields for `sample_name` (the generic, per-field path).

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
    """Output order of (a chunk of) a VCF: pass-through records and genotyped sites (one or two records)."""

    def __init__(self):
        self.entries = []       # ("raw", rec, None, -1) | ("site", rec, mate_or_None, site_index)
        self.breakpoints = []

    def passthrough(self, rec):
        self.entries.append(("raw", rec, None, -1))

    def site(self, rec, mate, breakpoint):
        self.entries.append(("site", rec, mate, len(self.breakpoints)))
        self.breakpoints.append(breakpoint)


def warn(msg):
    sys.stderr.write(msg)


# ---------------------------------------------------------------------------------------------
# pack
def pack_sample_python(sample, plan, gather, min_aligned):
    """Gather + pack every planned site of one sample into a wide EvidenceBatch (Python path: the parity
    checker of the native packer, and the route for inputs the native reader does not open)."""
    packer = ev.BatchPacker(sample.bam.gettid, sample.library_table())
    for bp in plan.breakpoints:
        fragments, too_many = gather(sample, bp)
        packer.add_site(bp, fragments, skip=too_many)
    return packer.finish()


def pack_sample(sample, plan, gather, min_aligned, mode=None, max_reads=None, threads=0):
    """Gather + pack every planned site of one sample -> wide EvidenceBatch.  With `mode` (packer.MODE_SSO /
    MODE_CLASSIC) and an indexed .bam on disk the native packer (libsvgt_pack.so) does it in one call on
    `threads` threads; otherwise the Python gather `gather(sample, breakpoint)` runs per site."""
    if mode is not None:
        from . import packer as native_packer
        if native_packer.usable(sample):
            return native_packer.pack_sample(sample, plan, mode, max_reads, Z, threads=threads)
    return pack_sample_python(sample, plan, gather, min_aligned)


def to_compact(batch, min_aligned, alloc=None, threads=0):
    """Wide -> compact rows (natively when libsvgt_pack.so is there; numpy otherwise: same bytes)."""
    if isinstance(batch, cp.CompactBatch):
        return batch
    from . import packer as native_packer
    if native_packer.available():
        return native_packer.compact_from_wide(batch, min_aligned=min_aligned, alloc=alloc, threads=threads)
    return cp.compact_from_wide(batch, alloc=alloc, min_aligned=min_aligned)


def merge_samples(batches, alloc=None):
    """(site x sample) as ONE batch: the compact batches of several samples over the same sites, concatenated
    sample-major, their library tables joined (fragment rows re-pointed at their sample's libraries).
    Returns the merged batch; rows [s * n, (s + 1) * n) of its scores belong to sample s."""
    if len(batches) == 1:
        return batches[0]
    alloc = alloc or (lambda name, shape, dtype: np.empty(shape, dtype=dtype))
    libs = ev.LibraryTable([src for b in batches for src in b.libs.sources])
    if libs.n_lib - 1 > cp.LIB_MAX:
        raise ValueError("compact schema holds library indices up to %d" % cp.LIB_MAX)
    n = batches[0].n_sites
    total_rows = sum(b.n_rows for b in batches)
    sites = alloc("sites", (n * len(batches), cp.CSITE_WORDS), np.int32)
    rows = alloc("rows", (total_rows, cp.CROW_WORDS), np.int32)
    r0 = lib0 = 0
    for k, b in enumerate(batches):
        assert b.n_sites == n and b.min_aligned == batches[0].min_aligned
        s = b.sites.copy()
        off = b.row_offsets() + r0
        s[:, 8:10] = off.view(np.int32).reshape(-1, 2)
        sites[k * n:(k + 1) * n] = s
        dst = rows[r0:r0 + b.n_rows]
        dst[:] = b.rows
        if lib0:                                        # fragment rows: library index lives in bits 16..24 of word 3
            nf = b.sites[:, 10].astype(np.int64)
            tot = nf + b.sites[:, 11]
            start = np.cumsum(tot) - tot
            is_frag = np.zeros(b.n_rows + 1, dtype=np.int32)
            np.add.at(is_frag, start, 1)
            np.add.at(is_frag, start + nf, -1)
            mask = np.cumsum(is_frag[:-1]) > 0
            dst[mask, 3] += np.int32(lib0 << cp.LIB_SHIFT)
        r0 += b.n_rows
        lib0 += b.libs.n_lib
    out = cp.CompactBatch.__new__(cp.CompactBatch)
    out.sites, out.rows, out.libs, out.order, out.min_aligned = sites, rows, libs, None, batches[0].min_aligned
    order = alloc("order", (sites.shape[0],), np.int32)
    order[:] = out.length_order()
    out.order = order
    return out


# ---------------------------------------------------------------------------------------------
# write
class ChunkWriter(object):
    """VCF text of a scored chunk.  Records whose sample columns are exactly this run's samples and carry
    nothing but GT (every record of an ordinary input VCF) get their FORMAT text from the native formatter --
    all rows of the chunk in one call -- and their QUAL from array arithmetic; anything else (other samples in
    the file, FORMAT values already present) goes through the per-field record model (apply_row)."""

    def __init__(self, header, sample_names, classic):
        self.header, self.samples, self.classic = header, list(sample_names), classic
        self.fast_header = header.samples == self.samples
        keys = set(CALL_KEYS)
        self.order = [k for k in header.format_ids() if k in keys]
        if len(self.order) != len(keys):
            self.fast_header = False
        self.fmt_full = ":".join(self.order)
        from . import packer as native_packer
        self.native = native_packer if native_packer.available() else None

    def eligible(self, rec):
        if not self.fast_header or rec.active_formats != ["GT"] or len(rec.calls) != len(self.samples):
            return False
        return all(len(c.values) == 1 for c in rec.calls.values())

    def _texts(self, rows):
        """(call text per row, style per row) of one sample's rows."""
        gt = rows["GT"]
        style = np.zeros(rows.shape[0], dtype=np.uint8)
        style[(gt == GT_BLANK)] = 1
        if self.classic:
            style[gt == GT_SKIPPED] = 2
        else:
            style[gt == GT_SKIPPED] = 1
        if self.native is not None:
            return self.native.format_calls(rows, self.order, style), style
        out = []
        blank = dict(_BLANK_ITEMS)
        for i in range(rows.shape[0]):
            row = rows[i]
            if style[i] == 2:
                out.append(":".join("./." if k == "GT" else "." for k in self.order))
                continue
            if style[i] == 1:
                out.append(":".join(str(blank[k]) for k in self.order))
                continue
            vals = dict(_count_items(row))
            vals["GL"] = ",".join("%.0f" % x for x in row["GL"])
            g = int(row["GT"])
            if g == GT_UNDERFLOW:
                vals.update(GQ=".", SQ=".", GT="./.")
            else:
                vals.update(GQ=int(row["GQ"]), SQ="%0.2f" % float(row["SQ"]), GT=GT_TEXT[g])
            out.append(":".join(str(vals[k]) for k in self.order))
        return out, style

    def lines(self, plan, rows_by_sample):
        """VCF lines of the chunk, in plan order."""
        n = len(plan.breakpoints)
        texts, styles = [], []
        for name in self.samples:
            t, s = self._texts(rows_by_sample[name])
            texts.append(t)
            styles.append(s)
        # QUAL: each sample in order adds its SQ when called; in classic a no-evidence sample resets it to 0
        base = np.zeros(n, dtype=np.float64)
        for kind, rec, mate, idx in plan.entries:
            if kind == "site":
                base[idx] = rec.qual
        qual = base.copy()
        for name in self.samples:
            rows = rows_by_sample[name]
            gt = rows["GT"]
            qual = np.where(gt >= 0, qual + rows["SQ"], qual)
            if self.classic:
                qual = np.where(gt == GT_BLANK, 0.0, qual)
        if self.native is not None:
            qtext = self.native.format_quals(qual)
        else:
            qtext = ["%0.2f" % v for v in qual.tolist()]
        all_dots = None
        if self.classic:                                    # every sample skipped: the record only ever got GT
            all_dots = np.ones(n, dtype=bool)
            for s in styles:
                all_dots &= s == 2
        out = []
        for kind, rec, mate, idx in plan.entries:
            if kind == "site" and self.eligible(rec) and (mate is None or self.eligible(mate)):
                if all_dots is not None and all_dots[idx]:
                    fmt, calls = "GT", "\t".join("./." for _ in self.samples)
                else:
                    fmt, calls = self.fmt_full, "\t".join(t[idx] for t in texts)
                tail = "\t" + fmt + "\t" + calls
                out.append("\t".join((rec.chrom, str(rec.pos), rec.var_id, rec.ref, rec.alt, qtext[idx], rec.filter,
                                      rec.info_string())) + tail)
                if mate is not None:                        # BND mates share one genotype (classic.py:516-521)
                    out.append("\t".join((mate.chrom, str(mate.pos), mate.var_id, mate.ref, mate.alt, qtext[idx],
                                          mate.filter, mate.info_string())) + tail)
                continue
            if kind == "site":
                for name in self.samples:
                    apply_row(rec, name, rows_by_sample[name][idx], classic=self.classic)
            out.append(rec.render())
            if mate is not None:
                mate.adopt_calls(rec)
                out.append(mate.render())
        return out


# ---------------------------------------------------------------------------------------------
# the pipeline both entry points run
def walk_records(lines, header, sum_quals, max_ci_dist, batch_size, open_bnds):
    """Yield SitePlans of up to `batch_size` breakpoints from VCF record lines (reference classic.py:212-258:
    pass-through of records without a usable SVTYPE, BND mates paired through MATEID)."""
    from . import vcf
    plan = SitePlan()
    for line in lines:
        if line.startswith("#") or not line.strip():
            continue
        rec = vcf.VcfRecord(line.rstrip().split("\t"), header)
        if not sum_quals:
            rec.qual = 0
        if not rec.has_svtype():
            warn("Warning: SVTYPE missing at variant %s. Skipping.\n" % rec.var_id)
            plan.passthrough(rec)
            continue
        svtype = rec.svtype()
        if svtype not in ("BND", "DEL", "DUP", "INV"):
            warn("Warning: Unsupported SVTYPE at variant %s (%s). Skipping.\n" % (rec.var_id, svtype))
            plan.passthrough(rec)
            continue
        if svtype == "BND":
            mate_id = rec.info["MATEID"]
            if mate_id not in open_bnds:
                open_bnds[rec.var_id] = rec
                continue
            first = open_bnds.pop(mate_id)
            plan.site(first, rec, vcf.bnd_breakpoint(first, rec, max_ci_dist))
        else:
            plan.site(rec, None, vcf.simple_breakpoint(rec, max_ci_dist))
        if len(plan.breakpoints) >= batch_size:
            yield plan
            plan = SitePlan()
    if plan.entries:
        yield plan


def run_pipeline(samples, plans, write_lines, gather, mode, classic, min_aligned, split_weight, disc_weight, max_reads,
                 header, threads=0):
    """Pack -> score -> write every plan; packing of plan k + 1 overlaps scoring / writing of plan k."""
    assoc = ev.ASSOC_CLASSIC if classic else ev.ASSOC_SSO
    writer = ChunkWriter(header, [s.name for s in samples], classic)
    allocs = pinned_allocs()
    one = len(samples) == 1

    def pack(plan, k):
        out = []
        for s in samples:
            wide = pack_sample(s, plan, gather, min_aligned, mode=mode, max_reads=max_reads, threads=threads)
            # one sample: its compact rows are written straight into this chunk's pinned arena
            out.append(to_compact(wide, min_aligned, alloc=allocs[k & 1] if one else None, threads=threads))
        return out

    def finish(plan, packed, k):
        n = len(plan.breakpoints)
        rows_by_sample = {}
        if n:
            merged = merge_samples(packed, alloc=allocs[k & 1])
            rows = score(merged, min_aligned=min_aligned, split_slop=SPLIT_SLOP, split_weight=split_weight,
                         disc_weight=disc_weight, assoc_mode=assoc)
            if writer.native is not None:                   # SQ (and so QUAL) from the bit-exact GL with the host libm
                rows = writer.native.host_sq(np.ascontiguousarray(rows), threads=threads)
            for k, s in enumerate(samples):
                rows_by_sample[s.name] = rows[k * n:(k + 1) * n]
        else:
            for s in samples:
                rows_by_sample[s.name] = np.zeros(0, dtype=ev.OUT_DTYPE)
        write_lines(writer.lines(plan, rows_by_sample))

    it = iter(plans)
    with ThreadPoolExecutor(max_workers=1) as pool:
        try:
            plan = next(it)
        except StopIteration:
            return
        k = 0
        fut = pool.submit(pack, plan, k)
        while True:
            packed = fut.result()
            try:
                nxt = next(it)
                fut = pool.submit(pack, nxt, k + 1)
            except StopIteration:
                nxt = None
            finish(plan, packed, k)
            if nxt is None:
                break
            plan, k = nxt, k + 1
