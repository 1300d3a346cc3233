"""Minimal BAM/BAI reader exposing the subset of the pysam API the genotyper uses.

pysam (htslib) is the reference's third-party I/O layer (reference
`setup.py:24`); it is absent from this image.  Read gathering stays on the CPU
by design (BASELINE.json north_star), so the drop-in needs *some* BAM access:
when pysam is importable it is used, otherwise this stdlib (zlib + struct)
reader stands in.  It honours exactly the pysam semantics the reference's
gather/scoring code relies on:

  fetch(chrom, start, stop)   file-order records with pos < stop and end > start
  count(..., read_callback='all')  skips flags 0x4|0x100|0x200|0x400
  get_overlap(s, e)           overlap with M/=/X blocks, advancing over M/D/N/=/X
  reference_end               pos + sum(M, D, N, =, X)
  query_alignment_length      sum(M, I, =, X)
  query_length                l_seq (0 when sequences are stripped)

Region queries use the .bai index (binning + linear index) so a query only
inflates the BGZF blocks it needs.
"""
from __future__ import annotations

import os
import struct
import zlib
from collections import OrderedDict

_CIGAR_CONSUMES_REF = (True, False, True, True, False, False, False, True, True)   # MIDNSHP=X
_CIGAR_IS_ALIGNED = (True, False, False, False, False, False, False, True, True)
_CIGAR_CONSUMES_QUERY_ALN = (True, True, False, False, False, False, False, True, True)

FUNMAP, FREVERSE, FMUNMAP, FMREVERSE = 0x4, 0x10, 0x8, 0x20
FSECONDARY, FQCFAIL, FDUP, FSUPPLEMENTARY = 0x100, 0x200, 0x400, 0x800


class BgzfFile(object):
    """Random-access reader over a BGZF file addressed by virtual offsets."""

    def __init__(self, path, cache_blocks=64):
        self._f = open(path, "rb")
        self._cache = OrderedDict()
        self._cache_blocks = cache_blocks
        self._coff = 0          # compressed offset of current block
        self._data = b""        # inflated current block
        self._csize = 0         # compressed size of current block
        self._uoff = 0          # offset within inflated block
        self._load(0)

    def close(self):
        self._f.close()

    def _load(self, coff):
        hit = self._cache.get(coff)
        if hit is not None:
            self._cache.move_to_end(coff)
            self._coff, (self._data, self._csize) = coff, hit
            return
        self._f.seek(coff)
        hdr = self._f.read(18)
        if len(hdr) < 18:
            self._coff, self._data, self._csize = coff, b"", 0
            return
        if hdr[:4] != b"\x1f\x8b\x08\x04":
            raise IOError("not a BGZF block at offset %d" % coff)
        xlen = struct.unpack_from("<H", hdr, 10)[0]
        extra = hdr[12:18] + self._f.read(xlen - 6)
        bsize = None
        p = 0
        while p + 4 <= len(extra):
            si1, si2, slen = extra[p], extra[p + 1], struct.unpack_from("<H", extra, p + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", extra, p + 4)[0]
            p += 4 + slen
        if bsize is None:
            raise IOError("BGZF block without BC subfield")
        csize = bsize + 1
        payload = self._f.read(csize - 12 - xlen - 8)
        self._f.read(8)
        data = zlib.decompress(payload, -15) if payload else b""
        self._cache[coff] = (data, csize)
        if len(self._cache) > self._cache_blocks:
            self._cache.popitem(last=False)
        self._coff, self._data, self._csize = coff, data, csize

    def seek(self, voffset):
        self._load(voffset >> 16)
        self._uoff = voffset & 0xFFFF

    def tell(self):
        if self._uoff >= len(self._data) and self._csize:
            return (self._coff + self._csize) << 16
        return (self._coff << 16) | self._uoff

    def read(self, n):
        out = []
        while n > 0:
            avail = len(self._data) - self._uoff
            if avail <= 0:
                if self._csize == 0:
                    break
                self._load(self._coff + self._csize)
                self._uoff = 0
                if self._csize == 0:
                    break
                continue
            take = avail if avail < n else n
            out.append(self._data[self._uoff:self._uoff + take])
            self._uoff += take
            n -= take
        return b"".join(out)


class AlignedSegment(object):
    """One BAM record; attribute names follow pysam.AlignedSegment."""

    __slots__ = ("_hdr", "reference_id", "reference_start", "mapping_quality", "flag",
                 "next_reference_id", "next_reference_start", "template_length",
                 "query_name", "cigar", "query_length", "_tagbytes", "_tags",
                 "_ref_end", "query_sequence")

    def __init__(self, hdr, buf):
        (self.reference_id, self.reference_start, l_name, self.mapping_quality, _bin,
         n_cig, self.flag, l_seq, self.next_reference_id, self.next_reference_start,
         self.template_length) = struct.unpack_from("<iiBBHHHiiii", buf, 0)
        self._hdr = hdr
        p = 32
        self.query_name = buf[p:p + l_name - 1].decode("ascii")
        p += l_name
        raw = struct.unpack_from("<%dI" % n_cig, buf, p)
        self.cigar = [(v & 0xF, v >> 4) for v in raw]
        p += 4 * n_cig
        self.query_length = l_seq
        p += (l_seq + 1) // 2 + l_seq
        self._tagbytes = buf[p:]
        self._tags = None
        self._ref_end = -1
        self.query_sequence = None

    # ---- flags -----------------------------------------------------
    @property
    def is_unmapped(self): return bool(self.flag & FUNMAP)
    @property
    def mate_is_unmapped(self): return bool(self.flag & FMUNMAP)
    @property
    def is_reverse(self): return bool(self.flag & FREVERSE)
    @property
    def mate_is_reverse(self): return bool(self.flag & FMREVERSE)
    @property
    def is_secondary(self): return bool(self.flag & FSECONDARY)
    @property
    def is_supplementary(self): return bool(self.flag & FSUPPLEMENTARY)
    @property
    def is_duplicate(self): return bool(self.flag & FDUP)
    @property
    def is_qcfail(self): return bool(self.flag & FQCFAIL)

    # ---- coordinates ----------------------------------------------
    @property
    def pos(self): return self.reference_start

    @property
    def reference_name(self):
        return self._hdr.get_reference_name(self.reference_id)

    @property
    def reference_end(self):
        if self._ref_end == -1:
            if self.is_unmapped or not self.cigar:
                self._ref_end = None
            else:
                e = self.reference_start
                for op, n in self.cigar:
                    if _CIGAR_CONSUMES_REF[op]:
                        e += n
                self._ref_end = e
        return self._ref_end

    @property
    def cigartuples(self): return self.cigar

    @property
    def query_alignment_length(self):
        return sum(n for op, n in self.cigar if _CIGAR_CONSUMES_QUERY_ALN[op])

    def infer_query_length(self):
        return sum(n for op, n in self.cigar if op in (0, 1, 4, 7, 8)) or None

    def get_overlap(self, start, end):
        pos = self.reference_start
        overlap = 0
        for op, n in self.cigar:
            if _CIGAR_IS_ALIGNED[op]:
                o = min(pos + n, end) - max(pos, start)
                if o > 0:
                    overlap += o
            if _CIGAR_CONSUMES_REF[op]:
                pos += n
        return overlap

    def get_blocks(self):
        """Aligned (M/=/X) reference blocks as [start, end) pairs."""
        pos = self.reference_start
        blocks = []
        for op, n in self.cigar:
            if _CIGAR_IS_ALIGNED[op]:
                blocks.append((pos, pos + n))
            if _CIGAR_CONSUMES_REF[op]:
                pos += n
        return blocks

    # ---- tags -------------------------------------------------------
    def _parse_tags(self):
        tags = {}
        b = self._tagbytes
        p, n = 0, len(b)
        while p + 3 <= n:
            key = b[p:p + 2].decode("ascii")
            t = chr(b[p + 2])
            p += 3
            if t == "A":
                tags[key] = chr(b[p]); p += 1
            elif t == "c":
                tags[key] = struct.unpack_from("<b", b, p)[0]; p += 1
            elif t == "C":
                tags[key] = b[p]; p += 1
            elif t == "s":
                tags[key] = struct.unpack_from("<h", b, p)[0]; p += 2
            elif t == "S":
                tags[key] = struct.unpack_from("<H", b, p)[0]; p += 2
            elif t == "i":
                tags[key] = struct.unpack_from("<i", b, p)[0]; p += 4
            elif t == "I":
                tags[key] = struct.unpack_from("<I", b, p)[0]; p += 4
            elif t == "f":
                tags[key] = struct.unpack_from("<f", b, p)[0]; p += 4
            elif t in "ZH":
                e = b.index(b"\x00", p)
                tags[key] = b[p:e].decode("ascii"); p = e + 1
            elif t == "B":
                sub = chr(b[p]); cnt = struct.unpack_from("<I", b, p + 1)[0]; p += 5
                fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
                sz = struct.calcsize(fmt)
                tags[key] = list(struct.unpack_from("<%d%s" % (cnt, fmt), b, p)); p += cnt * sz
            else:
                raise ValueError("bad BAM tag type %r" % t)
        self._tags = tags

    def has_tag(self, key):
        if self._tags is None:
            self._parse_tags()
        return key in self._tags

    def get_tag(self, key):
        if self._tags is None:
            self._parse_tags()
        return self._tags[key]

    def set_tag(self, key, value, value_type=None):
        if self._tags is None:
            self._parse_tags()
        self._tags[key] = value


def _reg2bins(beg, end):
    end -= 1
    bins = [0]
    for shift, base in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(base + (beg >> shift), base + (end >> shift) + 1))
    return bins


class _BaiIndex(object):
    def __init__(self, path):
        with open(path, "rb") as f:
            b = f.read()
        if b[:4] != b"BAI\x01":
            raise IOError("bad BAI magic in %s" % path)
        p = 4
        n_ref = struct.unpack_from("<i", b, p)[0]; p += 4
        self.bins, self.linear = [], []
        self.mapped = self.unmapped = 0
        for _ in range(n_ref):
            n_bin = struct.unpack_from("<i", b, p)[0]; p += 4
            d = {}
            for _ in range(n_bin):
                bin_id, n_chunk = struct.unpack_from("<Ii", b, p); p += 8
                chunks = [struct.unpack_from("<QQ", b, p + 16 * k) for k in range(n_chunk)]
                p += 16 * n_chunk
                if bin_id == 37450:      # htslib pseudo-bin: per-reference counts
                    if n_chunk == 2:
                        self.mapped += chunks[1][0]
                        self.unmapped += chunks[1][1]
                else:
                    d[bin_id] = chunks
            n_intv = struct.unpack_from("<i", b, p)[0]; p += 4
            self.linear.append(struct.unpack_from("<%dQ" % n_intv, b, p)); p += 8 * n_intv
            self.bins.append(d)
        if p + 8 <= len(b):
            self.unmapped += struct.unpack_from("<Q", b, p)[0]

    def chunks(self, tid, beg, end):
        if tid < 0 or tid >= len(self.bins):
            return []
        lin = self.linear[tid]
        w = beg >> 14
        min_off = lin[w] if w < len(lin) else (lin[-1] if lin else 0)
        out = []
        d = self.bins[tid]
        for b_ in _reg2bins(beg, end):
            for cb, ce in d.get(b_, ()):
                if ce > min_off:
                    out.append((cb, ce))
        out.sort()
        merged = []
        for cb, ce in out:
            if merged and cb <= merged[-1][1]:
                if ce > merged[-1][1]:
                    merged[-1][1] = ce
            else:
                merged.append([cb, ce])
        return merged


class AlignmentFile(object):
    """pysam.AlignmentFile look-alike for coordinate-sorted, indexed BAM."""

    def __init__(self, filename, mode="rb", **_ignored):
        if mode not in ("rb", "r"):
            raise ValueError("bamio.AlignmentFile only reads BAM (mode %r)" % mode)
        self.filename = str(filename)
        self._bgzf = BgzfFile(self.filename)
        if self._bgzf.read(4) != b"BAM\x01":
            raise IOError("%s is not a BAM file" % filename)
        l_text = struct.unpack("<i", self._bgzf.read(4))[0]
        self.text = self._bgzf.read(l_text).split(b"\x00")[0].decode("ascii", "replace")
        n_ref = struct.unpack("<i", self._bgzf.read(4))[0]
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            l_name = struct.unpack("<i", self._bgzf.read(4))[0]
            self.references.append(self._bgzf.read(l_name)[:-1].decode("ascii"))
            self.lengths.append(struct.unpack("<i", self._bgzf.read(4))[0])
        self.references, self.lengths = tuple(self.references), tuple(self.lengths)
        self._tid = {n: i for i, n in enumerate(self.references)}
        self._first_record = self._bgzf.tell()
        self.header = self._parse_header(self.text)
        self._index = None
        for cand in (self.filename + ".bai", os.path.splitext(self.filename)[0] + ".bai"):
            if os.path.exists(cand):
                self._index = _BaiIndex(cand)
                break

    @staticmethod
    def _parse_header(text):
        hdr = {}
        for line in text.split("\n"):
            if not line.startswith("@") or line.startswith("@CO"):
                continue
            f = line.rstrip("\r").split("\t")
            rec = {}
            for kv in f[1:]:
                if len(kv) >= 3 and kv[2] == ":":
                    rec[kv[:2]] = kv[3:]
            key = f[0][1:]
            if key == "HD":
                hdr[key] = rec
            else:
                hdr.setdefault(key, []).append(rec)
        return hdr

    # ---- pysam-like accessors ----------------------------------------
    @property
    def nreferences(self): return len(self.references)
    @property
    def mapped(self):
        return self._index.mapped if self._index else 0
    @property
    def unmapped(self):
        return self._index.unmapped if self._index else 0

    def gettid(self, name): return self._tid.get(name, -1)
    get_tid = gettid

    def get_reference_name(self, tid):
        return self.references[tid] if 0 <= tid < len(self.references) else None
    getrname = get_reference_name

    def close(self):
        self._bgzf.close()

    # ---- iteration ------------------------------------------------------
    def _next(self):
        szb = self._bgzf.read(4)
        if len(szb) < 4:
            return None
        buf = self._bgzf.read(struct.unpack("<i", szb)[0])
        return AlignedSegment(self, buf)

    def fetch(self, contig=None, start=None, stop=None, reference=None, end=None, until_eof=False):
        if contig is None:
            contig = reference
        if stop is None:
            stop = end
        if contig is None:
            self._bgzf.seek(self._first_record)
            while True:
                r = self._next()
                if r is None:
                    return
                if r.reference_id < 0 and not until_eof:
                    return
                yield r
        tid = self.gettid(contig)
        if tid < 0:
            raise ValueError("invalid contig %r" % (contig,))
        if self._index is None:
            raise ValueError("fetch on a region requires a .bai index")
        beg = 0 if start is None else max(0, int(start))
        fin = self.lengths[tid] if stop is None else int(stop)
        if fin <= beg:
            return
        for cb, ce in self._index.chunks(tid, beg, fin):
            self._bgzf.seek(cb)
            while self._bgzf.tell() < ce:
                r = self._next()
                if r is None or r.reference_id != tid or r.reference_start >= fin:
                    break
                rend = r.reference_end
                if rend is None or rend <= r.reference_start:
                    rend = r.reference_start + 1
                if rend > beg:
                    yield r
            else:
                continue
            # a record at/after `fin` (or another contig) ends the whole query
            if r is not None and (r.reference_id != tid or r.reference_start >= fin):
                return

    def count(self, contig=None, start=None, stop=None, read_callback="nofilter", **kw):
        n = 0
        for r in self.fetch(contig, start, stop):
            if read_callback == "all":
                if r.flag & (FUNMAP | FSECONDARY | FQCFAIL | FDUP):
                    continue
            elif callable(read_callback):
                if not read_callback(r):
                    continue
            n += 1
        return n
