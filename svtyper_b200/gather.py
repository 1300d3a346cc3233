"""Read gathering and split-read candidate QC: the host-side producers of evidence rows.

Host-side feature extraction: BAM records around both breakends are
grouped into fragments by query name, each primary read is tested as a split-read candidate,
and `evidence.BatchPacker` turns the fragments into the rows the scoring kernel streams.
Nothing here scores evidence.  Behaviour follows the reference's `gather_reads`
(svtyper/classic.py:54-100, svtyper/singlesample.py:139-205), `SamFragment.add_read`
(svtyper/parsers.py:748-768) and `SplitRead.is_valid` (svtyper/parsers.py:959-1058).
"""
from __future__ import annotations

import re

_CIGAR_OPS = "MIDNSHP=X"
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")

MIN_NON_OVERLAP = 20      # reference parsers.py:960-962
MIN_INDEL = 50
MAX_UNMAPPED_BASES = 50


def cigar_from_string(text):
    """'36M2D64M' -> [(0, 36), (2, 2), (0, 64)]  (BAM op codes)."""
    return [(_CIGAR_OPS.index(op), int(n)) for n, op in _CIGAR_RE.findall(text)]


def reference_end_of(start, cigar):
    """Coordinate just past the last aligned base: start + sum(M, D, N, =, X)."""
    return start + sum(n for op, n in cigar if op in (0, 2, 3, 7, 8))


def query_span(cigar, is_reverse):
    """(query_start, query_end, query_length) of the aligned part of the ORIGINAL read
    (leading clip skipped; reverse-strand alignments are read right to left)."""
    ops = cigar[::-1] if is_reverse else cigar
    start = end = length = 0
    for i, (op, n) in enumerate(ops):
        if op in (4, 5):
            if i == 0:
                start += n
                end += n
            length += n
        elif op in (0, 1, 7, 8):
            end += n
            length += n
    return start, end, length


def _is_clip(op):
    return op == 4 or op == 5


def _left_clipped(cigar):
    """Is the longer clip on the reference-left side of the alignment?"""
    (lop, llen), (rop, rlen) = cigar[0], cigar[-1]
    left, right = _is_clip(lop), _is_clip(rop)
    return (left and not right) or (left and right and llen > rlen)


class Piece(object):
    """One aligned piece of a (possibly chimeric) read."""
    __slots__ = ("chrom", "reference_start", "reference_end", "is_reverse", "mapping_quality", "cigar",
                 "qstart", "qend", "qlen")

    def __init__(self, chrom, start, end, is_reverse, cigar, mapq):
        self.chrom, self.reference_start, self.reference_end = chrom, start, end
        self.is_reverse, self.cigar, self.mapping_quality = is_reverse, cigar, mapq
        self.qstart, self.qend, self.qlen = query_span(cigar, is_reverse)

    def start_diagonal(self):
        clip = (self.qlen - self.qend) if self.is_reverse else self.qstart
        return self.reference_start - clip

    def end_diagonal(self):
        aligned = (self.qlen - self.qstart) if self.is_reverse else self.qend
        return self.reference_end - aligned


class SplitCandidate(object):
    __slots__ = ("query_left", "query_right", "is_soft_clip")

    def __init__(self, left, right, soft):
        self.query_left, self.query_right, self.is_soft_clip = left, right, soft


def split_candidate(read):
    """The read's split / soft-clip candidate, or None if it fails the QC rules."""
    cigar = read.cigar
    own = Piece(read.reference_name, read.reference_start, read.reference_end, read.is_reverse, cigar,
                read.mapping_quality)
    if not read.has_tag("SA"):
        first, last = _is_clip(cigar[0][0]), _is_clip(cigar[-1][0])
        if not (first or last):
            return None
        longest = max(cigar[0][1] * first, cigar[-1][1] * last)
        if longest > 0 and (read.query_length - read.query_alignment_length) <= MAX_UNMAPPED_BASES:
            ghost = Piece(None, 1, 1, read.is_reverse, cigar, 0)
            if _left_clipped(cigar):
                return SplitCandidate(ghost, own, True)
            return SplitCandidate(own, ghost, True)
        return None

    entries = read.get_tag("SA").rstrip(";").split(";")
    if len(entries) > 1:
        return None
    chrom, pos, strand, cig, mapq = entries[0].split(",")[:5]
    mate_pos = int(pos) - 1                      # SA is one-based
    mate_cigar = cigar_from_string(cig)
    mate = Piece(chrom, mate_pos, reference_end_of(mate_pos, mate_cigar), strand == "-", mate_cigar, int(mapq))
    if read.reference_name == chrom:
        left, right = (mate, own) if read.reference_start > mate_pos else (own, mate)
    else:
        left, right = (mate, own) if _left_clipped(cigar) else (own, mate)

    # the two pieces must each cover enough of the read that the other does not
    overlap = max(0, 1 + min(left.qend, right.qend) - max(left.qstart, right.qstart))
    if min(1 + left.qend - left.qstart - overlap, 1 + right.qend - right.qstart - overlap) < MIN_NON_OVERLAP:
        return None
    if left.chrom == right.chrom and left.is_reverse == right.is_reverse:
        if left.is_reverse:
            ins = right.end_diagonal() - left.start_diagonal()
        else:
            ins = left.end_diagonal() - right.start_diagonal()
        if abs(ins) < MIN_INDEL:
            return None
        desert = right.qstart - left.qend - 1
        if desert > 0 and desert - max(0, ins) > MAX_UNMAPPED_BASES:
            return None
    return SplitCandidate(left, right, False)


class Fragment(object):
    """All alignments of one molecule seen so far (what the packer consumes)."""
    __slots__ = ("lib_index", "primary_reads", "split_reads", "_seen")

    def __init__(self, lib_index):
        self.lib_index = lib_index
        self.primary_reads, self.split_reads = [], []
        self._seen = set()

    def add(self, read):
        key = (read.query_name, read.flag)
        if key in self._seen:
            return
        self._seen.add(key)
        if read.is_supplementary or read.is_secondary:
            return
        self.primary_reads.append(read)
        cand = split_candidate(read)
        if cand is not None:
            self.split_reads.append(cand)


def _collect(reads, sample, fragments, limit=None):
    """Add usable reads to `fragments`; True if more than `limit` records were seen
    (classic semantics: the position in the fetch, not the number kept; classic.py:79-91)."""
    for i, read in enumerate(reads):
        if read.is_unmapped or read.is_duplicate:
            continue
        lib = sample.rg_to_lib[read.get_tag("RG")]
        if lib not in sample.active:
            continue
        if limit is not None and i > limit:
            return True
        frag = fragments.get(read.query_name)
        if frag is None:
            frag = fragments[read.query_name] = Fragment(lib)
        frag.add(read)
    return False


def gather_classic(sample, breakpoint, z, max_reads):
    """({qname: Fragment}, too_many) the way classic.sv_genotype gathers (classic.py:54-100)."""
    bam = sample.bam
    flank = sample.fetch_flank(z)
    fragments = {}
    for side in ("A", "B"):
        end = breakpoint[side]
        length = bam.lengths[bam.gettid(end["chrom"])]
        reads = bam.fetch(end["chrom"], max(end["pos"] + end["ci"][0] - flank, 0),
                          min(end["pos"] + end["ci"][1] + flank, length))
        if _collect(reads, sample, fragments, max_reads):
            return {}, True
    return fragments, False


def breakpoint_regions(sample, breakpoint, z):
    """((chrom, left, right) x 2) fetch windows (singlesample.py:139-157)."""
    bam = sample.bam
    flank = sample.fetch_flank(z)
    out = []
    for side in ("A", "B"):
        end = breakpoint[side]
        length = bam.lengths[bam.gettid(end["chrom"])]
        out.append((end["chrom"], int(max(end["pos"] + end["ci"][0] - flank, 0)),
                    int(min(end["pos"] + end["ci"][1] + flank, length))))
    return tuple(out)


def gather_sso(sample, breakpoint, z, max_reads, bam=None):
    """({qname: Fragment}, over_threshold) the way svtyper-sso gathers (singlesample.py:168-205):
    both windows are counted first, then fetched."""
    bam = sample.bam if bam is None else bam
    regions = breakpoint_regions(sample, breakpoint, z)
    if max_reads is not None:
        for chrom, left, right in regions:
            if bam.count(chrom, start=left, stop=right, read_callback="all") > max_reads:
                return {}, True
    fragments = {}
    for chrom, left, right in regions:
        _collect(bam.fetch(chrom, start=left, stop=right), sample, fragments)
    return fragments, False
