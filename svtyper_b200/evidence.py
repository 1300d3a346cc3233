"""Evidence schema: the HBM layout the scoring kernel streams, and the host packer.

A *batch* is everything one kernel launch needs to genotype `n_sites`
breakpoints of one sample:

  site rows      int32 [n_sites, 16]  (64 B)   one per SV breakpoint pair
  fragment rows  int32 [n_frag,  8]   (32 B)   one per read pair, grouped per site,
                                               inside a site in sorted(query_name)
                                               order (reference singlesample.py:364,
                                               classic.py:296 -- the fp64 sums are
                                               order-sensitive)
  split rows     int32 [n_split, 8]   (32 B)   one per valid split/soft-clip candidate
  library table  f64 [n_lib, 4] + int32 [n_lib, 4] + uint32 hist[]  (insert-size PDF
                                               counts; reference parsers.py:579-583)
  LUTs           f64 pm[256], f64 log10[n_log], f64 consts[32]   built with CPython
                                               `math` (no device
                                               transcendentals on the integer path)
  output rows    80 B  = f64 GL[3], f64 SQ, int32 x 12

Word layouts (little-endian int32 words):

site row (SITE_WORDS = 16)
  0 posA   1 posB        breakend positions AFTER the reverse-strand +1
                         (reference parsers.py:205-207, classic.py:276-277)
  2 ciA0 3 ciA1 4 ciB0 5 ciB1
  6 tidA   7 tidB        BAM reference ids of the breakend chromosomes (-1 unknown)
  8 var_length           DEL only: posB - posA BEFORE the +1 (parsers.py:182)
  9 meta                 bits 0-1 svtype (0 DEL 1 DUP 2 INV 3 BND), bit 2 o1_is_reverse,
                         bit 3 o2_is_reverse, bit 4 SKIP (too-many-reads row)
  10,11 frag_off (int64) 12 n_frag
  13,14 split_off (int64) 15 n_split

fragment row (FRAG_WORDS = 8): the two primary alignments of one SamFragment
  0 a_start 1 a_end      readA = first primary seen (parsers.py:767-768)
  2 b_start 3 b_end      readB = second primary seen
  4 tidA    5 tidB
  6 mapqA | mapqB<<8 | lib<<16
  7 flags (F_* below)

split row (SPLIT_WORDS = 8): SplitRead.query_left / query_right (parsers.py:1017-1028)
  0 l_tid 1 l_start 2 l_end 3 r_tid 4 r_start 5 r_end
  6 mapqL | mapqR<<8 | flags<<16   (S_* below)
  7 fragment ordinal inside the site (informational)
"""
from __future__ import annotations

import math

import numpy as np

SITE_WORDS, FRAG_WORDS, SPLIT_WORDS = 16, 8, 8
SITE_BYTES, FRAG_BYTES, SPLIT_BYTES, OUT_BYTES = 64, 32, 32, 80

SV_DEL, SV_DUP, SV_INV, SV_BND = 0, 1, 2, 3
SVTYPE_CODE = {"DEL": SV_DEL, "DUP": SV_DUP, "INV": SV_INV, "BND": SV_BND}

SITE_O1_REV, SITE_O2_REV, SITE_SKIP = 1 << 2, 1 << 3, 1 << 4

F_HAS_A = 1 << 0      # slot A holds a primary alignment
F_HAS_B = 1 << 1      # slot B holds a primary alignment
F_REV_A = 1 << 2
F_REV_B = 1 << 3
F_PAIRED = 1 << 4     # num_primary == 2: slot A/B are readA/readB (parsers.py:827)
F_CONT = 1 << 5       # row continues the previous row's fragment (num_primary > 2)
F_EXTRA = 1 << 6      # row only carries extra aligned intervals for the NEXT main row
F_MULTI_A = 1 << 7    # slot A read has D/N gaps: take its ref-seq hit from EXTRA rows
F_MULTI_B = 1 << 8

S_SOFT_CLIP = 1 << 0  # SplitRead.is_soft_clip (parsers.py:983)
S_FIRST = 1 << 1      # first split of its fragment (sso per-fragment sub-totals)

TID_NONE = -2         # SplitPiece.chrom is None / not in the BAM header

GT_UNDERFLOW, GT_BLANK, GT_SKIPPED = -1, -2, -3

ASSOC_SSO, ASSOC_CLASSIC = 0, 1

OUT_DTYPE = np.dtype([
    ("GL", "<f8", (3,)), ("SQ", "<f8"),
    ("GT", "<i4"), ("GQ", "<i4"), ("DP", "<i4"), ("RO", "<i4"), ("AO", "<i4"),
    ("QR", "<i4"), ("QA", "<i4"), ("RS", "<i4"), ("AS", "<i4"), ("ASC", "<i4"),
    ("RP", "<i4"), ("AP", "<i4"),
])
assert OUT_DTYPE.itemsize == OUT_BYTES

# consts[] slots
C_NONDUP_ALT, C_NONDUP_REF, C_DUP_ALT, C_DUP_REF = 0, 3, 6, 9
C_CONC_PRIOR, C_DISC_PRIOR, C_POW10_MIN_X = 12, 13, 14
N_CONSTS = 32


def _pow10_min_x() -> float:
    """Smallest double x with 10.0**x > 0 under this interpreter's libm."""
    lo, hi = -400.0, -300.0          # 10**lo == 0, 10**hi > 0
    assert 10.0 ** lo == 0.0 and 10.0 ** hi > 0.0
    while True:
        mid = lo + (hi - lo) / 2
        if mid <= lo or mid >= hi:
            break
        if 10.0 ** mid > 0.0:
            hi = mid
        else:
            lo = mid
    while 10.0 ** math.nextafter(hi, -math.inf) > 0.0:
        hi = math.nextafter(hi, -math.inf)
    return hi


def build_luts(n_log: int):
    """Host look-up tables built with CPython `math` (same calls as the reference).

    pm[q]      = 1 - 10 ** (-q / 10.0)            reference utils.py:74-75
    logt[n]    = math.log(n, 10)                  reference statistics.py:16-17
    consts     = math.log(p, 10), math.log(1 - p, 10) for the two prior sets
                 (statistics.py:25-35), conc/disc priors (parsers.py:863-864)
    """
    pm = np.array([1 - 10 ** (-q / 10.0) for q in range(256)], dtype=np.float64)
    logt = np.zeros(max(2, int(n_log)), dtype=np.float64)
    for n in range(1, logt.size):
        logt[n] = math.log(n, 10)
    consts = np.zeros(N_CONSTS, dtype=np.float64)
    for base, p_alt in ((C_NONDUP_ALT, [1e-3, 0.5, 0.9]), (C_DUP_ALT, [1e-2, 0.2, 1 / 3.0])):
        for g, p in enumerate(p_alt):
            consts[base + g] = math.log(p, 10)
            consts[base + 3 + g] = math.log(1 - p, 10)
    disc_prior = 0.05
    consts[C_CONC_PRIOR] = 1 - disc_prior
    consts[C_DISC_PRIOR] = disc_prior
    consts[C_POW10_MIN_X] = _pow10_min_x()
    return pm, logt, consts


class LibraryTable(object):
    """Per-library constants + insert-size histogram counts, flattened for the device.

    lib_f64[l] = (flank = mean + sd*3, two_sd = 2*sd, N = sum(hist), mean)
    lib_i32[l] = (hist_off, hist_len, nondel_L, 0)
        nondel_L: the non-DEL `var_length = mean + sd*3` (parsers.py:874-875) when that
        float is integral-valued and >= 0 (then it hits the Counter as a key), else -1
        (float keys that are not integral never match: SURVEY.md H4).
    hist[hist_off + i] = count of insert size i (0 <= i < hist_len)
    """

    def __init__(self, libs):
        """libs: iterable of (mean, sd, {insert_size: count})."""
        f64, i32, chunks, off = [], [], [], 0
        libs = list(libs)
        self.sources = libs              # kept for host-side consumers (fetch flank, tests)
        for mean, sd, hist in libs:
            mean, sd = float(mean), float(sd)
            flank = mean + sd * 3
            keys = [int(k) for k in hist if int(k) >= 0 and int(hist[k]) != 0]
            hlen = (max(keys) + 1) if keys else 0
            h = np.zeros(hlen, dtype=np.uint32)
            total = 0
            for k, v in hist.items():
                total += int(v)
                if int(k) >= 0 and int(v) != 0:
                    h[int(k)] = int(v)
            nondel = int(flank) if (flank == math.floor(flank) and 0 <= flank < 2 ** 31) else -1
            f64.append((flank, 2 * sd, float(total), mean))
            i32.append((off, hlen, nondel, 0))
            chunks.append(h)
            off += hlen
        self.lib_f64 = np.array(f64, dtype=np.float64).reshape(-1, 4)
        self.lib_i32 = np.array(i32, dtype=np.int32).reshape(-1, 4)
        self.hist = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
        if self.hist.size == 0:
            self.hist = np.zeros(1, np.uint32)
        self.n_lib = len(f64)

    def max_flank(self):
        return float(self.lib_f64[:, 0].max()) if self.n_lib else 0.0


class EvidenceBatch(object):
    """Host-side (numpy) batch in the device layout."""

    def __init__(self, sites, frags, splits, libs: LibraryTable, order=None):
        self.sites = np.ascontiguousarray(sites, dtype=np.int32).reshape(-1, SITE_WORDS)
        self.frags = np.ascontiguousarray(frags, dtype=np.int32).reshape(-1, FRAG_WORDS)
        self.splits = np.ascontiguousarray(splits, dtype=np.int32).reshape(-1, SPLIT_WORDS)
        self.libs = libs
        self.order = None if order is None else np.ascontiguousarray(order, dtype=np.int32)

    @property
    def n_sites(self): return self.sites.shape[0]
    @property
    def n_frag(self): return self.frags.shape[0]
    @property
    def n_split(self): return self.splits.shape[0]

    def frag_counts(self): return self.sites[:, 12]
    def split_counts(self): return self.sites[:, 15]

    def algorithmic_bytes(self) -> int:
        """Bytes one pass must move: 64/site + 32/fragment row + 32/split row + 80/site out."""
        return (self.n_sites * (SITE_BYTES + OUT_BYTES) + self.n_frag * FRAG_BYTES
                + self.n_split * SPLIT_BYTES)

    def length_order(self):
        """Site permutation bucketing warps by work (descending fragment+split rows)."""
        work = self.sites[:, 12].astype(np.int64) + self.sites[:, 15]
        return np.argsort(-work, kind="stable").astype(np.int32)

    def log_table_size(self, split_weight=1.0, disc_weight=1.0) -> int:
        """Upper bound on QR+QA+1 for any site of the batch (sizes the log10 LUT)."""
        if self.n_sites == 0:
            return 2
        nf = int(self.sites[:, 12].max())
        ns = int(self.sites[:, 15].max())
        bound = abs(float(split_weight)) * (2 * nf + ns) + abs(float(disc_weight)) * 2 * nf
        return int(bound) + 8

    def slice_sites(self, lo, hi):
        """Contiguous site range [lo, hi) as an independent batch (multi-GPU shards)."""
        s = self.sites[lo:hi].copy()
        if s.shape[0] == 0:
            return EvidenceBatch(s, np.zeros((0, FRAG_WORDS), np.int32),
                                 np.zeros((0, SPLIT_WORDS), np.int32), self.libs)
        foff = s[:, 10:12].copy().view(np.int64).ravel()
        soff = s[:, 13:15].copy().view(np.int64).ravel()
        f0, f1 = int(foff[0]), int(foff[-1] + s[-1, 12])
        s0, s1 = int(soff[0]), int(soff[-1] + s[-1, 15])
        s[:, 10:12] = (foff - f0).astype(np.int64).view(np.int32).reshape(-1, 2)
        s[:, 13:15] = (soff - s0).astype(np.int64).view(np.int32).reshape(-1, 2)
        return EvidenceBatch(s, self.frags[f0:f1], self.splits[s0:s1], self.libs)


def _aligned_intervals(read):
    """Maximal gap-free reference intervals of a read (merged M/=/X blocks).

    `get_overlap(w0, w1) >= w1 - w0` (parsers.py:813) holds iff [w0, w1) lies inside one
    of these: insertions/clips do not break reference contiguity, D/N do.
    """
    if hasattr(read, "get_blocks"):
        blocks = read.get_blocks()
    else:
        blocks, pos = [], read.reference_start
        for op, n in read.cigar:
            if op in (0, 7, 8):
                blocks.append((pos, pos + n))
            if op in (0, 2, 3, 7, 8):
                pos += n
    merged = []
    for s, e in blocks:
        if merged and s == merged[-1][1]:
            merged[-1][1] = e
        elif e > s:
            merged.append([s, e])
    return merged


class BatchPacker(object):
    """Incrementally packs gathered fragments into an EvidenceBatch.

    `add_site(breakpoint, fragments, ...)` takes the reference-shaped breakpoint dict
    (parsers.py:190-203) and a {query_name: fragment} mapping whose values expose
    `lib_index`, `primary_reads` and `split_reads` (reference SamFragment fields,
    parsers.py:729-768, or this repo's own gather objects).
    """

    def __init__(self, tid_of, libs: LibraryTable):
        self._tid_of = tid_of
        self.libs = libs
        self._sites, self._frags, self._splits = [], [], []
        self._nf = self._ns = 0

    def _tid(self, chrom):
        if chrom is None:
            return TID_NONE
        t = self._tid_of(chrom)
        return TID_NONE if (t is None or t < 0) else t

    def add_site(self, breakpoint, fragments, skip=False, lib_index_of=None):
        A, B = breakpoint["A"], breakpoint["B"]
        meta = SVTYPE_CODE[breakpoint["svtype"]]
        if A["is_reverse"]:
            meta |= SITE_O1_REV
        if B["is_reverse"]:
            meta |= SITE_O2_REV
        if skip:
            meta |= SITE_SKIP
        f_off, s_off = self._nf, self._ns
        n_f = n_s = 0
        if not skip:
            for ordinal, qname in enumerate(sorted(fragments.keys())):
                frag = fragments[qname]
                lib = frag.lib_index if lib_index_of is None else lib_index_of(frag)
                rows, srows = self._pack_fragment(frag, lib, ordinal)
                self._frags.extend(rows)
                self._splits.extend(srows)
                n_f += len(rows)
                n_s += len(srows)
        self._nf += n_f
        self._ns += n_s
        tA = self._tid_of(A["chrom"])
        tB = self._tid_of(B["chrom"])
        self._sites.append((
            int(A["pos"]), int(B["pos"]), int(A["ci"][0]), int(A["ci"][1]),
            int(B["ci"][0]), int(B["ci"][1]),
            -1 if tA is None else int(tA), -1 if tB is None else int(tB),
            int(breakpoint.get("var_length", 0) or 0), meta,
            f_off & 0xFFFFFFFF, f_off >> 32, n_f, s_off & 0xFFFFFFFF, s_off >> 32, n_s))

    def _pack_fragment(self, frag, lib, ordinal):
        prim = list(frag.primary_reads)
        rows = []
        # rows of (slotA read or None, slotB read or None, flags)
        groups = []
        if len(prim) == 2:
            groups.append((prim[0], prim[1], F_PAIRED))
        else:
            for i, r in enumerate(prim):
                groups.append((r, None, F_CONT if i > 0 else 0))
        for ra, rb, fl in groups:
            iv_a = _aligned_intervals(ra) if ra is not None else []
            iv_b = _aligned_intervals(rb) if rb is not None else []
            multi_a = ra is not None and not (len(iv_a) == 1 and iv_a[0][0] == ra.reference_start
                                              and iv_a[0][1] == ra.reference_end)
            multi_b = rb is not None and not (len(iv_b) == 1 and iv_b[0][0] == rb.reference_start
                                              and iv_b[0][1] == rb.reference_end)
            xa = iv_a if multi_a else []
            xb = iv_b if multi_b else []
            for k in range(max(len(xa), len(xb))):
                efl = F_EXTRA | (fl & F_CONT)
                w = [0, 0, 0, 0, 0, 0, 0, 0]
                if k < len(xa):
                    efl |= F_HAS_A
                    w[0], w[1], w[4] = xa[k][0], xa[k][1], ra.reference_id
                if k < len(xb):
                    efl |= F_HAS_B
                    w[2], w[3], w[5] = xb[k][0], xb[k][1], rb.reference_id
                w[6] = (lib & 0xFFFF) << 16
                w[7] = efl
                rows.append(tuple(w))
            w = [0, 0, 0, 0, 0, 0, 0, 0]
            if ra is not None:
                fl |= F_HAS_A | (F_REV_A if ra.is_reverse else 0) | (F_MULTI_A if multi_a else 0)
                w[0], w[1], w[4] = ra.reference_start, ra.reference_end, ra.reference_id
                w[6] |= min(int(ra.mapping_quality), 255)
            if rb is not None:
                fl |= F_HAS_B | (F_REV_B if rb.is_reverse else 0) | (F_MULTI_B if multi_b else 0)
                w[2], w[3], w[5] = rb.reference_start, rb.reference_end, rb.reference_id
                w[6] |= min(int(rb.mapping_quality), 255) << 8
            w[6] |= (lib & 0xFFFF) << 16
            w[7] = fl
            rows.append(tuple(w))
        srows = []
        for i, sp in enumerate(frag.split_reads):
            L, R = sp.query_left, sp.query_right
            sfl = (S_SOFT_CLIP if sp.is_soft_clip else 0) | (S_FIRST if i == 0 else 0)
            meta = (min(int(L.mapping_quality), 255) | (min(int(R.mapping_quality), 255) << 8)
                    | (sfl << 16))
            srows.append((self._tid(L.chrom), int(L.reference_start), int(L.reference_end),
                          self._tid(R.chrom), int(R.reference_start), int(R.reference_end),
                          meta, ordinal))
        return rows, srows

    def finish(self, bucket=True) -> EvidenceBatch:
        sites = (np.array(self._sites, dtype=np.int64).astype(np.int32)
                 if self._sites else np.zeros((0, SITE_WORDS), np.int32))
        # words 10/13 were stored as unsigned low halves; int64->int32 cast wraps correctly
        frags = (np.array(self._frags, dtype=np.int64).astype(np.int32)
                 if self._frags else np.zeros((0, FRAG_WORDS), np.int32))
        splits = (np.array(self._splits, dtype=np.int64).astype(np.int32)
                  if self._splits else np.zeros((0, SPLIT_WORDS), np.int32))
        b = EvidenceBatch(sites, frags, splits, self.libs)
        if bucket:
            b.order = b.length_order()
        return b
