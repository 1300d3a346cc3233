"""Compact evidence schema (v2): what the default scoring kernel streams from HBM.

The wide schema of `evidence.py` (32-byte fragment / split rows, 64-byte site rows) stays the
interchange layout of the Python gather path and of the CPU oracle.  The product path -- native
packer, host->device copies, the tally kernel -- moves the SAME evidence in 16-byte rows:

  site row      int32 [n_sites, 12]  (48 B)
      0 posA 1 posB 2 ciA0 3 ciA1 4 ciB0 5 ciB1 6 var_length
      7 meta     bits 0-1 svtype, 2 o1_is_reverse, 3 o2_is_reverse, 4 SKIP,
                 5 SAME (both breakends on one contig: tidA == tidB)
      8,9 row_off (int64)   first row of the site in `rows`
      10 n_frag  11 n_split  the site's rows are n_frag fragment rows, then n_split split rows

  fragment row  int32 [4]  (16 B)    one read pair / lone primary, sorted(query_name) order
      0 a_start   reference_start of read A          (reference parsers.py:785-796 get_ispan/ospan)
      1 b_end     reference_end of read B (0 if the row has one read)
      2 lenA [0:14) | lenB [14:28) | contig class [28:32)
                 lenX = reference_end - reference_start of the read's single gap-free block;
                 class bits: 28 read A on the site's contig A, 29 read A on contig B,
                             30 read B on contig A,            31 read B on contig B
                 (the only thing the path ever does with a read's contig is compare it with the
                 two breakend contigs -- parsers.py:805,833-834 -- so the comparison result is
                 what the row carries)
      3 mapqA [0:8) | mapqB [8:16) | library [16:25) | flags [25:32)
                 25 PAIRED (num_primary == 2)  26 REV_A  27 REV_B  28 CONT  29 (reserved)
                 30 MULTI_A  31 MULTI_B
      A read whose aligned blocks are not one gap-free run (D / N operations), or whose span
      does not fit 14 bits, is MULTI: the packer evaluates SamFragment.is_ref_seq for it
      (parsers.py:801-816: CIGAR-block overlap with either breakend window, which needs the
      batch's min_aligned -- recorded in CompactBatch.min_aligned and checked at launch) and the
      row carries the outcome in its length field (lenX = 1 hit, 0 no hit).  Single-block reads,
      practically all of them, are tested on the device from a_start / b_end / lenX.

  split row     int32 [4]  (16 B)    SplitRead.query_left / query_right (parsers.py:1017-1028)
      0 l_start  1 r_start
      2 lenL [0:16) | lenR [16:32)
      3 mapqL [0:8) | mapqR [8:16) | 16 SOFT_CLIP | 17 FIRST | 18 WIDE | 19 XEND | class [28:32)
                 class bits: 28 left piece on contig A, 29 left on B, 30 right on A, 31 right on B
      A piece longer than 65535 reference bases sets WIDE and is followed by an XEND row holding
      the two true ends (words 0, 1); a WIDE row never sits in the last slot of a 32-row chunk
      of its site's split rows (an all-zero filler row is put in front of it when it would).

`compact_from_wide()` / `wide_from_compact()` convert between the two; the second one needs no
contig ids (it invents 0 / 1 for the two breakend contigs), which is all the oracle looks at.
"""
from __future__ import annotations

import numpy as np

from . import evidence as ev

CSITE_WORDS, CROW_WORDS = 12, 4
CSITE_BYTES, CROW_BYTES = 48, 16

CS_SAME = 1 << 5

LEN_BITS = 14
LEN_MAX = (1 << LEN_BITS) - 1
CLS_A_ON_A, CLS_A_ON_B, CLS_B_ON_A, CLS_B_ON_B = 1 << 28, 1 << 29, 1 << 30, 1 << 31
LIB_SHIFT, LIB_BITS = 16, 9
LIB_MAX = (1 << LIB_BITS) - 1
CF_PAIRED, CF_REV_A, CF_REV_B, CF_CONT, CF_RESERVED, CF_MULTI_A, CF_MULTI_B = (1 << b for b in range(25, 32))

SLEN_MAX = 0xFFFF
CSP_SOFT, CSP_FIRST, CSP_WIDE, CSP_XEND = 1 << 16, 1 << 17, 1 << 18, 1 << 19

FAKE_NONE = -3        # wide_from_compact(): "some other contig"


class CompactBatch(object):
    """Host-side (numpy) batch in the compact device layout."""

    def __init__(self, sites, rows, libs, order=None, min_aligned=20):
        self.sites = np.ascontiguousarray(sites, dtype=np.int32).reshape(-1, CSITE_WORDS)
        self.rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, CROW_WORDS)
        self.libs = libs
        self.order = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
        self.min_aligned = int(min_aligned)     # the MULTI rows' hit bits were evaluated with this -m

    @property
    def n_sites(self): return self.sites.shape[0]
    @property
    def n_rows(self): return self.rows.shape[0]
    @property
    def n_frag(self): return int(self.sites[:, 10].sum(dtype=np.int64))
    @property
    def n_split(self): return int(self.sites[:, 11].sum(dtype=np.int64))

    def frag_counts(self): return self.sites[:, 10]
    def split_counts(self): return self.sites[:, 11]

    def row_offsets(self):
        return np.ascontiguousarray(self.sites[:, 8:10]).view(np.int64).ravel()

    def algorithmic_bytes(self) -> int:
        """Bytes one pass must move: 48/site + 16/row + 80/site out."""
        return self.n_sites * (CSITE_BYTES + ev.OUT_BYTES) + self.n_rows * CROW_BYTES

    def survey_bytes(self) -> int:
        """SURVEY.md 8(d)'s formula 48 + 16 F + 32 S + 80 on this batch's row counts."""
        return self.n_sites * 128 + 16 * self.n_frag + 32 * self.n_split

    def work(self):
        return self.sites[:, 10].astype(np.int64) + self.sites[:, 11]

    def length_order(self):
        """Site permutation bucketing warps by work (descending rows)."""
        return np.argsort(-self.work(), kind="stable").astype(np.int32)

    def suggest_unit_mode(self, resident_warps=148 * 20, unit_sites=6, margin=1.5) -> int:
        """svgt_cbatch_t.unit_mode for this batch: ramped work units (3) iff the heaviest unit -- the
        `unit_sites` largest sites, which the work-descending launch order puts on one warp -- outweighs `margin`
        times a warp's fair share of the batch (then it is the critical path: measured -8 % on the heavy-tailed
        stress shape at 200k sites, ratio 1.65; -2 % on mixed100k, ratio 1.58; -27 % at 10k sites); full units (1)
        otherwise (the ramp costs 7-9 % on 125k-500k evenly sized sites; at 125k sites of the benchmark shape the
        ratio is 1.09).  Identity launch order (no `order`): the ramp has nothing sorted to spread -> 1."""
        if self.order is None or self.n_sites == 0:
            return 1
        w = self.work()
        k = min(unit_sites, w.shape[0])
        heaviest = int(np.partition(w, w.shape[0] - k)[-k:].sum())
        fair = float(w.sum()) / max(resident_warps, 1)
        return 3 if heaviest > margin * fair else 1

    def log_table_size(self, split_weight=1.0, disc_weight=1.0) -> int:
        if self.n_sites == 0:
            return 2
        nf = int(self.sites[:, 10].max())
        ns = int(self.sites[:, 11].max())
        bound = abs(float(split_weight)) * (2 * nf + ns) + abs(float(disc_weight)) * 2 * nf
        return int(bound) + 8

    def slice_sites(self, lo, hi):
        """Contiguous site range [lo, hi) as an independent batch (multi-GPU shards, slices)."""
        s = self.sites[lo:hi].copy()
        if s.shape[0] == 0:
            return CompactBatch(s, np.zeros((0, CROW_WORDS), np.int32), self.libs, min_aligned=self.min_aligned)
        off = np.ascontiguousarray(s[:, 8:10]).view(np.int64).ravel()
        r0 = int(off[0])
        r1 = int(off[-1]) + int(s[-1, 10]) + int(s[-1, 11])
        s[:, 8:10] = (off - r0).view(np.int32).reshape(-1, 2)
        return CompactBatch(s, self.rows[r0:r1], self.libs, min_aligned=self.min_aligned)


def _u32(a):
    return (np.asarray(a, dtype=np.int64) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


def compact_from_wide(batch: ev.EvidenceBatch, alloc=None, min_aligned=20) -> CompactBatch:
    """Re-encode a wide EvidenceBatch (evidence.py) as a CompactBatch.

    Vectorised; rows come out site by site (fragment rows, then split rows), so the result is
    always laid out in site order whatever the wide batch's offsets were.  The wide layout's EXTRA
    interval rows are consumed here: is_ref_seq of a gapped (MULTI) read is evaluated against the
    two breakend windows (pos -/+ min_aligned, parsers.py:801-816) and travels as one bit.
    `alloc(name, shape, dtype)` may supply the destination arrays (e.g. pinned host memory).
    """
    alloc = alloc or (lambda name, shape, dtype: np.empty(shape, dtype=dtype))
    m = int(min_aligned)
    S = batch.sites.astype(np.int64)
    n = S.shape[0]
    nf_w = S[:, 12].copy()
    ns_w = S[:, 15].copy()
    foff = (S[:, 10] & 0xFFFFFFFF) | (S[:, 11] << 32)
    soff = (S[:, 13] & 0xFFFFFFFF) | (S[:, 14] << 32)
    skip = (S[:, 9] & ev.SITE_SKIP) != 0
    nf_w[skip] = 0            # SKIP sites carry no rows the path may look at
    ns_w[skip] = 0

    def gather_index(off, cnt):
        tot = int(cnt.sum())
        if tot == 0:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        site_of = np.repeat(np.arange(n, dtype=np.int64), cnt)
        start = np.cumsum(cnt) - cnt
        idx = off[site_of] + (np.arange(tot, dtype=np.int64) - start[site_of])
        return idx, site_of

    # ---- fragment rows, gathered in site order
    fidx, fsite = gather_index(foff, nf_w)
    F = batch.frags[fidx].astype(np.int64) if fidx.size else np.zeros((0, ev.FRAG_WORDS), np.int64)
    tA, tB = S[fsite, 6], S[fsite, 7]
    fl = F[:, 7]
    lib = (F[:, 6] >> 16) & 0xFFFF
    if fidx.size and int(lib.max()) > LIB_MAX:
        raise ValueError("compact schema holds library indices up to %d" % LIB_MAX)
    isx = (fl & ev.F_EXTRA) != 0
    hasA = (fl & ev.F_HAS_A) != 0
    hasB = (fl & ev.F_HAS_B) != 0
    paired = (fl & ev.F_PAIRED) != 0
    main = ~isx
    if fidx.size and not (hasA[main].all() and hasB[main & paired].all()):
        raise ValueError("wide fragment rows must hold read A, and read B when PAIRED")
    onAA, onAB = F[:, 4] == tA, F[:, 4] == tB
    onBA, onBB = F[:, 5] == tA, F[:, 5] == tB
    clsA = np.where(onAA, CLS_A_ON_A, 0) | np.where(onAB, CLS_A_ON_B, 0)
    clsB = np.where(onBA, CLS_B_ON_A, 0) | np.where(onBB, CLS_B_ON_B, 0)
    lenA = F[:, 1] - F[:, 0]
    lenB = F[:, 3] - F[:, 2]
    # is_ref_seq of one gap-free interval [s, e) against both windows (the oracle's ref_seq_hit())
    wA0, wA1 = S[fsite, 0] - m, S[fsite, 0] + m
    wB0, wB1 = S[fsite, 1] - m, S[fsite, 1] + m

    def interval_hit(on_a, on_b, s_, e_):
        return (on_a & (wA0 >= 0) & (s_ <= wA0) & (e_ >= wA1)) | (on_b & (wB0 >= 0) & (s_ <= wB0) & (e_ >= wB1))

    hitA_iv = hasA & interval_hit(onAA, onAB, F[:, 0], F[:, 1])
    hitB_iv = hasB & interval_hit(onBA, onBB, F[:, 2], F[:, 3])
    # EXTRA rows feed the next main row of their site: OR their hits into that row's group
    NW = F.shape[0]
    grp = np.cumsum(main) - main            # EXTRA rows share the id of the main row that follows them
    n_main = int(main.sum())
    pendA = np.zeros(n_main + 1, dtype=bool)
    pendB = np.zeros(n_main + 1, dtype=bool)
    if isx.any():
        main_site = np.full(n_main + 1, -1, dtype=np.int64)
        main_site[:n_main] = fsite[main]
        ok = isx & (main_site[grp] == fsite)          # a trailing EXTRA run with no main row in its site feeds nothing
        np.logical_or.at(pendA, grp[ok & hitA_iv], True)
        np.logical_or.at(pendB, grp[ok & hitB_iv], True)
    multiA = main & (((fl & ev.F_MULTI_A) != 0) | (lenA < 0) | (lenA > LEN_MAX))
    multiB = main & hasB & (((fl & ev.F_MULTI_B) != 0) | (lenB < 0) | (lenB > LEN_MAX))
    # a read flagged MULTI takes the EXTRA rows' verdict; one that is merely too long for the length field its own
    hitA = np.where((fl & ev.F_MULTI_A) != 0, pendA[grp], hitA_iv)
    hitB = np.where((fl & ev.F_MULTI_B) != 0, pendB[grp], hitB_iv)
    cflags = (np.where(paired, CF_PAIRED, 0) | np.where(fl & ev.F_REV_A, CF_REV_A, 0)
              | np.where(fl & ev.F_REV_B, CF_REV_B, 0) | np.where(fl & ev.F_CONT, CF_CONT, 0)
              | np.where(multiA, CF_MULTI_A, 0) | np.where(multiB, CF_MULTI_B, 0))
    NF = n_main
    nf_c = np.bincount(fsite[main], minlength=n).astype(np.int64) if NF else np.zeros(n, np.int64)
    R = np.zeros((NF, 4), dtype=np.int64)
    if NF:
        la = np.where(multiA, hitA.astype(np.int64), lenA)[main]
        lb = np.where(multiB, hitB.astype(np.int64), np.where(hasB, lenB, 0))[main]
        hb = hasB[main]
        R[:, 0] = F[main, 0]
        R[:, 1] = np.where(hb, F[main, 3], 0)
        R[:, 2] = la | (lb << LEN_BITS) | clsA[main] | np.where(hb, clsB[main], 0)
        R[:, 3] = ((F[main, 6] & 0xFF) | np.where(hb, F[main, 6] & 0xFF00, 0) | (lib[main] << LIB_SHIFT) | cflags[main])

    # ---- split rows
    sidx, ssite = gather_index(soff, ns_w)
    Q = batch.splits[sidx].astype(np.int64) if sidx.size else np.zeros((0, ev.SPLIT_WORDS), np.int64)
    tA, tB = S[ssite, 6], S[ssite, 7]
    sfl = (Q[:, 6] >> 16) & 0xFFFF
    lenL, lenR = Q[:, 2] - Q[:, 1], Q[:, 5] - Q[:, 4]
    wide = (lenL < 0) | (lenL > SLEN_MAX) | (lenR < 0) | (lenR > SLEN_MAX)
    cls = (np.where(Q[:, 0] == tA, 1 << 28, 0) | np.where(Q[:, 0] == tB, 1 << 29, 0)
           | np.where(Q[:, 3] == tA, 1 << 30, 0) | np.where(Q[:, 3] == tB, 1 << 31, 0))
    C = np.zeros((Q.shape[0], 4), dtype=np.int64)
    C[:, 0], C[:, 1] = Q[:, 1], Q[:, 4]
    C[:, 2] = np.where(wide, 0, (lenL & 0xFFFF) | ((lenR & 0xFFFF) << 16))
    C[:, 3] = ((Q[:, 6] & 0xFFFF) | np.where(sfl & ev.S_SOFT_CLIP, CSP_SOFT, 0) | np.where(sfl & ev.S_FIRST, CSP_FIRST, 0)
               | np.where(wide, CSP_WIDE, 0) | cls)
    ns_c = ns_w.copy()
    if wide.any():
        # rare: rebuild the split rows of the affected sites one by one (filler + XEND rows)
        pieces, start = [], np.cumsum(ns_w) - ns_w
        for g in range(n):
            a, b = int(start[g]), int(start[g] + ns_w[g])
            if not wide[a:b].any():
                pieces.append(C[a:b])
                continue
            rows_g = []
            for j in range(a, b):
                if wide[j]:
                    if len(rows_g) % 32 == 31:
                        rows_g.append(np.zeros(4, np.int64))
                    rows_g.append(C[j])
                    rows_g.append(np.array([Q[j, 2], Q[j, 5], 0, CSP_XEND], dtype=np.int64))
                else:
                    rows_g.append(C[j])
            ns_c[g] = len(rows_g)
            pieces.append(np.array(rows_g, dtype=np.int64).reshape(-1, 4))
        C = np.concatenate(pieces) if pieces else C

    # ---- interleave per site: fragment rows, then split rows
    tot = nf_c + ns_c
    row_off = np.cumsum(tot) - tot
    NR = int(tot.sum())
    rows = alloc("rows", (NR, CROW_WORDS), np.int32)
    if NF:
        fs = np.repeat(np.arange(n, dtype=np.int64), nf_c)
        dst = row_off[fs] + (np.arange(NF, dtype=np.int64) - (np.cumsum(nf_c) - nf_c)[fs])
        rows[dst] = _u32(R)
    if C.shape[0]:
        ss = np.repeat(np.arange(n, dtype=np.int64), ns_c)
        dst = row_off[ss] + nf_c[ss] + (np.arange(C.shape[0], dtype=np.int64) - (np.cumsum(ns_c) - ns_c)[ss])
        rows[dst] = _u32(C)

    sites = alloc("sites", (n, CSITE_WORDS), np.int32)
    cs = np.zeros((n, CSITE_WORDS), dtype=np.int64)
    cs[:, 0:6] = S[:, 0:6]
    cs[:, 6] = S[:, 8]
    cs[:, 7] = (S[:, 9] & 0x1F) | np.where(S[:, 6] == S[:, 7], CS_SAME, 0)
    cs[:, 8], cs[:, 9] = row_off & 0xFFFFFFFF, row_off >> 32
    cs[:, 10], cs[:, 11] = nf_c, ns_c
    sites[:] = _u32(cs)
    out = CompactBatch.__new__(CompactBatch)
    out.sites, out.rows, out.libs, out.order, out.min_aligned = sites, rows, batch.libs, None, m
    if batch.order is not None:
        order = alloc("order", (n,), np.int32)
        order[:] = out.length_order()
        out.order = order
    return out


def wide_from_compact(cb: CompactBatch) -> ev.EvidenceBatch:
    """Decode a CompactBatch into the wide layout (what the oracle reads).

    Contig ids are invented: 0 for contig A, 0 or 1 for contig B (SAME bit), FAKE_NONE for a read
    on neither -- the path only ever tests reads for equality with the breakend contigs.
    """
    S = cb.sites.astype(np.int64)
    n = S.shape[0]
    nf, ns = S[:, 10], S[:, 11]
    off = (S[:, 8] & 0xFFFFFFFF) | (S[:, 9] << 32)
    same = (S[:, 7] & CS_SAME) != 0
    tidA = np.zeros(n, np.int64)
    tidB = np.where(same, 0, 1)

    def rows_of(first, cnt):
        tot = int(cnt.sum())
        if tot == 0:
            return np.zeros((0, 4), np.int64), np.zeros(0, np.int64)
        site_of = np.repeat(np.arange(n, dtype=np.int64), cnt)
        start = np.cumsum(cnt) - cnt
        idx = first[site_of] + (np.arange(tot, dtype=np.int64) - start[site_of])
        return cb.rows[idx].astype(np.int64) & 0xFFFFFFFF, site_of

    def tid_from(cls_on_a, cls_on_b, site_of):
        return np.where(cls_on_a, tidA[site_of], np.where(cls_on_b, tidB[site_of], FAKE_NONE))

    def s32(a):
        return (a & 0xFFFFFFFF).astype(np.uint32).view(np.int32).astype(np.int64)

    m = cb.min_aligned
    R, fs = rows_of(off, nf)
    w2, w3 = R[:, 2], R[:, 3]
    paired = (w3 & CF_PAIRED) != 0
    a_start, b_end = s32(R[:, 0]), s32(R[:, 1])
    lenA, lenB = w2 & LEN_MAX, (w2 >> LEN_BITS) & LEN_MAX
    multiA, multiB = (w3 & CF_MULTI_A) != 0, (w3 & CF_MULTI_B) != 0
    # a second read that is not part of a pair shows only through its fields (one on neither
    # breakend contig with no span cannot contribute anything, so dropping it changes nothing)
    hasB = paired | ((w2 & (CLS_B_ON_A | CLS_B_ON_B)) != 0) | (lenB != 0) | multiB | (b_end != 0)
    onAA, onAB = (w2 & CLS_A_ON_A) != 0, (w2 & CLS_A_ON_B) != 0
    onBA, onBB = (w2 & CLS_B_ON_A) != 0, (w2 & CLS_B_ON_B) != 0
    tA_read = tid_from(onAA, onAB, fs)
    tB_read = tid_from(onBA, onBB, fs)
    lib = (w3 >> LIB_SHIFT) & LIB_MAX
    # a MULTI read with its hit bit set becomes MULTI + one EXTRA row whose interval is the window it covers
    posA, posB = s32(S[fs, 0]), s32(S[fs, 1])
    xA = multiA & (lenA != 0)
    xB = multiB & (lenB != 0)
    n_out = 1 + xA.astype(np.int64) + xB.astype(np.int64)
    pos = np.cumsum(n_out) - n_out
    NW = int(n_out.sum())
    F = np.zeros((NW, ev.FRAG_WORDS), dtype=np.int64)

    def window_of(on_a, on_b):
        use_a = on_a & (posA - m >= 0)
        use_b = ~use_a & on_b & (posB - m >= 0)
        if not (use_a | use_b)[xsel].all():
            raise ValueError("MULTI hit bit set on a read that can reach neither breakend window")
        return np.where(use_a, posA, posB) - m, np.where(use_a, posA, posB) + m

    for xsel, on_a, on_b, slot_b in ((xA, onAA, onAB, False), (xB, onBA, onBB, True)):
        if not xsel.any():
            continue
        lo_, hi_ = window_of(on_a, on_b)
        p = (pos + (xA & slot_b))[xsel]
        cfl = np.where(w3 & CF_CONT, ev.F_CONT, 0)[xsel]
        if slot_b:
            F[p, 2], F[p, 3], F[p, 5] = lo_[xsel], hi_[xsel], tB_read[xsel]
            F[p, 7] = ev.F_EXTRA | ev.F_HAS_B | cfl
        else:
            F[p, 0], F[p, 1], F[p, 4] = lo_[xsel], hi_[xsel], tA_read[xsel]
            F[p, 7] = ev.F_EXTRA | ev.F_HAS_A | cfl
        F[p, 6] = (lib << 16)[xsel]
    pm_ = pos + n_out - 1
    F[pm_, 0] = a_start
    F[pm_, 1] = a_start + np.where(multiA, 0, lenA)
    F[pm_, 2] = np.where(hasB, b_end - np.where(multiB, 0, lenB), 0)
    F[pm_, 3] = np.where(hasB, b_end, 0)
    F[pm_, 4] = tA_read
    F[pm_, 5] = np.where(hasB, tB_read, 0)
    F[pm_, 6] = (w3 & 0xFFFF) | (lib << 16)
    F[pm_, 7] = (ev.F_HAS_A | np.where(hasB, ev.F_HAS_B, 0) | np.where(paired, ev.F_PAIRED, 0)
                 | np.where(w3 & CF_REV_A, ev.F_REV_A, 0) | np.where(w3 & CF_REV_B, ev.F_REV_B, 0)
                 | np.where(w3 & CF_CONT, ev.F_CONT, 0) | np.where(multiA, ev.F_MULTI_A, 0)
                 | np.where(multiB, ev.F_MULTI_B, 0))
    nf_wide = np.bincount(fs, weights=n_out, minlength=n).astype(np.int64) if R.shape[0] else np.zeros(n, np.int64)

    C, ss = rows_of(off + nf, ns)
    w2, w3 = C[:, 2], C[:, 3]
    xend = (w3 & CSP_XEND) != 0
    wide = (w3 & CSP_WIDE) != 0
    l_start, r_start = s32(C[:, 0]), s32(C[:, 1])
    l_end, r_end = l_start + (w2 & 0xFFFF), r_start + ((w2 >> 16) & 0xFFFF)
    if wide.any():
        wi = np.nonzero(wide)[0]
        l_end[wi], r_end[wi] = l_start[wi + 1], r_start[wi + 1]
    Q = np.zeros((C.shape[0], ev.SPLIT_WORDS), dtype=np.int64)
    Q[:, 0] = tid_from((w3 & (1 << 28)) != 0, (w3 & (1 << 29)) != 0, ss)
    Q[:, 1], Q[:, 2] = l_start, l_end
    Q[:, 3] = tid_from((w3 & (1 << 30)) != 0, (w3 & (1 << 31)) != 0, ss)
    Q[:, 4], Q[:, 5] = r_start, r_end
    sfl = np.where(w3 & CSP_SOFT, ev.S_SOFT_CLIP, 0) | np.where(w3 & CSP_FIRST, ev.S_FIRST, 0)
    Q[:, 6] = (w3 & 0xFFFF) | (sfl << 16)
    Q[xend] = 0
    Q[xend, 0] = Q[xend, 3] = FAKE_NONE       # contributes (0.0 + 0.0) / 2 to the pending sub-total

    sites = np.zeros((n, ev.SITE_WORDS), dtype=np.int64)
    sites[:, 0:6] = s32(S[:, 0:6])
    sites[:, 6], sites[:, 7] = tidA, tidB
    sites[:, 8] = s32(S[:, 6])
    sites[:, 9] = S[:, 7] & 0x1F
    f0 = np.cumsum(nf_wide) - nf_wide
    s0 = np.cumsum(ns) - ns
    sites[:, 10], sites[:, 11], sites[:, 12] = f0 & 0xFFFFFFFF, f0 >> 32, nf_wide
    sites[:, 13], sites[:, 14], sites[:, 15] = s0 & 0xFFFFFFFF, s0 >> 32, ns
    return ev.EvidenceBatch(_u32(sites), _u32(F), _u32(Q), cb.libs,
                            None if cb.order is None else cb.order.copy())
