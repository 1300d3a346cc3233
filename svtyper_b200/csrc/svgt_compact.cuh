/*
 * svgt_compact.cuh -- row scorers of the compact-schema tally kernel (svgt_compact.cu).
 *
 * Rows are 16 bytes (svtyper_b200/compact.py): a fragment row is {a_start, b_end, lenA | lenB | contig
 * class, mapqA | mapqB | library | flags}; a split row is {l_start, r_start, lenL | lenR, mapqL | mapqR |
 * flags | contig class}.  One LDS.128 per lane fetches a row; everything a row is tested against is
 * warp-uniform per site (CSiteF, three LDS.128) or per (site, library) (WinF, two LDS.128).
 *
 * Same arithmetic as the wide-row scorers (reference citations in svgt_lean.cuh / svgt_coop.cuh); what
 * changes with the encoding:
 *   is_ref_seq (parsers.py:801-816)  a read [s, s + len) covers the window [w0, w0 + 2m) iff
 *       (unsigned)(w0 - s) < max(len - 2m + 1, 0); for read B, stored by its end e,
 *       (unsigned)(e - w1) < max(len - 2m + 1, 0).  Exact for every int32 input: a wrapped difference is
 *       >= 2^30 while the bound is < 2^14.
 *   contig tests (parsers.py:805,833-834)  are the four class bits of the row.
 *   window validity (max(0, pos - m) shortens a window at the contig start: no read can cover it) is a
 *       per-site mask over the class bits, so such sites take the same chain as every other.
 * Reads the 16-byte form cannot test on the device (gapped alignments, spans beyond 14 bits) are MULTI rows
 * whose is_ref_seq outcome the packer evaluated (one bit in the length field); extra primaries of a fragment
 * are CONT rows.  Both leave the straight-line chain through one vote; p_concordant ties and libraries
 * outside the integer rewrite are decoded into the wide form (decode_wide) for the literal fp64 row
 * (slow_row), which is shared with the wide-row kernels.
 */
#pragma once
#include "svgt_device.cuh"

namespace {

/* ---- pieces shared by the scorers below ---- */
constexpr int kWLibs = 4;           /* libraries with per-site windows cached in shared memory */

/* per-(warp, g) site scalars, read warp-uniformly.  First 32 bytes are the hot ones. */
struct SiteS {
    int tA, tB, wA0, wA1;
    int wB0, wB1, meta, var_length;
    int posA, posB, ciA0, ciA1;
    int ciB0, ciB1, dAB, nf;
    long long foff, soff;
    int ns, slot, pad1, pad2;
};  /* 96 B */

__device__ __forceinline__ void set_win(unsigned &lo_out, unsigned &w1_out, int lo, int hi, bool enable)
{
    lo_out = (unsigned)lo;
    w1_out = (enable && hi >= lo) ? (unsigned)(hi - lo) + 1u : 0u;
}


/* everything the integer fast path does not cover, evaluated the long way for one row:
 * libraries beyond the window cache or not provably integer-exact, histogram counts >= 2^26,
 * breakends within min_aligned of the contig start, malformed DEL lengths, p_concordant ties */
__device__ __noinline__ void slow_row(const SvgtParams &p, const Tables &t, const SiteS &S, const int4 lo,
                                      const int4 hi, const LibK *s_lib, int m, int &err, bool &alt, bool &refA,
                                      bool &refB, bool &pc)
{
    const int lib = (int)(((unsigned)hi.z) >> 16);
    if (lib >= p.n_lib) { err = SVGT_ERR_LIB_INDEX; alt = refA = refB = pc = false; return; }
    LibK Ls;
    if (lib >= SVGT_SMEM_LIBS) { int e = 0; Ls = derive_lib(p, lib, &e); }
    const LibK &L = (lib < SVGT_SMEM_LIBS) ? s_lib[lib] : Ls;
    const int svtype = S.meta & 3;
    const bool is_del = svtype == SV_DEL;
    const int o1 = (S.meta >> 2) & 1, o2 = (S.meta >> 3) & 1;
    const bool small_del = is_del && ((double)((long long)S.posB - S.posA) < L.two_sd);
    alt = !small_del && straddle_literal(lo, hi, S.tA, S.posA, S.ciA0, S.ciA1, S.tB, S.posB, S.ciB0, S.ciB1, o1, o2,
                                         m, L.flank);
    if (svtype == SV_INV)
        alt = alt || straddle_literal(lo, hi, S.tA, S.posA, S.ciA0, S.ciA1, S.tB, S.posB, S.ciB0, S.ciB1, !o1, !o2, m,
                                      L.flank);
    refA = !small_del && straddle_literal(lo, hi, S.tA, S.posA, 0, 0, S.tA, S.posA, 0, 0, 0, 1, m, L.flank);
    refB = !small_del && straddle_literal(lo, hi, S.tB, S.posB, 0, 0, S.tB, S.posB, 0, 0, 0, 1, m, L.flank);
    pc = p_concordant(t, L, lo.x, lo.w, is_del, S.var_length);
}

/* one 32-row fragment chunk of site S scored by the warp (phase A): what every lane parks for its row */
struct FragOut {
    double s, p_ref, p_alt;         /* ref_seq / ref_span / alt_span addends this row parks (see below)     */
    int lead;                       /* warp-uniform: leading rows that continue the previous chunk's fragment */
};

/*
 * Continuation rows are resolved HERE, lane-parallel, so that phase B is one add per row:
 * in the sso order (singlesample.py:254-259, :367-378) a fragment's reads are summed into a sub-total
 * first -- sub = ((0 + a1) + b1) + a2 ... over its rows (CONT rows; EXTRA interval rows add nothing) --
 * and the sub-total is added to the site sum when the next fragment starts.  The running sub-total is
 * folded forward along each fragment's rows with shuffles, in row order, and parked in the fragment's
 * LAST row of the chunk; its earlier rows park 0.0 (x + 0.0 is exact), so phase B just does
 * acc += pend, pend = s for every row and a fragment that continues in the next chunk stays pending.
 * Only continuation rows at the very start of a chunk (their fragment began in the previous chunk) are
 * left to phase B: `lead` of them update the carried sub-total first.
 */
struct SplitOut { double vseq, vclip; int lead; };

/* sums parked in the site's output row between the two launches */
struct ParkedSums { double ref_seq, alt_seq, alt_clip, ref_span, alt_span; };


constexpr unsigned kOneHi = 0x3FF00000u, kHalfHi = 0x3FE00000u;   /* high words of 1.0 and 0.5 */
constexpr unsigned kTrapLk = 0x80000000u;
constexpr int kHistLenBits = 14;                                  /* packed hist length < 16384 */

/* per-(site, library slot); slot kWLibs is the trap entry for library indices beyond the cache */
struct WinF {
    unsigned altA_lo, altA_w1, altB_lo, altB_w1;
    unsigned FL, FL1, Lk, hpk;   /* hpk = shared byte address of the library's counts | len << 18 */
};

/* per-CTA: where the lean copy of a cached library's histogram lives */
struct LibF { unsigned addr; int len; int ok; int pad; };

__device__ __forceinline__ WinF make_winf(const SiteS &S, const LibK &L, const LibF &F, int m, unsigned zero_addr)
{
    WinF w;
    const int svtype = S.meta & 3;
    const bool is_del = svtype == SV_DEL;
    const int Lk = is_del ? S.var_length : L.nondel_L;
    if (!F.ok || (is_del && Lk < 0)) {
        w.altA_lo = w.altA_w1 = w.altB_lo = w.altB_w1 = 0u;
        w.FL = 0u; w.FL1 = 0u; w.Lk = kTrapLk; w.hpk = zero_addr;      /* len 0: both keys clamp onto the zero word */
        return w;
    }
    const int o1 = (S.meta >> 2) & 1, o2 = (S.meta >> 3) & 1;
    const bool en = !(is_del && (S.dAB < L.ceil2sd));       /* singlesample.py:289,328 small deletions */
    const int FL = L.FL;
    const int LA = S.posA + S.ciA0 - m, HA = S.posA + S.ciA1 - m;
    const int LB = S.posB + S.ciB0 + m + 1, HB = S.posB + S.ciB1 + m + 1;
    set_win(w.altA_lo, w.altA_w1, LA - (o1 ? 0 : FL), HA + (o1 ? FL : 0), en);
    set_win(w.altB_lo, w.altB_w1, LB - (o2 ? 0 : FL), HB + (o2 ? FL : 0), en);
    w.FL = (unsigned)FL;
    w.FL1 = (en && FL >= 0) ? (unsigned)FL + 1u : 0u;
    w.Lk = (!is_del && Lk < 0) ? 0x7fffffffu : (unsigned)Lk;
    w.hpk = F.addr | ((unsigned)F.len << 18);
    return w;
}

/* sso association of a fragment's extra rows (singlesample.py:254-259, :367-378): the note above FragOut */
struct FoldOut { double s, p_ref, p_alt; int lead; };

__device__ __noinline__ FoldOut fold_continuations(const int lane, const int n, const unsigned nm, const unsigned vm,
                                                   const double va, const double vb, double s, const double p_ref,
                                                   const double p_alt)
{
    const unsigned full = 0xffffffffu;
    FoldOut o;
    o.s = s; o.p_ref = p_ref; o.p_alt = p_alt;
    const unsigned NN = vm & ~nm;                   /* rows that continue a fragment */
    o.lead = nm ? __ffs(nm) - 1 : (n < 32 ? n : 32);
    const bool nonnew = (NN >> lane) & 1u;
    const bool inner = nonnew && lane >= o.lead;    /* continues a fragment that starts in this chunk */
    const unsigned below = nm & ((1u << lane) - 1u);
    const int dist = inner ? lane - (31 - __clz(below)) : 0;
    const bool pe_too = __ballot_sync(full, inner && (p_ref != 0.0 || p_alt != 0.0)) != 0u;
    for (int k = 1; k < 32; ++k) {
        if (!__any_sync(full, dist >= k)) break;
        const double up = __shfl_up_sync(full, o.s, 1);
        if (dist == k) o.s = __dadd_rn(__dadd_rn(up, va), vb);
        if (pe_too) {
            const double ur = __shfl_up_sync(full, o.p_ref, 1), ua = __shfl_up_sync(full, o.p_alt, 1);
            if (dist == k) { o.p_ref = __dadd_rn(ur, p_ref); o.p_alt = __dadd_rn(ua, p_alt); }
        }
    }
    const bool has_next = lane + 1 < 32 && ((NN >> (lane + 1 < 32 ? lane + 1 : 31)) & 1u);
    if (lane >= o.lead && has_next) {
        o.s = 0.0;
        if (pe_too) { o.p_ref = 0.0; o.p_alt = 0.0; }
    }
    return o;
}

/* splits that are not the FIRST of their fragment are folded into the first one's sub-totals, as in
 * score_split_chunk(); .p_ref / .p_alt of the result carry alt_seq / alt_clip */
__device__ __noinline__ FoldOut fold_splits(const int lane, const int n, const unsigned nm, const double vs0,
                                            const double vc0)
{
    const unsigned full = 0xffffffffu;
    const unsigned vm = n >= 32 ? full : ((1u << n) - 1u);
    FoldOut o;
    o.s = 0.0; o.p_ref = vs0; o.p_alt = vc0;
    const unsigned NN = vm & ~nm;
    o.lead = nm ? __ffs(nm) - 1 : (n < 32 ? n : 32);
    const bool nonnew = (NN >> lane) & 1u;
    const bool inner = nonnew && lane >= o.lead;
    const unsigned below = nm & ((1u << lane) - 1u);
    const int dist = inner ? lane - (31 - __clz(below)) : 0;
    for (int k = 1; k < 32; ++k) {
        if (!__any_sync(full, dist >= k)) break;
        const double us = __shfl_up_sync(full, o.p_ref, 1), uc = __shfl_up_sync(full, o.p_alt, 1);
        if (dist == k) { o.p_ref = __dadd_rn(us, vs0); o.p_alt = __dadd_rn(uc, vc0); }
    }
    const bool has_next = lane + 1 < 32 && ((NN >> (lane + 1 < 32 ? lane + 1 : 31)) & 1u);
    if (lane >= o.lead && has_next) { o.p_ref = 0.0; o.p_alt = 0.0; }
    return o;
}



/* compact schema constants (svtyper_b200/compact.py) */
enum : unsigned {
    CS_SAME = 1u << 5,
    CLS_A_ON_A = 1u << 28, CLS_A_ON_B = 1u << 29, CLS_B_ON_A = 1u << 30, CLS_B_ON_B = 1u << 31,
    CF_PAIRED = 1u << 25, CF_REV_A = 1u << 26, CF_REV_B = 1u << 27, CF_CONT = 1u << 28,
    CF_MULTI_A = 1u << 30, CF_MULTI_B = 1u << 31,
    CSP_SOFT = 1u << 16, CSP_FIRST = 1u << 17, CSP_WIDE = 1u << 18, CSP_XEND = 1u << 19
};
constexpr int kLenBits = 14;
constexpr unsigned kLenMask = (1u << kLenBits) - 1u;
constexpr unsigned kLibMaskC = 0x1FFu;

/* per-site constants, warp-uniform LDS.128s (two for the one-contig chain, four for the general one) */
struct CSiteF {
    int wA0, wA1, wB0, wB1;
    int pat, del, fast, m21;     /* pat: (w3 & 0x0E000000) of an alt-orientation pair; m21 = 2 * min_aligned - 1;
                                    fast: 1 one contig, not INV, both is_ref_seq windows valid; 2 general chain */
    int sgnA, sgnB, inv, same;
    unsigned mAA, mAB, mBA, mBB; /* class bit a read must carry to cover window A / B (0: the window is invalid) */
};

/* the wide form of a compact fragment row for the literal path (slow_row): invented contig ids
 * (0 = contig A, tB = contig B); MULTI rows keep a zero-length interval, their hit is not slow_row's business */
__device__ __forceinline__ void decode_wide(const int4 r, const int tB, int4 &lo, int4 &hi)
{
    const unsigned w2 = (unsigned)r.z, w3 = (unsigned)r.w;
    const int tidA = (w2 & CLS_A_ON_A) ? 0 : ((w2 & CLS_A_ON_B) ? tB : -3);
    const int tidB = (w2 & CLS_B_ON_A) ? 0 : ((w2 & CLS_B_ON_B) ? tB : -3);
    const int lenA = (w3 & CF_MULTI_A) ? 0 : (int)(w2 & kLenMask);
    const int lenB = (w3 & CF_MULTI_B) ? 0 : (int)((w2 >> kLenBits) & kLenMask);
    const bool hasB = (w3 & CF_PAIRED) || (w2 & (CLS_B_ON_A | CLS_B_ON_B));
    lo = make_int4(r.x, r.x + lenA, r.y - lenB, r.y);
    hi.x = tidA; hi.y = tidB;
    hi.z = (int)((w3 & 0xFFFFu) | (((w3 >> 16) & kLibMaskC) << 16));
    hi.w = F_HAS_A | (hasB ? F_HAS_B : 0) | ((w3 & CF_REV_A) ? F_REV_A : 0) | ((w3 & CF_REV_B) ? F_REV_B : 0) |
           ((w3 & CF_PAIRED) ? F_PAIRED : 0) | ((w3 & CF_CONT) ? F_CONT : 0);
}

/*
 * One row of a fast site on one contig (not an inversion).  Outputs as in fast_row(): high words of the
 * {0,1} factors of a and b, of the {0,.5,1} factor of p_ref and the {0,1} factor of p_alt, and `tie`.
 *   %5 a_start  %6 b_end  %7 w2  %8 w3
 *   %9..%12 wA0 wA1 wB0 wB1    %13 pat  %14 del  %15 m21
 *   %16..%19 altA_lo altA_w1 altB_lo altB_w1    %20 FL  %21 FL1  %22 Lk  %23 hpk
 */
__device__ __forceinline__ void crow_fast(const int4 r, const int4 f0, const int4 f1, const uint4 w0, const uint4 w1,
                                          double &hA, double &hB, double &wref, double &walt, int &tie)
{
    asm("{\n\t"
        ".reg .pred eA, eB, p, q, hA, hB, pe0, pa, pr0, ra, rb, pc, pt, pdel, both, any, x1, ron, aon, ptrap;\n\t"
        ".reg .b32 t, u, d, x, o, k2, len, hb, i1, i2, a1, a2, h1, h2, l19, zr, lmA, lmB;\n\t"
        "mov.b32 zr, 0;\n\t"
        /* span bounds: max(len - 2m + 1, 0) */
        "and.b32 t, %7, 0x3FFF;\n\t"
        "sub.s32 t, t, %15;\n\t"
        "max.s32 lmA, t, 0;\n\t"
        "bfe.u32 t, %7, 14, 14;\n\t"
        "sub.s32 t, t, %15;\n\t"
        "max.s32 lmB, t, 0;\n\t"
        /* contig class: on the site's (single) contig */
        "and.b32 t, %7, 0x10000000;\n\t"
        "setp.ne.s32 eA, t, 0;\n\t"
        "and.b32 t, %7, 0x40000000;\n\t"
        "setp.ne.s32 eB, t, 0;\n\t"
        /* is_ref_seq, read A then read B */
        "sub.s32 u, %9, %5;\n\t"
        "setp.lt.and.u32 p, u, lmA, eA;\n\t"
        "sub.s32 u, %11, %5;\n\t"
        "setp.lt.and.u32 q, u, lmA, eA;\n\t"
        "or.pred hA, p, q;\n\t"
        "sub.s32 u, %6, %10;\n\t"
        "setp.lt.and.u32 p, u, lmB, eB;\n\t"
        "sub.s32 u, %6, %12;\n\t"
        "setp.lt.and.u32 q, u, lmB, eB;\n\t"
        "or.pred hB, p, q;\n\t"
        "selp.b32 t, 0x3FF00000, 0, hA;\n\t"
        "mov.b64 %0, {zr, t};\n\t"
        "selp.b32 t, 0x3FF00000, 0, hB;\n\t"
        "mov.b64 %1, {zr, t};\n\t"
        /* paired-end straddles */
        "and.pred pe0, eA, eB;\n\t"
        "and.b32 t, %8, 0x0E000000;\n\t"
        "setp.eq.and.s32 pa, t, %13, pe0;\n\t"
        "setp.eq.and.s32 pr0, t, 0x0A000000, pe0;\n\t"
        "sub.s32 d, %5, %16;\n\t"
        "setp.lt.and.u32 pa, d, %17, pa;\n\t"
        "sub.s32 d, %6, %18;\n\t"
        "setp.lt.and.u32 pa, d, %19, pa;\n\t"
        "add.s32 x, %5, %20;\n\t"
        "sub.s32 d, x, %9;\n\t"
        "setp.lt.and.u32 ra, d, %21, pr0;\n\t"
        "sub.s32 d, %6, %10;\n\t"
        "add.s32 d, d, -1;\n\t"
        "setp.lt.and.u32 ra, d, %21, ra;\n\t"
        "sub.s32 d, x, %11;\n\t"
        "setp.lt.and.u32 rb, d, %21, pr0;\n\t"
        "sub.s32 d, %6, %12;\n\t"
        "add.s32 d, d, -1;\n\t"
        "setp.lt.and.u32 rb, d, %21, rb;\n\t"
        /* p_concordant on the counts; keys outside the histogram clamp onto the zero sentinel */
        "sad.s32 o, %6, %5, 0;\n\t"
        "sub.s32 k2, o, %22;\n\t"
        "shr.u32 len, %23, 18;\n\t"
        "and.b32 hb, %23, 0x3ffff;\n\t"
        "min.u32 i1, o, len;\n\t"
        "min.u32 i2, k2, len;\n\t"
        "mad.lo.u32 a1, i1, 4, hb;\n\t"
        "mad.lo.u32 a2, i2, 4, hb;\n\t"
        "ld.shared.u32 h1, [a1];\n\t"
        "ld.shared.u32 h2, [a2];\n\t"
        "mul.lo.u32 l19, h1, 19;\n\t"
        "setp.gt.u32 pc, l19, h2;\n\t"
        "setp.eq.u32 pt, l19, h2;\n\t"
        "setp.ne.and.u32 pt, h2, 0, pt;\n\t"
        "setp.lt.s32 ptrap, %22, 0;\n\t"
        "or.pred pt, pt, ptrap;\n\t"
        "selp.s32 %4, 1, 0, pt;\n\t"
        /* weights */
        "setp.ne.s32 pdel, %14, 0;\n\t"
        "and.pred both, ra, rb;\n\t"
        "or.pred any, ra, rb;\n\t"
        "and.pred x1, both, !pdel;\n\t"
        "and.pred ron, any, !x1;\n\t"
        "and.pred ron, ron, pc;\n\t"
        "and.pred x1, pdel, pc;\n\t"
        "and.pred aon, pa, !x1;\n\t"
        "selp.b32 t, 0x3FF00000, 0x3FE00000, both;\n\t"
        "selp.b32 t, t, 0, ron;\n\t"
        "mov.b64 %2, {zr, t};\n\t"
        "selp.b32 t, 0x3FF00000, 0, aon;\n\t"
        "mov.b64 %3, {zr, t};\n\t"
        "}"
        : "=d"(hA), "=d"(hB), "=d"(wref), "=d"(walt), "=r"(tie)
        : "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w),                                                /* %5..%8   */
          "r"(f0.x), "r"(f0.y), "r"(f0.z), "r"(f0.w), "r"(f1.x), "r"(f1.y), "r"(f1.w),           /* %9..%15  */
          "r"(w0.x), "r"(w0.y), "r"(w0.z), "r"(w0.w), "r"(w1.x), "r"(w1.y), "r"(w1.z), "r"(w1.w) /* %16..%23 */);
}

/*
 * The same row for any other site -- breakends on two contigs, an inversion (the reciprocal orientation also
 * straddles, singlesample.py:296-303), or a breakend whose is_ref_seq window is cut off by the contig start:
 *   %24 sgnA  %25 sgnB  %26 inv    %27..%30 class masks of (read A, window A), (A, B), (read B, window A), (B, B)
 */
__device__ __forceinline__ void crow_gen(const int4 r, const int4 f0, const int4 f1, const int4 f2, const uint4 f3,
                                         const uint4 w0, const uint4 w1, double &hA, double &hB, double &wref,
                                         double &walt, int &tie)
{
    asm("{\n\t"
        ".reg .pred eaa, eab, eba, ebb, haa, hab, hba, hbb, p, q, hA, hB, pe0, pa, prc, pr0, ra, rb, pc, pt, pdel, pinv, both, any, x1, ron, aon, ptrap;\n\t"
        ".reg .b32 t, u, d, x, o, k2, len, hb, i1, i2, a1, a2, h1, h2, l19, zr, lmA, lmB;\n\t"
        "mov.b32 zr, 0;\n\t"
        "and.b32 t, %7, 0x3FFF;\n\t"
        "sub.s32 t, t, %15;\n\t"
        "max.s32 lmA, t, 0;\n\t"
        "bfe.u32 t, %7, 14, 14;\n\t"
        "sub.s32 t, t, %15;\n\t"
        "max.s32 lmB, t, 0;\n\t"
        "and.b32 t, %7, 0x10000000;\n\t"
        "setp.ne.s32 eaa, t, 0;\n\t"
        "and.b32 t, %7, 0x20000000;\n\t"
        "setp.ne.s32 eab, t, 0;\n\t"
        "and.b32 t, %7, 0x40000000;\n\t"
        "setp.ne.s32 eba, t, 0;\n\t"
        "and.b32 t, %7, 0x80000000;\n\t"
        "setp.ne.s32 ebb, t, 0;\n\t"
        /* is_ref_seq: on contig A against window A, on contig B against window B (valid windows only) */
        "and.b32 t, %7, %27;\n\t"
        "setp.ne.s32 haa, t, 0;\n\t"
        "and.b32 t, %7, %28;\n\t"
        "setp.ne.s32 hab, t, 0;\n\t"
        "and.b32 t, %7, %29;\n\t"
        "setp.ne.s32 hba, t, 0;\n\t"
        "and.b32 t, %7, %30;\n\t"
        "setp.ne.s32 hbb, t, 0;\n\t"
        "sub.s32 u, %9, %5;\n\t"
        "setp.lt.and.u32 p, u, lmA, haa;\n\t"
        "sub.s32 u, %11, %5;\n\t"
        "setp.lt.and.u32 q, u, lmA, hab;\n\t"
        "or.pred hA, p, q;\n\t"
        "sub.s32 u, %6, %10;\n\t"
        "setp.lt.and.u32 p, u, lmB, hba;\n\t"
        "sub.s32 u, %6, %12;\n\t"
        "setp.lt.and.u32 q, u, lmB, hbb;\n\t"
        "or.pred hB, p, q;\n\t"
        "selp.b32 t, 0x3FF00000, 0, hA;\n\t"
        "mov.b64 %0, {zr, t};\n\t"
        "selp.b32 t, 0x3FF00000, 0, hB;\n\t"
        "mov.b64 %1, {zr, t};\n\t"
        /* alt straddle: read A on contig A, read B on contig B, alt orientation; INV also the reciprocal one */
        "and.pred pe0, eaa, ebb;\n\t"
        "and.b32 t, %8, 0x0E000000;\n\t"
        "setp.eq.and.s32 pa, t, %13, pe0;\n\t"
        "sub.s32 d, %5, %16;\n\t"
        "setp.lt.and.u32 pa, d, %17, pa;\n\t"
        "sub.s32 d, %6, %18;\n\t"
        "setp.lt.and.u32 pa, d, %19, pa;\n\t"
        "setp.ne.s32 pinv, %26, 0;\n\t"
        "and.pred prc, pe0, pinv;\n\t"
        "xor.b32 u, %13, 0x0C000000;\n\t"
        "setp.eq.and.s32 prc, t, u, prc;\n\t"
        "sub.s32 d, %5, %16;\n\t"
        "mad.lo.s32 d, %24, %20, d;\n\t"
        "setp.lt.and.u32 prc, d, %17, prc;\n\t"
        "sub.s32 d, %6, %18;\n\t"
        "mad.lo.s32 d, %25, %20, d;\n\t"
        "setp.lt.and.u32 prc, d, %19, prc;\n\t"
        "or.pred pa, pa, prc;\n\t"
        /* reference FR pairs: both reads on contig A around A, both on contig B around B */
        "setp.eq.s32 pr0, t, 0x0A000000;\n\t"
        "and.pred ra, eaa, eba;\n\t"
        "and.pred ra, ra, pr0;\n\t"
        "and.pred rb, eab, ebb;\n\t"
        "and.pred rb, rb, pr0;\n\t"
        "add.s32 x, %5, %20;\n\t"
        "sub.s32 d, x, %9;\n\t"
        "setp.lt.and.u32 ra, d, %21, ra;\n\t"
        "sub.s32 d, %6, %10;\n\t"
        "add.s32 d, d, -1;\n\t"
        "setp.lt.and.u32 ra, d, %21, ra;\n\t"
        "sub.s32 d, x, %11;\n\t"
        "setp.lt.and.u32 rb, d, %21, rb;\n\t"
        "sub.s32 d, %6, %12;\n\t"
        "add.s32 d, d, -1;\n\t"
        "setp.lt.and.u32 rb, d, %21, rb;\n\t"
        /* p_concordant */
        "sad.s32 o, %6, %5, 0;\n\t"
        "sub.s32 k2, o, %22;\n\t"
        "shr.u32 len, %23, 18;\n\t"
        "and.b32 hb, %23, 0x3ffff;\n\t"
        "min.u32 i1, o, len;\n\t"
        "min.u32 i2, k2, len;\n\t"
        "mad.lo.u32 a1, i1, 4, hb;\n\t"
        "mad.lo.u32 a2, i2, 4, hb;\n\t"
        "ld.shared.u32 h1, [a1];\n\t"
        "ld.shared.u32 h2, [a2];\n\t"
        "mul.lo.u32 l19, h1, 19;\n\t"
        "setp.gt.u32 pc, l19, h2;\n\t"
        "setp.eq.u32 pt, l19, h2;\n\t"
        "setp.ne.and.u32 pt, h2, 0, pt;\n\t"
        "setp.lt.s32 ptrap, %22, 0;\n\t"
        "or.pred pt, pt, ptrap;\n\t"
        "selp.s32 %4, 1, 0, pt;\n\t"
        /* weights */
        "setp.ne.s32 pdel, %14, 0;\n\t"
        "and.pred both, ra, rb;\n\t"
        "or.pred any, ra, rb;\n\t"
        "and.pred x1, both, !pdel;\n\t"
        "and.pred ron, any, !x1;\n\t"
        "and.pred ron, ron, pc;\n\t"
        "and.pred x1, pdel, pc;\n\t"
        "and.pred aon, pa, !x1;\n\t"
        "selp.b32 t, 0x3FF00000, 0x3FE00000, both;\n\t"
        "selp.b32 t, t, 0, ron;\n\t"
        "mov.b64 %2, {zr, t};\n\t"
        "selp.b32 t, 0x3FF00000, 0, aon;\n\t"
        "mov.b64 %3, {zr, t};\n\t"
        "}"
        : "=d"(hA), "=d"(hB), "=d"(wref), "=d"(walt), "=r"(tie)
        : "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w),                                                 /* %5..%8   */
          "r"(f0.x), "r"(f0.y), "r"(f0.z), "r"(f0.w), "r"(f1.x), "r"(f1.y), "r"(f1.w),            /* %9..%15  */
          "r"(w0.x), "r"(w0.y), "r"(w0.z), "r"(w0.w), "r"(w1.x), "r"(w1.y), "r"(w1.z), "r"(w1.w), /* %16..%23 */
          "r"(f2.x), "r"(f2.y), "r"(f2.z), "r"(f3.x), "r"(f3.y), "r"(f3.z), "r"(f3.w)             /* %24..%30 */);
}

/*
 * One 32-row fragment chunk (phase A) in three stages: the straight-line chain, the rare-row fix-ups (entered
 * by one vote), the addends.  Rows beyond the site's last one are zero.  (Running the stages of two sites'
 * chunks side by side -- two independent chains per lane -- was measured: 1.73 ms vs 1.60 ms per 1M sites, the
 * extra registers cost more than the interleaving hides.)
 */
struct CRow { double hA, hB, wref, walt, pmA, pmB; int tie; unsigned vm, nm; bool special; };

/* stage 1: per-row loads (window constants of the row's library, prob_mapq of both reads) + the predicate chain;
 * `gen`: take the general chain (any site); otherwise the one-contig chain (F.fast == 1 sites only) */
__device__ __forceinline__ double ld_pm(const unsigned pm_addr, const unsigned byte_off)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(pm_addr + byte_off));
    return v;
}

__device__ __forceinline__ void crow_stage1(const CSiteF &F, const WinF *wf, const unsigned pm_addr, const int4 r, const bool gen,
                                            CRow &st)
{
    const int4 f0 = *reinterpret_cast<const int4 *>(&F.wA0);
    const int4 f1 = *reinterpret_cast<const int4 *>(&F.pat);
    const unsigned z = (unsigned)r.w;
    const unsigned lib = min((z >> 16) & kLibMaskC, (unsigned)kWLibs);
    const uint4 w0 = *reinterpret_cast<const uint4 *>(&wf[lib].altA_lo);
    const uint4 w1 = *reinterpret_cast<const uint4 *>(&wf[lib].FL);
    st.pmA = ld_pm(pm_addr, (z << 3) & 0x7F8u);       /* prob_mapq of both reads: byte offsets (mapq * 8) out of the packed word */
    st.pmB = ld_pm(pm_addr, (z >> 5) & 0x7F8u);
    if (!gen) crow_fast(r, f0, f1, w0, w1, st.hA, st.hB, st.wref, st.walt, st.tie);
    else crow_gen(r, f0, f1, *reinterpret_cast<const int4 *>(&F.sgnA), *reinterpret_cast<const uint4 *>(&F.mAA), w0, w1, st.hA,
                  st.hB, st.wref, st.walt, st.tie);
    st.vm = 0u; st.nm = 0u; st.special = false;
}

__device__ __forceinline__ bool crow_is_rare(const int4 r, const CRow &st)
{
    return ((unsigned)r.w & (CF_MULTI_A | CF_MULTI_B | CF_CONT)) != 0u || st.tie != 0;
}

/* stage 2 (entered by the whole warp when any lane's row is rare): MULTI rows take the packer's is_ref_seq
 * verdict, CONT rows (extra primaries) mark who starts a fragment, a p_concordant tie or a library outside the
 * integer rewrites re-scores the row the literal way */
__device__ __forceinline__ void crow_stage2(const SvgtParams &p, const Tables &t, const SiteS &S, const CSiteF &F,
                                            const LibK *s_lib, const int lane, const int n, const int m, const int4 r,
                                            CRow &st, int &err)
{
    const unsigned full = 0xffffffffu;
    const unsigned z = (unsigned)r.w;
    if (z & CF_MULTI_A) st.hA = ((unsigned)r.z & 1u) ? 1.0 : 0.0;
    if (z & CF_MULTI_B) st.hB = (((unsigned)r.z >> kLenBits) & 1u) ? 1.0 : 0.0;
    if (__any_sync(full, (z & CF_CONT) != 0u)) {        /* who starts a fragment */
        st.vm = n >= 32 ? full : ((1u << n) - 1u);
        st.nm = __ballot_sync(full, lane < n && !(z & CF_CONT));
        st.special = st.nm != st.vm;
    }
    if (st.tie != 0 && (z & CF_PAIRED)) {                /* the literal row */
        int4 lo, hi;
        decode_wide(r, S.tB, lo, hi);
        bool alt, refA, refB, pc;
        slow_row(p, t, S, lo, hi, s_lib, m, err, alt, refA, refB, pc);
        const bool is_del = F.del != 0;
        const bool both = refA & refB;
        const bool ref_on = (refA | refB) & (!both | is_del) & pc;
        const bool alt_on = alt & !(is_del & pc);
        st.wref = ref_on ? (both ? 1.0 : 0.5) : 0.0;
        st.walt = alt_on ? 1.0 : 0.0;
    }
}

/* stage 3: the addends (singlesample.py:254-259: a = pm[A] if read A covers a breakend; :305-350: p_alt, p_ref),
 * continuation rows folded into their fragment's last row of the chunk (see the note above FragOut) */
template <int ASSOC, bool RARE>
__device__ __forceinline__ FragOut crow_stage3(const int lane, const int n, const int4 r, const CRow &st)
{
    const unsigned full = 0xffffffffu;
    const double prod = __dmul_rn(st.pmA, st.pmB);
    const double vb = __dmul_rn(st.pmB, st.hB);
    FragOut o;
    o.s = __fma_rn(st.pmA, st.hA, vb);                 /* pmA * {0,1} is exact: one rounding, a + b */
    o.p_ref = __dmul_rn(prod, st.wref); o.p_alt = __dmul_rn(prod, st.walt);
    o.lead = 0;
    if (!RARE) return o;                                /* no vote: the chunk has no continuation rows to fold */
    if (ASSOC == SVGT_ASSOC_CLASSIC) {
        /* phase B adds a and b one by one: the caller parks their LUT indices (crow_lut_pair) */
    } else if (st.special && ((st.vm & ~st.nm) & ~(st.nm << 1)) == 0u) {
        /* every continuation row sits right below the row that starts its fragment (and none leads the chunk):
         * one fold step -- the lower row takes (s_up + a) + b and the weights, the upper row parks zeros */
        const unsigned NN = st.vm & ~st.nm;
        const bool cont = (NN >> lane) & 1u, has_next = (NN >> 1 >> lane) & 1u;
        const double us = __shfl_up_sync(full, o.s, 1), ur = __shfl_up_sync(full, o.p_ref, 1);
        const double ua = __shfl_up_sync(full, o.p_alt, 1);
        if (cont) {
            o.s = __dadd_rn(__dadd_rn(us, __dmul_rn(st.pmA, st.hA)), vb);
            o.p_ref = __dadd_rn(ur, o.p_ref); o.p_alt = __dadd_rn(ua, o.p_alt);
        }
        if (has_next) { o.s = 0.0; o.p_ref = 0.0; o.p_alt = 0.0; }
    } else if (st.special) {
        const FoldOut q = fold_continuations(lane, n, st.nm, st.vm, __dmul_rn(st.pmA, st.hA), vb, o.s, o.p_ref, o.p_alt);
        o.s = q.s; o.p_ref = q.p_ref; o.p_alt = q.p_alt; o.lead = q.lead;
    }
    return o;
}

/* the prob_mapq LUT indices of a row's own ref_seq addends a, b (0 = none), packed as a double: what phase B reads
 * where it must add a and b apart (every row under the classic association, `lead` rows under the sso one) */
__device__ __forceinline__ double crow_lut_pair(const int4 r, const CRow &st)
{
    const unsigned z = (unsigned)r.w;
    const int ia = st.hA != 0.0 ? (int)(z & 0xFFu) : 0, ib = st.hB != 0.0 ? (int)((z >> 8) & 0xFFu) : 0;
    return __hiloint2double(ib, ia);
}

/* ---- split rows (parsers.py:1122-1215, singlesample.py:262-274), pre-digested per site ---- */
struct CSplitF {
    int loL, loR, w1, kind;          /* lo = pos - slop of the left / right breakend; w1 = 2 * slop + 1;
                                        kind of a SOFT-CLIPPED row: 0 as a plain one (DEL), 1 DUP, 2 INV, 3 none */
    unsigned mLL, mLR, mRL, mRR;     /* class bit: left piece on the left / right breakend's contig, right piece .. */
    int rL, rR, pad0, pad1;
};

__device__ __forceinline__ CSplitF make_csplitf(const SiteS &S, int slop)
{
    CSplitF f;
    const int o1 = (S.meta >> 2) & 1, o2 = (S.meta >> 3) & 1, svtype = S.meta & 3;
    const bool swap = (S.tA != S.tB) || (S.posA > S.posB);      /* parsers.py:1143-1161 */
    f.loL = (swap ? S.posB : S.posA) - slop; f.loR = (swap ? S.posA : S.posB) - slop;
    f.w1 = 2 * slop + 1;
    f.rL = swap ? o2 : o1; f.rR = swap ? o1 : o2;
    f.kind = svtype == SV_DEL ? 0 : svtype == SV_DUP ? 1 : svtype == SV_INV ? 2 : 3;
    f.mLL = swap ? (1u << 29) : (1u << 28); f.mLR = swap ? (1u << 28) : (1u << 29);
    f.mRL = swap ? (1u << 31) : (1u << 30); f.mRR = swap ? (1u << 30) : (1u << 31);
    f.pad0 = 0; f.pad1 = 0;
    return f;
}

template <int ASSOC>
__device__ __forceinline__ SplitOut score_csplit_chunk(const CSplitF &F, const unsigned pm_addr, const int lane, const int n,
                                                       const int4 r)
{
    const unsigned full = 0xffffffffu;
    const int4 f0 = *reinterpret_cast<const int4 *>(&F.loL);
    const uint4 f1 = *reinterpret_cast<const uint4 *>(&F.mLL);
    const int2 f2 = *reinterpret_cast<const int2 *>(&F.rL);
    const unsigned z = (unsigned)r.w;
    const bool soft = (z & CSP_SOFT) != 0u;
    const bool first = (z & CSP_FIRST) != 0u;
    int l_end = r.x + (int)((unsigned)r.z & 0xFFFFu), r_end = r.y + (int)((unsigned)r.z >> 16);
    if (__any_sync(full, (z & CSP_WIDE) != 0u)) {          /* true ends ride in the next (XEND) row */
        const int xe = __shfl_down_sync(full, r.x, 1), ye = __shfl_down_sync(full, r.y, 1);
        if (z & CSP_WIDE) { l_end = xe; r_end = ye; }
    }
    const unsigned w1 = (unsigned)f0.z;
    const int cl = f2.x ? r.x : l_end, cr = f2.y ? r.x : l_end;       /* left piece vs L / R side */
    const int dl = f2.x ? r.y : r_end, dr = f2.y ? r.y : r_end;       /* right piece vs L / R side */
    const bool lL = ((z & f1.x) != 0u) & ((unsigned)(cl - f0.x) < w1);
    const bool lR = ((z & f1.y) != 0u) & ((unsigned)(cr - f0.y) < w1);
    const bool rLs = ((z & f1.z) != 0u) & ((unsigned)(dl - f0.x) < w1);
    const bool rRs = ((z & f1.w) != 0u) & ((unsigned)(dr - f0.y) < w1);
    const int kind = soft ? f0.w : 0;
    const bool Ls = kind == 0 ? lL : kind == 1 ? lR : kind == 2 ? (lL | lR) : false;
    const bool Rs = kind == 0 ? rRs : kind == 1 ? rLs : kind == 2 ? (rLs | rRs) : false;
    const double x = ld_pm(pm_addr, Ls ? ((z << 3) & 0x7F8u) : 0u);
    const double y = ld_pm(pm_addr, Rs ? ((z >> 5) & 0x7F8u) : 0u);
    const double p_alt = __dmul_rn(__dadd_rn(x, y), 0.5);       /* (.. + ..) / 2.0, exact either way */
    SplitOut o;
    o.vseq = soft ? 0.0 : p_alt; o.vclip = soft ? p_alt : 0.0; o.lead = 0;
    const bool rv = lane < n;
    if (ASSOC == SVGT_ASSOC_SSO && __any_sync(full, rv && !first)) {     /* extra splits of one fragment */
        const unsigned vm = n >= 32 ? full : ((1u << n) - 1u);
        const unsigned nm = __ballot_sync(full, rv && first);
        const unsigned NN = vm & ~nm;
        if ((NN & ~(nm << 1)) == 0u) {
            const bool cont = (NN >> lane) & 1u, has_next = (NN >> 1 >> lane) & 1u;
            const double us = __shfl_up_sync(full, o.vseq, 1), uc = __shfl_up_sync(full, o.vclip, 1);
            if (cont) { o.vseq = __dadd_rn(us, o.vseq); o.vclip = __dadd_rn(uc, o.vclip); }
            if (has_next) { o.vseq = 0.0; o.vclip = 0.0; }
        } else {
            const FoldOut q = fold_splits(lane, n, nm, o.vseq, o.vclip);
            o.vseq = q.p_ref; o.vclip = q.p_alt; o.lead = q.lead;
        }
    }
    return o;
}

}  // namespace
