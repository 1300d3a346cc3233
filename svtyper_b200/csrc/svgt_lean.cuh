/*
 * svgt_lean.cuh -- the lean row scorer of the default tally kernel (svgt_lean.cu, variant 5).
 *
 * Same arithmetic as score_frag_chunk() in svgt_coop.cuh (reference citations there); what changes
 * is how much work a row costs.  The cooperative kernel is bound by instruction issue and by
 * shared-memory wavefronts, not by HBM (profiles/README.md), so for the sites that make up
 * practically every real batch -- both breakends on one contig, both is_ref_seq windows valid, not
 * an inversion: a "fast site" -- the row is scored by one straight-line predicate chain:
 *
 *   - site constants are two warp-uniform LDS.128 (SiteF), pre-digested once per site;
 *   - the per-(site, library) state is two LDS.128 (WinF): the two alt windows as (lo, width+1),
 *     FL and FL+1 for the four reference windows (they all have width FL+1 and are anchored on the
 *     site's wA0/wA1/wB0/wB1, so x = a_start + FL serves both a-side tests), the second histogram
 *     key and the packed shared-memory address + length of the library's histogram;
 *   - histogram look-ups are index-clamped onto a zero sentinel behind each library's counts
 *     instead of predicated (min.u32 replaces setp + mov 0);
 *   - weights are applied on the otherwise idle fp64 pipe: p_ref = (pmA * pmB) * {0, 0.5, 1},
 *     p_alt = (pmA * pmB) * {0, 1}, a + b = fma(pmA, {0,1}, pmB * {0,1}) -- every one of these is
 *     exact where the reference selects or halves, so results stay bit-identical -- which removes
 *     four of six LUT reads and the 64-bit selects;
 *   - a library that cannot take the integer rewrites (or an index beyond the cached four) maps to
 *     a trap entry: the lane reports it with the p_concordant tie flag and the literal fp64 row
 *     (slow_row) recomputes its weights.
 * EXTRA / MULTI / CONT rows (gapped reads, extra primaries; evidence.py) are resolved by the same
 * ballots as in the cooperative kernel, entered only when a chunk has such rows.
 */
#pragma once
#include "svgt_coop.cuh"

namespace {

constexpr unsigned kOneHi = 0x3FF00000u, kHalfHi = 0x3FE00000u;   /* high words of 1.0 and 0.5 */
constexpr unsigned kTrapLk = 0x80000000u;
constexpr int kHistLenBits = 14;                                  /* packed hist length < 16384 */

/* fast-site constants, read warp-uniformly as two LDS.128 (three for the general chain) */
struct SiteF {
    int tA, wA0, wA1, wB0;
    int wB1, pat, del, fast;     /* pat: (fl & 0x5C) of an alt-orientation pair; del: DEL site;
                                    fast: 0 generic scorer, 1 fast_row (one contig, not INV), 2 fast_row_gen */
    int tB, sgnA, sgnB, inv;     /* general chain: second contig; the reciprocal (INV) alt windows are the alt
                                    windows shifted by sgn * FL; inv: the site is an inversion */
};

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

/*
 * One row of a fast site.  Outputs: high words of the {0,1} factors of a and b, of the {0,.5,1}
 * factor of p_ref and the {0,1} factor of p_alt, and `tie` (p_concordant tie or trap library:
 * the caller re-scores the row the literal way).
 *   hits    parsers.py:801-816       is_ref_seq of either read against either breakend window
 *   pa      parsers.py:821-857       is_pair_straddle(alt) via the (site, library) windows
 *   ra, rb  same, reference FR pairs at breakend A / B
 *   pc      parsers.py:861-882       p_concordant > 0.5  <=>  19 * h1 > h2 on the counts
 *   weights singlesample.py:305-318, :336-350
 */
__device__ __forceinline__ void fast_row(const int4 lo, const int4 hi, const int4 f0, const int4 f1, const uint4 w0,
                                         const uint4 w1, double &hA, double &hB, double &wref, double &walt, int &tie)
{
    asm("{\n\t"
        ".reg .pred e1, e2, pfa, pfb, p, q, hA, hB, pe0, pa, pr0, ra, rb, pc, pt, pdel, both, any, x1, ron, aon, ptrap;\n\t"
        ".reg .b32 t, d, x, o, k2, len, hb, i1, i2, a1, a2, h1, h2, l19, zr;\n\t"
        "mov.b32 zr, 0;\n\t"
        /* tids and presence flags */
        "setp.eq.s32 e1, %9, %12;\n\t"
        "setp.eq.s32 e2, %10, %12;\n\t"
        "and.b32 t, %11, 1;\n\t"
        "setp.ne.and.s32 pfa, t, 0, e1;\n\t"
        "and.b32 t, %11, 2;\n\t"
        "setp.ne.and.s32 pfb, t, 0, e2;\n\t"
        /* is_ref_seq, read A then read B */
        "setp.le.and.s32 p, %5, %13, pfa;\n\t"
        "setp.ge.and.s32 p, %6, %14, p;\n\t"
        "setp.le.and.s32 q, %5, %15, pfa;\n\t"
        "setp.ge.and.s32 q, %6, %16, q;\n\t"
        "or.pred hA, p, q;\n\t"
        "setp.le.and.s32 p, %7, %13, pfb;\n\t"
        "setp.ge.and.s32 p, %8, %14, p;\n\t"
        "setp.le.and.s32 q, %7, %15, pfb;\n\t"
        "setp.ge.and.s32 q, %8, %16, q;\n\t"
        "or.pred hB, p, q;\n\t"
        "selp.b32 t, 0x3FF00000, 0, hA;\n\t"
        "mov.b64 %0, {zr, t};\n\t"
        "selp.b32 t, 0x3FF00000, 0, hB;\n\t"
        "mov.b64 %1, {zr, t};\n\t"
        /* paired-end straddles */
        "and.pred pe0, e1, e2;\n\t"
        "and.b32 t, %11, 0x5C;\n\t"
        "setp.eq.and.s32 pa, t, %17, pe0;\n\t"
        "setp.eq.and.s32 pr0, t, 0x18, pe0;\n\t"
        "sub.s32 d, %5, %19;\n\t"
        "setp.lt.and.u32 pa, d, %20, pa;\n\t"
        "sub.s32 d, %8, %21;\n\t"
        "setp.lt.and.u32 pa, d, %22, pa;\n\t"
        "add.s32 x, %5, %23;\n\t"
        "sub.s32 d, x, %13;\n\t"
        "setp.lt.and.u32 ra, d, %24, pr0;\n\t"
        "sub.s32 d, %8, %14;\n\t"
        "add.s32 d, d, -1;\n\t"
        "setp.lt.and.u32 ra, d, %24, ra;\n\t"
        "sub.s32 d, x, %15;\n\t"
        "setp.lt.and.u32 rb, d, %24, pr0;\n\t"
        "sub.s32 d, %8, %16;\n\t"
        "add.s32 d, d, -1;\n\t"
        "setp.lt.and.u32 rb, d, %24, rb;\n\t"
        /* p_concordant on the counts; keys outside the histogram clamp onto the zero sentinel */
        "sad.s32 o, %8, %5, 0;\n\t"
        "sub.s32 k2, o, %25;\n\t"
        "shr.u32 len, %26, 18;\n\t"
        "and.b32 hb, %26, 0x3ffff;\n\t"
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
        "setp.lt.s32 ptrap, %25, 0;\n\t"
        "or.pred pt, pt, ptrap;\n\t"
        "selp.s32 %4, 1, 0, pt;\n\t"
        /* weights */
        "setp.ne.s32 pdel, %18, 0;\n\t"
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
        : "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.w),          /* %5..%11  */
          "r"(f0.x), "r"(f0.y), "r"(f0.z), "r"(f0.w), "r"(f1.x), "r"(f1.y), "r"(f1.z),          /* %12..%18 */
          "r"(w0.x), "r"(w0.y), "r"(w0.z), "r"(w0.w), "r"(w1.x), "r"(w1.y), "r"(w1.z), "r"(w1.w) /* %19..%26 */);
}

/*
 * The same row for a site whose breakends lie on two contigs (BND) and / or an inversion: every test
 * names its own contig, and an INV pair may also straddle in the reciprocal orientation
 * (singlesample.py:296-303), whose windows are the alt windows shifted by -/+ FL.
 */
/* (inline on purpose: an out-of-line call here costs the one-contig loop 15 % -- measured -- through the
 * registers it pins across the call) */
__device__ __forceinline__
void fast_row_gen(const int4 lo, const int4 hi, const int4 f0, const int4 f1, const int4 f2,
                                             const uint4 w0, const uint4 w1, double &hA, double &hB, double &wref,
                                             double &walt, int &tie)
{
    asm("{\n\t"
        ".reg .pred eaa, eab, eba, ebb, pfl, pf1, pf2, p, q, hA, hB, pe0, pa, prc, pr0, ra, rb, pc, pt, pdel, pinv, both, any, x1, ron, aon, ptrap;\n\t"
        ".reg .b32 t, u, d, x, o, k2, len, hb, i1, i2, a1, a2, h1, h2, l19, zr;\n\t"
        "mov.b32 zr, 0;\n\t"
        "setp.eq.s32 eaa, %9, %12;\n\t"
        "setp.eq.s32 eab, %9, %27;\n\t"
        "setp.eq.s32 eba, %10, %12;\n\t"
        "setp.eq.s32 ebb, %10, %27;\n\t"
        /* is_ref_seq, read A: on contig tA against window A, on contig tB against window B */
        "and.b32 t, %11, 1;\n\t"
        "setp.ne.s32 pfl, t, 0;\n\t"
        "and.pred pf1, pfl, eaa;\n\t"
        "and.pred pf2, pfl, eab;\n\t"
        "setp.le.and.s32 p, %5, %13, pf1;\n\t"
        "setp.ge.and.s32 p, %6, %14, p;\n\t"
        "setp.le.and.s32 q, %5, %15, pf2;\n\t"
        "setp.ge.and.s32 q, %6, %16, q;\n\t"
        "or.pred hA, p, q;\n\t"
        "and.b32 t, %11, 2;\n\t"
        "setp.ne.s32 pfl, t, 0;\n\t"
        "and.pred pf1, pfl, eba;\n\t"
        "and.pred pf2, pfl, ebb;\n\t"
        "setp.le.and.s32 p, %7, %13, pf1;\n\t"
        "setp.ge.and.s32 p, %8, %14, p;\n\t"
        "setp.le.and.s32 q, %7, %15, pf2;\n\t"
        "setp.ge.and.s32 q, %8, %16, q;\n\t"
        "or.pred hB, p, q;\n\t"
        "selp.b32 t, 0x3FF00000, 0, hA;\n\t"
        "mov.b64 %0, {zr, t};\n\t"
        "selp.b32 t, 0x3FF00000, 0, hB;\n\t"
        "mov.b64 %1, {zr, t};\n\t"
        /* alt straddle: read A on tA, read B on tB, alt orientation; INV also the reciprocal one */
        "and.pred pe0, eaa, ebb;\n\t"
        "and.b32 t, %11, 0x5C;\n\t"
        "setp.eq.and.s32 pa, t, %17, pe0;\n\t"
        "sub.s32 d, %5, %19;\n\t"
        "setp.lt.and.u32 pa, d, %20, pa;\n\t"
        "sub.s32 d, %8, %21;\n\t"
        "setp.lt.and.u32 pa, d, %22, pa;\n\t"
        "setp.ne.s32 pinv, %30, 0;\n\t"
        "and.pred prc, pe0, pinv;\n\t"
        "xor.b32 u, %17, 0xC;\n\t"
        "setp.eq.and.s32 prc, t, u, prc;\n\t"
        "sub.s32 d, %5, %19;\n\t"
        "mad.lo.s32 d, %28, %23, d;\n\t"
        "setp.lt.and.u32 prc, d, %20, prc;\n\t"
        "sub.s32 d, %8, %21;\n\t"
        "mad.lo.s32 d, %29, %23, d;\n\t"
        "setp.lt.and.u32 prc, d, %22, prc;\n\t"
        "or.pred pa, pa, prc;\n\t"
        /* reference FR pairs: both reads on tA around A, both on tB around B */
        "setp.eq.s32 pr0, t, 0x18;\n\t"
        "and.pred ra, eaa, eba;\n\t"
        "and.pred ra, ra, pr0;\n\t"
        "and.pred rb, eab, ebb;\n\t"
        "and.pred rb, rb, pr0;\n\t"
        "add.s32 x, %5, %23;\n\t"
        "sub.s32 d, x, %13;\n\t"
        "setp.lt.and.u32 ra, d, %24, ra;\n\t"
        "sub.s32 d, %8, %14;\n\t"
        "add.s32 d, d, -1;\n\t"
        "setp.lt.and.u32 ra, d, %24, ra;\n\t"
        "sub.s32 d, x, %15;\n\t"
        "setp.lt.and.u32 rb, d, %24, rb;\n\t"
        "sub.s32 d, %8, %16;\n\t"
        "add.s32 d, d, -1;\n\t"
        "setp.lt.and.u32 rb, d, %24, rb;\n\t"
        /* p_concordant on the counts; keys outside the histogram clamp onto the zero sentinel */
        "sad.s32 o, %8, %5, 0;\n\t"
        "sub.s32 k2, o, %25;\n\t"
        "shr.u32 len, %26, 18;\n\t"
        "and.b32 hb, %26, 0x3ffff;\n\t"
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
        "setp.lt.s32 ptrap, %25, 0;\n\t"
        "or.pred pt, pt, ptrap;\n\t"
        "selp.s32 %4, 1, 0, pt;\n\t"
        /* weights */
        "setp.ne.s32 pdel, %18, 0;\n\t"
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
        : "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.w),          /* %5..%11  */
          "r"(f0.x), "r"(f0.y), "r"(f0.z), "r"(f0.w), "r"(f1.x), "r"(f1.y), "r"(f1.z),          /* %12..%18 */
          "r"(w0.x), "r"(w0.y), "r"(w0.z), "r"(w0.w), "r"(w1.x), "r"(w1.y), "r"(w1.z), "r"(w1.w), /* %19..%26 */
          "r"(f2.x), "r"(f2.y), "r"(f2.z), "r"(f2.w)                                             /* %27..%30 */);
}

/* EXTRA interval rows feed the next main row's MULTI slots; hits of an EXTRA run that reaches the end of
 * the chunk are carried into the site's next chunk (bit g of carryA / carryB).  Same ballots as in
 * score_frag_chunk(); out of line because < 1 % of rows are such rows. */
struct MultiOut { unsigned hAhi, hBhi, carryA, carryB, nm, vm; };

__device__ __noinline__ MultiOut resolve_multi(const int fl, const int lane, const int n, const int g, unsigned hAhi,
                                               unsigned hBhi, unsigned carryA, unsigned carryB)
{
    const unsigned full = 0xffffffffu;
    const bool rv = lane < n;
    const unsigned vm = n >= 32 ? full : ((1u << n) - 1u);
    const bool isx = (fl & F_EXTRA) != 0;
    bool hitA = hAhi != 0u, hitB = hBhi != 0u;
    const bool cA = (carryA >> g) & 1u, cB = (carryB >> g) & 1u;
    const unsigned E = __ballot_sync(full, isx);
    const unsigned HA = __ballot_sync(full, isx && hitA), HB = __ballot_sync(full, isx && hitB);
    const unsigned below = (1u << lane) - 1u;
    const unsigned zz = ~E & below;
    unsigned runm;
    bool reach0;
    if (zz == 0u) { runm = below; reach0 = true; }
    else { const int pz = 31 - __clz(zz); runm = below & ~((2u << pz) - 1u); reach0 = false; }
    const bool pA = ((HA & runm) != 0u) || (reach0 && cA);
    const bool pB = ((HB & runm) != 0u) || (reach0 && cB);
    if (fl & F_MULTI_A) hitA = pA;
    if (fl & F_MULTI_B) hitB = pB;
    const unsigned zt = ~E & vm;
    bool nA, nB;
    if (zt == 0u) { nA = cA || (HA != 0u); nB = cB || (HB != 0u); }
    else {
        const int pz = 31 - __clz(zt);
        const unsigned rt = vm & ~((2u << pz) - 1u);
        nA = (HA & rt) != 0u; nB = (HB & rt) != 0u;
    }
    MultiOut r;
    r.carryA = (carryA & ~(1u << g)) | ((unsigned)nA << g);
    r.carryB = (carryB & ~(1u << g)) | ((unsigned)nB << g);
    r.nm = __ballot_sync(full, rv && !(fl & (F_CONT | F_EXTRA)));
    r.vm = vm;
    if (isx) { hitA = false; hitB = false; }
    r.hAhi = hitA ? kOneHi : 0u; r.hBhi = hitB ? kOneHi : 0u;
    return r;
}

/* sso association of a fragment's extra rows (singlesample.py:254-259, :367-378): see score_frag_chunk() */
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

/* one 32-row fragment chunk of a FAST site (phase A); rows beyond the site's last one were zero-filled */
template <int ASSOC>
__device__ __forceinline__ FragOut score_frag_chunk_fast(const SvgtParams &p, const Tables &t, const SiteS &S,
                                                         const SiteF &F, const WinF *wf, const double *s_pm,
                                                         const LibK *s_lib, const int lane, const int step, const int g,
                                                         const int m, const int4 lo, const int4 hi, unsigned &carryA,
                                                         unsigned &carryB, int &err)
{
    const unsigned full = 0xffffffffu;
    const int4 f0 = *reinterpret_cast<const int4 *>(&F.tA);
    const int4 f1 = *reinterpret_cast<const int4 *>(&F.wB1);
    const int fl = hi.w;
    const unsigned z = (unsigned)hi.z;
    const unsigned lib = min(z >> 16, (unsigned)kWLibs);
    const uint4 w0 = *reinterpret_cast<const uint4 *>(&wf[lib].altA_lo);
    const uint4 w1 = *reinterpret_cast<const uint4 *>(&wf[lib].FL);
    /* prob_mapq of both reads: byte offsets (mapq * 8) straight out of the packed word */
    const char *pmb = reinterpret_cast<const char *>(s_pm);
    const double pmA = *reinterpret_cast<const double *>(pmb + ((z << 3) & 0x7F8u));
    const double pmB = *reinterpret_cast<const double *>(pmb + ((z >> 5) & 0x7F8u));

    double hA, hB, wref, walt;                      /* {0,1}, {0,1}, {0,.5,1}, {0,1} */
    int tie;
    if (f1.w == 1) fast_row(lo, hi, f0, f1, w0, w1, hA, hB, wref, walt, tie);
    else fast_row_gen(lo, hi, f0, f1, *reinterpret_cast<const int4 *>(&F.tB), w0, w1, hA, hB, wref, walt, tie);

    /* one vote for everything rare: EXTRA / MULTI / CONT rows (evidence.py), a p_concordant tie or a
     * library the integer rewrites do not cover; carried EXTRA hits re-enter through the carry bits */
    unsigned vm = 0u, nm = 0u;
    bool special = false;                                       /* the chunk has CONT / EXTRA rows (warp-uniform) */
    const bool xrow = (fl & (F_EXTRA | F_MULTI_A | F_MULTI_B | F_CONT)) != 0;
    if (__any_sync(full, xrow || tie != 0) || (carryA | carryB) != 0u) {
        if (__any_sync(full, (fl & (F_EXTRA | F_MULTI_A | F_MULTI_B)) != 0) || (((carryA | carryB) >> g) & 1u)) {
            const MultiOut r = resolve_multi(fl, lane, S.nf - step * 32, g, (unsigned)__double2hiint(hA),
                                             (unsigned)__double2hiint(hB), carryA, carryB);
            hA = __hiloint2double((int)r.hAhi, 0); hB = __hiloint2double((int)r.hBhi, 0);
            carryA = r.carryA; carryB = r.carryB; nm = r.nm; vm = r.vm;
            special = nm != vm;
        } else if (__any_sync(full, xrow)) {            /* CONT rows only: no hits to move, just who starts a fragment */
            const int n = S.nf - step * 32;
            vm = n >= 32 ? full : ((1u << n) - 1u);
            nm = __ballot_sync(full, lane < n && !(fl & F_CONT));
            special = nm != vm;
        }
        if (tie != 0 && (fl & (F_PAIRED | F_EXTRA)) == F_PAIRED) {      /* the literal row */
            bool alt, refA, refB, pc;
            slow_row(p, t, S, lo, hi, s_lib, m, err, alt, refA, refB, pc);
            const bool is_del = F.del != 0;
            const bool both = refA & refB;
            const bool ref_on = (refA | refB) & (!both | is_del) & pc;
            const bool alt_on = alt & !(is_del & pc);
            wref = ref_on ? (both ? 1.0 : 0.5) : 0.0;
            walt = alt_on ? 1.0 : 0.0;
        }
    }

    /* singlesample.py:254-259: a = pm[A] if read A covers a breakend; :305-350: p_alt, p_ref */
    const double prod = __dmul_rn(pmA, pmB);
    const double vb = __dmul_rn(pmB, hB);
    FragOut o;
    o.s = __fma_rn(pmA, hA, vb);                       /* pmA * {0,1} is exact: one rounding, a + b */
    o.p_ref = __dmul_rn(prod, wref); o.p_alt = __dmul_rn(prod, walt);
    o.ia = 0; o.ib = 0; o.lead = 0; o.need_idx = false;
    if (ASSOC == SVGT_ASSOC_CLASSIC) {                  /* phase B adds a and b one by one: park their LUT indices */
        o.ia = hA != 0.0 ? (int)(z & 0xFFu) : 0; o.ib = hB != 0.0 ? (int)((z >> 8) & 0xFFu) : 0;
    } else if (special && ((vm & ~nm) & ~(nm << 1)) == 0u) {
        /* every continuation row sits right below the row that starts its fragment (and none leads the
         * chunk): fold_continuations() is one step -- the lower row takes (s_up + a) + b and the weights,
         * the upper row parks zeros.  Folding weights that are zero anyway changes nothing (x + 0.0). */
        const unsigned NN = vm & ~nm;
        const bool cont = (NN >> lane) & 1u, has_next = (NN >> 1 >> lane) & 1u;
        const double us = __shfl_up_sync(full, o.s, 1), ur = __shfl_up_sync(full, o.p_ref, 1);
        const double ua = __shfl_up_sync(full, o.p_alt, 1);
        if (cont) {
            o.s = __dadd_rn(__dadd_rn(us, __dmul_rn(pmA, hA)), vb);
            o.p_ref = __dadd_rn(ur, o.p_ref); o.p_alt = __dadd_rn(ua, o.p_alt);
        }
        if (has_next) { o.s = 0.0; o.p_ref = 0.0; o.p_alt = 0.0; }
    } else if (special) {
        const FoldOut r = fold_continuations(lane, S.nf - step * 32, nm, vm, __dmul_rn(pmA, hA), vb, o.s, o.p_ref, o.p_alt);
        o.s = r.s; o.p_ref = r.p_ref; o.p_alt = r.p_alt; o.lead = r.lead;
        if (lane < r.lead) { o.ia = hA != 0.0 ? (int)(z & 0xFFu) : 0; o.ib = hB != 0.0 ? (int)((z >> 8) & 0xFFu) : 0; }
    }
    return o;
}

/* ---- split rows of any site (parsers.py:1122-1215, singlesample.py:262-274), pre-digested per site ---- */
struct SplitF {
    int tL, tR, loL, loR;        /* breakends left to right (parsers.py:1143-1161); lo = pos - slop */
    int w1, rL, rR, kind;        /* w1 = 2 * slop + 1; kind of a SOFT-CLIPPED row: 0 as a plain one (DEL), 1 DUP, 2 INV, 3 none */
};

__device__ __forceinline__ SplitF make_splitf(const SiteS &S, int slop)
{
    SplitF f;
    const int o1 = (S.meta >> 2) & 1, o2 = (S.meta >> 3) & 1, svtype = S.meta & 3;
    const bool swap = (S.tA != S.tB) || (S.posA > S.posB);
    f.tL = swap ? S.tB : S.tA; f.tR = swap ? S.tA : S.tB;
    f.loL = (swap ? S.posB : S.posA) - slop; f.loR = (swap ? S.posA : S.posB) - slop;
    f.w1 = 2 * slop + 1;
    f.rL = swap ? o2 : o1; f.rR = swap ? o1 : o2;
    f.kind = svtype == SV_DEL ? 0 : svtype == SV_DUP ? 1 : svtype == SV_INV ? 2 : 3;
    return f;
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

/* one 32-row split chunk: parks {alt_seq, alt_clip} addends; same arithmetic as score_split_chunk() */
template <int ASSOC>
__device__ __forceinline__ SplitOut score_split_chunk_lean(const SplitF &F, const double *s_pm, const int lane, const int n,
                                                           const int4 q0, const int4 q1)
{
    const unsigned full = 0xffffffffu;
    const int4 f0 = *reinterpret_cast<const int4 *>(&F.tL);
    const int4 f1 = *reinterpret_cast<const int4 *>(&F.w1);
    const bool rv = lane < n;
    const unsigned z = (unsigned)q1.z;
    const bool soft = (z >> 16) & S_SOFT_CLIP;
    const bool first = (z >> 16) & S_FIRST;
    const unsigned w1 = (unsigned)f1.x;
    const int cl = f1.y ? q0.y : q0.z, cr = f1.z ? q0.y : q0.z;       /* left piece vs L / R side */
    const int dl = f1.y ? q1.x : q1.y, dr = f1.z ? q1.x : q1.y;       /* right piece vs L / R side */
    const bool lL = (q0.x == f0.x) & ((unsigned)(cl - f0.z) < w1);
    const bool lR = (q0.x == f0.y) & ((unsigned)(cr - f0.w) < w1);
    const bool rLs = (q0.w == f0.x) & ((unsigned)(dl - f0.z) < w1);
    const bool rRs = (q0.w == f0.y) & ((unsigned)(dr - f0.w) < w1);
    const int kind = soft ? f1.w : 0;
    const bool Ls = rv & (kind == 0 ? lL : kind == 1 ? lR : kind == 2 ? (lL | lR) : false);
    const bool Rs = rv & (kind == 0 ? rRs : kind == 1 ? rLs : kind == 2 ? (rLs | rRs) : false);
    const double x = s_pm[Ls ? (z & 0xFFu) : 0u];
    const double y = s_pm[Rs ? ((z >> 8) & 0xFFu) : 0u];
    const double p_alt = __dmul_rn(__dadd_rn(x, y), 0.5);       /* (.. + ..) / 2.0, exact either way */
    SplitOut o;
    o.vseq = soft ? 0.0 : p_alt; o.vclip = soft ? p_alt : 0.0; o.lead = 0;
    if (ASSOC == SVGT_ASSOC_SSO && __any_sync(full, rv && !first)) {     /* extra splits of one fragment */
        const unsigned vm = n >= 32 ? full : ((1u << n) - 1u);
        const unsigned nm = __ballot_sync(full, rv && first);
        const unsigned NN = vm & ~nm;
        if ((NN & ~(nm << 1)) == 0u) {                  /* second split right below its first: one fold step */
            const bool cont = (NN >> lane) & 1u, has_next = (NN >> 1 >> lane) & 1u;
            const double us = __shfl_up_sync(full, o.vseq, 1), uc = __shfl_up_sync(full, o.vclip, 1);
            if (cont) { o.vseq = __dadd_rn(us, o.vseq); o.vclip = __dadd_rn(uc, o.vclip); }
            if (has_next) { o.vseq = 0.0; o.vclip = 0.0; }
        } else {
            const FoldOut r = fold_splits(lane, n, nm, o.vseq, o.vclip);
            o.vseq = r.p_ref; o.vclip = r.p_alt; o.lead = r.lead;
        }
    }
    return o;
}

/*
 * Parked rows, structure-of-arrays per site: ch[0] = a + b (or alt_seq), ch[1] = p_ref (or alt_clip),
 * ch[2] = p_alt.  Where phase B needs a and b separately -- every row under the classic association,
 * the `lead` rows that continue the previous chunk's fragment under the sso one -- ch[0] holds the two
 * prob_mapq LUT indices (ib, ia) instead of the sum.  Phase A stores are 8-byte-stride
 * (conflict-free); chain c of site g starts at bank 6g + 2c (mod 32), so the 24 chain lanes of phase B
 * read in the minimum two wavefronts.
 */
struct alignas(8) Parked { double ch[3][33]; };

template <int ASSOC>
__device__ __forceinline__ void park_frag_soa(Parked &P, int lane, const FragOut &o)
{
    if (ASSOC == SVGT_ASSOC_CLASSIC) P.ch[0][lane] = __hiloint2double(o.ib, o.ia);
    else {
        P.ch[0][lane] = o.s;
        if (o.lead > 0) { if (lane < o.lead) P.ch[0][lane] = __hiloint2double(o.ib, o.ia); }   /* warp-uniform, rare */
    }
    P.ch[1][lane] = o.p_ref; P.ch[2][lane] = o.p_alt;
}

/* phase B of one fragment chunk for chain c (0 ref_seq, 1 ref_span, 2 alt_span) of one site; see replay_frag() */
template <int ASSOC>
__device__ __forceinline__ void replay_frag_soa(const Parked &P, int c, int cnt, int lead, const double *s_pm, double &acc,
                                                double &pend)
{
    const double *px = &P.ch[c][0];
    const int2 *pi = reinterpret_cast<const int2 *>(&P.ch[0][0]);      /* .x = ia, .y = ib */
    if (cnt <= 0) return;                               /* no chunk of this site in the super-step: `lead` is stale */
    lead = lead < cnt ? lead : cnt;
    if (ASSOC == SVGT_ASSOC_CLASSIC) {
        /* classic.py:306-311,339-408: every read goes straight into the site sum */
        if (c == 0) {
            for (int j = 0; j < cnt; ++j) {
                const int2 ix = pi[j];
                acc = __dadd_rn(__dadd_rn(acc, s_pm[ix.x]), s_pm[ix.y]);
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) acc = __dadd_rn(acc, px[j]);
        }
    } else {
        for (int j = 0; j < lead; ++j) {                /* rows continuing the previous chunk's last fragment */
            if (c == 0) {
                const int2 ix = pi[j];
                pend = __dadd_rn(__dadd_rn(pend, s_pm[ix.x]), s_pm[ix.y]);
            } else {
                pend = __dadd_rn(pend, px[j]);
            }
        }
#pragma unroll 8
        for (int j = lead; j < cnt; ++j) {
            acc = __dadd_rn(acc, pend);
            pend = px[j];
        }
    }
}

/* phase B of one split chunk for chain c (0 alt_seq, 1 alt_clip) */
template <int ASSOC>
__device__ __forceinline__ void replay_split_soa(const Parked &P, int c, int cnt, int lead, double &acc, double &pend)
{
    const double *px = &P.ch[c][0];
    if (cnt <= 0) return;
    lead = lead < cnt ? lead : cnt;
    if (ASSOC == SVGT_ASSOC_CLASSIC) {
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) acc = __dadd_rn(acc, px[j]);
    } else {
        for (int j = 0; j < lead; ++j) pend = __dadd_rn(pend, px[j]);
#pragma unroll 8
        for (int j = lead; j < cnt; ++j) {
            acc = __dadd_rn(acc, pend);
            pend = px[j];
        }
    }
}

}  // namespace
