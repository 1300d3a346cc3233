/*
 * svgt_coop.cuh -- pieces shared by the warp-cooperative tally kernels (svgt_coop.cu: rows through
 * registers; svgt_ring.cu: rows through a cp.async.bulk shared-memory ring): per-site scalars and
 * integer windows, the PTX predicate chains, the slow (literal) row path and the ordered chain replay.
 */
#pragma once
#include "svgt_device.cuh"


namespace {

#ifndef SVGT_ROWBUFS
#define SVGT_ROWBUFS 3
#endif
#ifndef SVGT_DIAG
#define SVGT_DIAG 0          /* diagnostics only (wrong results), bit mask: 1 no scoring, 2 no phase B, 4 no split phase,
                                8 no fragment-row loads, 16 no L2 prefetch */
#endif
#ifndef SVGT_SPLIT_PIPE
#define SVGT_SPLIT_PIPE 0
#endif
#ifndef SVGT_USE_SAME
#define SVGT_USE_SAME 1
#endif
#ifndef SVGT_SPLIT_ROT
#define SVGT_SPLIT_ROT 0
#endif
#ifndef SVGT_L2_PREFETCH
#define SVGT_L2_PREFETCH 0
#endif
constexpr int kWLibs = 4;           /* libraries with per-site windows cached in smem      */
constexpr int kCoopWarps = SVGT_COOP_THREADS / 32;

/* per-(warp, g) site scalars, read warp-uniformly.  First 32 bytes are the hot ones. */
struct SiteS {
    int tA, tB, wA0, wA1;
    int wB0, wB1, meta, var_length;
    int posA, posB, ciA0, ciA1;
    int ciB0, ciB1, dAB, nf;
    long long foff, soff;
    int ns, slot, pad1, pad2;
};  /* 96 B */

/* per-(warp, g, library) windows as (lo, width+1): pass iff (unsigned)(v - lo) < w1 */
struct Win {
    unsigned altA_lo, altA_w1, altB_lo, altB_w1;
    unsigned recA_lo, recA_w1, recB_lo, recB_w1;
    unsigned rAa_lo, rAa_w1, rAb_lo, rAb_w1;
    unsigned rBa_lo, rBa_w1, rBb_lo, rBb_w1;
    unsigned Lk, hist_off, hist_len, flags;   /* p_concordant inputs; flags bit 0 = integer fast path valid.
                                                 80 B stride: the four cached libraries land in distinct banks */
};

template <int G>
struct alignas(128) WarpSmem {
    SiteS site[G];
    Win win[G][kWLibs];
    double contrib[G][33][4];   /* 32 rows + 32 B pad: chain lanes of different sites hit distinct banks */
    unsigned newmask[G];
    double zero[2];
};

__device__ __forceinline__ void set_win(unsigned &lo_out, unsigned &w1_out, int lo, int hi, bool enable)
{
    lo_out = (unsigned)lo;
    w1_out = (enable && hi >= lo) ? (unsigned)(hi - lo) + 1u : 0u;
}

__device__ __forceinline__ Win make_win(const SiteS &S, const LibK &L, int m, bool small_counts)
{
    Win w;
    const int svtype = S.meta & 3;
    const bool is_del = svtype == SV_DEL;
    const int o1 = (S.meta >> 2) & 1, o2 = (S.meta >> 3) & 1;
    const bool ok = L.safe != 0;
    const bool small_del = is_del && (S.dAB < L.ceil2sd);
    const int FL = L.FL;
    const int LA = S.posA + S.ciA0 - m, HA = S.posA + S.ciA1 - m;
    const int LB = S.posB + S.ciB0 + m + 1, HB = S.posB + S.ciB1 + m + 1;
    set_win(w.altA_lo, w.altA_w1, LA - (o1 ? 0 : FL), HA + (o1 ? FL : 0), ok && !small_del);
    set_win(w.altB_lo, w.altB_w1, LB - (o2 ? 0 : FL), HB + (o2 ? FL : 0), ok && !small_del);
    set_win(w.recA_lo, w.recA_w1, LA - (o1 ? FL : 0), HA + (o1 ? 0 : FL), ok && svtype == SV_INV);
    set_win(w.recB_lo, w.recB_w1, LB - (o2 ? FL : 0), HB + (o2 ? 0 : FL), ok && svtype == SV_INV);
    set_win(w.rAa_lo, w.rAa_w1, S.wA0 - FL, S.wA0, ok && !small_del);
    set_win(w.rAb_lo, w.rAb_w1, S.wA1 + 1, S.wA1 + 1 + FL, ok && !small_del);
    set_win(w.rBa_lo, w.rBa_w1, S.wB0 - FL, S.wB0, ok && !small_del);
    set_win(w.rBb_lo, w.rBb_w1, S.wB1 + 1, S.wB1 + 1 + FL, ok && !small_del);
    /* second histogram key is o - Lk: Lk = var_length (DEL) or the integral mean+3sd (others);
     * "no key" becomes 0x7fffffff, which no |b_end - a_start| of a straddling pair reaches */
    const int Lk = is_del ? S.var_length : L.nondel_L;
    w.Lk = (!is_del && Lk < 0) ? 0x7fffffffu : (unsigned)Lk;
    w.hist_off = (unsigned)L.hist_off;
    w.hist_len = (unsigned)L.hist_len;
    w.flags = (ok && small_counts && !(is_del && Lk < 0)) ? 1u : 0u;
    return w;
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

__device__ __forceinline__ bool in_win(int v, unsigned lo, unsigned w1) { return ((unsigned)v - lo) < w1; }

__device__ __forceinline__ void prefetch_l2(const void *ptr)
{
#if SVGT_L2_PREFETCH && !(SVGT_DIAG & 16)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;      /* measured: no effect once two chunks are in flight in registers (profiles/README.md) */
#endif
}


/*
 * Predicate chains in PTX.  The C++ forms of these tests compile to an ISETP plus a SEL per
 * boolean (every bool is materialised as 0/1 and recombined with LOP3); the kernel is ALU-pipe
 * bound, so the chains are written with setp.<cmp>.and so one compare also ANDs in the running
 * predicate.  All compares are exact for every int32 input (no subtract-and-test-sign tricks).
 */

/* is_ref_seq for both reads against both breakends (parsers.py:801-816), both windows valid.
 * SAME: both breakends on one contig (tA == tB), so one tid compare serves both windows. */
template <bool SAME>
__device__ __forceinline__ void hits_chain(int a_start, int a_end, int b_start, int b_end, int tidA, int tidB, int fl,
                                           int tA, int tB, int wA0, int wA1, int wB0, int wB1, int &hitA, int &hitB)
{
    if (SAME) {
        asm("{\n\t"
            ".reg .pred p, q;\n\t"
            ".reg .b32 t;\n\t"
            "setp.le.s32 p, %2, %11;\n\t"
            "setp.ge.and.s32 p, %3, %12, p;\n\t"
            "setp.le.s32 q, %2, %13;\n\t"
            "setp.ge.and.s32 q, %3, %14, q;\n\t"
            "or.pred p, p, q;\n\t"
            "setp.eq.and.s32 p, %6, %9, p;\n\t"
            "and.b32 t, %8, 1;\n\t"
            "setp.ne.and.s32 p, t, 0, p;\n\t"
            "selp.s32 %0, 1, 0, p;\n\t"
            "setp.le.s32 p, %4, %11;\n\t"
            "setp.ge.and.s32 p, %5, %12, p;\n\t"
            "setp.le.s32 q, %4, %13;\n\t"
            "setp.ge.and.s32 q, %5, %14, q;\n\t"
            "or.pred p, p, q;\n\t"
            "setp.eq.and.s32 p, %7, %9, p;\n\t"
            "and.b32 t, %8, 2;\n\t"
            "setp.ne.and.s32 p, t, 0, p;\n\t"
            "selp.s32 %1, 1, 0, p;\n\t"
            "}"
            : "=r"(hitA), "=r"(hitB)
            : "r"(a_start), "r"(a_end), "r"(b_start), "r"(b_end), "r"(tidA), "r"(tidB), "r"(fl), "r"(tA), "r"(tB),
              "r"(wA0), "r"(wA1), "r"(wB0), "r"(wB1));
    } else {
        asm("{\n\t"
            ".reg .pred p, q;\n\t"
            ".reg .b32 t;\n\t"
            "setp.le.s32 p, %2, %11;\n\t"
            "setp.ge.and.s32 p, %3, %12, p;\n\t"
            "setp.eq.and.s32 p, %6, %9, p;\n\t"
            "setp.le.s32 q, %2, %13;\n\t"
            "setp.ge.and.s32 q, %3, %14, q;\n\t"
            "setp.eq.and.s32 q, %6, %10, q;\n\t"
            "or.pred p, p, q;\n\t"
            "and.b32 t, %8, 1;\n\t"
            "setp.ne.and.s32 p, t, 0, p;\n\t"
            "selp.s32 %0, 1, 0, p;\n\t"
            "setp.le.s32 p, %4, %11;\n\t"
            "setp.ge.and.s32 p, %5, %12, p;\n\t"
            "setp.eq.and.s32 p, %7, %9, p;\n\t"
            "setp.le.s32 q, %4, %13;\n\t"
            "setp.ge.and.s32 q, %5, %14, q;\n\t"
            "setp.eq.and.s32 q, %7, %10, q;\n\t"
            "or.pred p, p, q;\n\t"
            "and.b32 t, %8, 2;\n\t"
            "setp.ne.and.s32 p, t, 0, p;\n\t"
            "selp.s32 %1, 1, 0, p;\n\t"
            "}"
            : "=r"(hitA), "=r"(hitB)
            : "r"(a_start), "r"(a_end), "r"(b_start), "r"(b_end), "r"(tidA), "r"(tidB), "r"(fl), "r"(tA), "r"(tB),
              "r"(wA0), "r"(wA1), "r"(wB0), "r"(wB1));
    }
}

/*
 * is_pair_straddle x3 (alt, ref at A, ref at B; parsers.py:821-857 through the per-(site, library)
 * windows), p_concordant as 19*h1 > h2 on the histogram counts (parsers.py:861-882)
 * and the selection of the prob_mapq LUT indices that realise
 *     p_alt = alt ? (DEL & p_conc ? 0 : pmA * pmB) : 0            singlesample.py:305-318
 *     p_ref = (refA | refB) & (!(refA & refB) | DEL) & p_conc ? pmA * pmB * (refA + refB) / 2 : 0   :336-350
 * Outputs LUT indices (0 selects pm[0] == 0.0; +256 selects the halved table) and a `tie` flag
 * (19*h1 == h2 != 0: the caller evaluates the literal fp64 expression).
 * SAME: tA == tB, so "both reads on the site's contig" is one predicate shared by all three tests.
 */
#define SVGT_PE_DECL                                                                                  \
    "{\n\t"                                                                                           \
    ".reg .pred pf, pa, pfr, ra, rb, p1, p2, pc, pt, pboth, pany, pdel, pron, paon;\n\t"             \
    ".reg .b32 d, o, k2, h1, h2, l19, t;\n\t"                                                         \
    ".reg .b64 ad;\n\t"                                                                               \
    "setp.ne.s32 pf, %10, 0;\n\t"
#define SVGT_PE_TESTS_ANY                                                                             \
    "setp.eq.and.s32 pa, %7, %11, pf;\n\t"                                                            \
    "setp.eq.and.s32 pa, %8, %12, pa;\n\t"                                                            \
    "setp.eq.and.s32 pa, %9, %13, pa;\n\t"                                                            \
    "sub.s32 d, %5, %15;\n\t"                                                                         \
    "setp.lt.and.u32 pa, d, %16, pa;\n\t"                                                             \
    "sub.s32 d, %6, %17;\n\t"                                                                         \
    "setp.lt.and.u32 pa, d, %18, pa;\n\t"                                                             \
    "setp.eq.and.s32 pfr, %9, 2, pf;\n\t"                                                             \
    "setp.eq.and.s32 ra, %7, %11, pfr;\n\t"                                                           \
    "setp.eq.and.s32 ra, %8, %11, ra;\n\t"                                                            \
    "sub.s32 d, %5, %19;\n\t"                                                                         \
    "setp.lt.and.u32 ra, d, %20, ra;\n\t"                                                             \
    "sub.s32 d, %6, %21;\n\t"                                                                         \
    "setp.lt.and.u32 ra, d, %22, ra;\n\t"                                                             \
    "setp.eq.and.s32 rb, %7, %12, pfr;\n\t"                                                           \
    "setp.eq.and.s32 rb, %8, %12, rb;\n\t"                                                            \
    "sub.s32 d, %5, %23;\n\t"                                                                         \
    "setp.lt.and.u32 rb, d, %24, rb;\n\t"                                                             \
    "sub.s32 d, %6, %25;\n\t"                                                                         \
    "setp.lt.and.u32 rb, d, %26, rb;\n\t"
#define SVGT_PE_TESTS_SAME                                                                            \
    "setp.eq.and.s32 pf, %7, %11, pf;\n\t"                                                            \
    "setp.eq.and.s32 pf, %8, %11, pf;\n\t"                                                            \
    "setp.eq.and.s32 pa, %9, %13, pf;\n\t"                                                            \
    "sub.s32 d, %5, %15;\n\t"                                                                         \
    "setp.lt.and.u32 pa, d, %16, pa;\n\t"                                                             \
    "sub.s32 d, %6, %17;\n\t"                                                                         \
    "setp.lt.and.u32 pa, d, %18, pa;\n\t"                                                             \
    "setp.eq.and.s32 pfr, %9, 2, pf;\n\t"                                                             \
    "sub.s32 d, %5, %19;\n\t"                                                                         \
    "setp.lt.and.u32 ra, d, %20, pfr;\n\t"                                                            \
    "sub.s32 d, %6, %21;\n\t"                                                                         \
    "setp.lt.and.u32 ra, d, %22, ra;\n\t"                                                             \
    "sub.s32 d, %5, %23;\n\t"                                                                         \
    "setp.lt.and.u32 rb, d, %24, pfr;\n\t"                                                            \
    "sub.s32 d, %6, %25;\n\t"                                                                         \
    "setp.lt.and.u32 rb, d, %26, rb;\n\t"
/* with SAME, pf has been narrowed to "fast and both reads on the contig": a pair elsewhere can
 * not straddle anything, so skipping its histogram look-ups changes nothing */
#define SVGT_PE_PCONC                                                                                 \
    "sad.s32 o, %6, %5, 0;\n\t"                                                                       \
    "sub.s32 k2, o, %27;\n\t"                                                                         \
    "mov.b32 h1, 0;\n\t"                                                                              \
    "mov.b32 h2, 0;\n\t"                                                                              \
    "setp.lt.and.u32 p1, o, %29, pf;\n\t"                                                             \
    "setp.lt.and.u32 p2, k2, %29, pf;\n\t"                                                            \
    "add.s32 t, o, %28;\n\t"                                                                          \
    "mad.wide.u32 ad, t, 4, %30;\n\t"                                                                 \
    "@p1 ld.u32 h1, [ad];\n\t"                                                                        \
    "add.s32 t, k2, %28;\n\t"                                                                         \
    "mad.wide.u32 ad, t, 4, %30;\n\t"                                                                 \
    "@p2 ld.u32 h2, [ad];\n\t"                                                                        \
    "mul.lo.u32 l19, h1, 19;\n\t"                                                                     \
    "setp.gt.u32 pc, l19, h2;\n\t"                                                                    \
    "setp.eq.u32 pt, l19, h2;\n\t"                                                                    \
    "setp.ne.and.u32 pt, h2, 0, pt;\n\t"                                                              \
    "selp.s32 %3, 1, 0, pt;\n\t"
#define SVGT_PE_WEIGHTS                                                                               \
    "setp.ne.s32 pdel, %14, 0;\n\t"                                                                   \
    "and.pred pboth, ra, rb;\n\t"                                                                     \
    "or.pred pany, ra, rb;\n\t"                                                                       \
    "and.pred p1, pboth, !pdel;\n\t"                                                                  \
    "and.pred pron, pany, !p1;\n\t"                                                                   \
    "and.pred pron, pron, pc;\n\t"                                                                    \
    "and.pred p2, pdel, pc;\n\t"                                                                      \
    "and.pred paon, pa, !p2;\n\t"                                                                     \
    "selp.s32 %0, %31, 0, paon;\n\t"                                                                  \
    "selp.s32 %1, %31, 0, pron;\n\t"                                                                  \
    "add.s32 t, %32, 256;\n\t"                                                                        \
    "selp.s32 %2, %32, t, pboth;\n\t"                                                                 \
    "selp.s32 %4, 1, 0, pa;\n\t"                                                                      \
    "}"
#define SVGT_PE_OPERANDS                                                                              \
    : "=r"(idx_alt), "=r"(idx_ref), "=r"(idx_refB), "=r"(tie), "=r"(alt_out)                           \
    : "r"(a_start), "r"(b_end), "r"(tidA), "r"(tidB), "r"(st), "r"(fastflag), "r"(tA), "r"(tB), "r"(o12),  \
      "r"(is_del), "r"(w0.x), "r"(w0.y), "r"(w0.z), "r"(w0.w), "r"(w2.x), "r"(w2.y), "r"(w2.z), "r"(w2.w), \
      "r"(w3.x), "r"(w3.y), "r"(w3.z), "r"(w3.w), "r"(Lk), "r"(hist_off), "r"(hist_len), "l"(hist), "r"(mqA), \
      "r"(mqB)

template <bool SAME>
__device__ __forceinline__ void pe_chain(int a_start, int b_end, int tidA, int tidB, int st, int fastflag,
                                         int tA, int tB, int o12, int is_del, uint4 w0, uint4 w2, uint4 w3,
                                         unsigned Lk, unsigned hist_off, unsigned hist_len, const unsigned *hist,
                                         int mqA, int mqB, int &idx_alt, int &idx_ref, int &idx_refB, int &tie,
                                         int &alt_out)
{
    if (SAME)
        asm(SVGT_PE_DECL SVGT_PE_TESTS_SAME SVGT_PE_PCONC SVGT_PE_WEIGHTS SVGT_PE_OPERANDS);
    else
        asm(SVGT_PE_DECL SVGT_PE_TESTS_ANY SVGT_PE_PCONC SVGT_PE_WEIGHTS SVGT_PE_OPERANDS);
}

/* one 32-row fragment chunk of site S scored by the warp (phase A): what every lane parks for its row */
struct FragOut {
    double s, p_ref, p_alt;         /* ref_seq / ref_span / alt_span addends this row parks (see below)     */
    int ia, ib;                     /* prob_mapq LUT indices of the row's own ref_seq addends a, b (0 = none) */
    int lead;                       /* warp-uniform: leading rows that continue the previous chunk's fragment */
    bool need_idx;                  /* warp-uniform: phase B may read ia / ib of this chunk (lean kernel only)  */
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
template <int ASSOC>
__device__ __forceinline__ FragOut score_frag_chunk(const SvgtParams &p, const Tables &t, const SiteS &S, const Win *wins,
                                                    const double *s_pm, const LibK *s_lib, const unsigned *hist,
                                                    const int lane, const int n, const int g, const int m,
                                                    const int4 lo, const int4 hi, unsigned &carryA, unsigned &carryB,
                                                    int &err)
{
    const unsigned full = 0xffffffffu;
        const int4 s0 = *reinterpret_cast<const int4 *>(&S.tA);   /* tA tB wA0 wA1 */
        const int4 s1 = *reinterpret_cast<const int4 *>(&S.wB0);  /* wB0 wB1 meta var_length */
        const bool rv = lane < n;
        const unsigned vm = n >= 32 ? full : ((1u << n) - 1u);
        const int fl = rv ? hi.w : 0;
        const int smeta = s1.z;
        const int svtype = smeta & 3;
        const bool is_del = svtype == SV_DEL;
        /* the PTX chains cover sites whose two ref-seq windows are valid and that are not INV
         * (reciprocal orientation); everything else takes the same tests in C++ (site-uniform) */
        const bool common = ((smeta >> 8) & 3) == 3 && svtype != SV_INV;
        const int mqA = hi.z & 0xFF, mqB = (hi.z >> 8) & 0xFF;
        const unsigned lib = ((unsigned)hi.z) >> 16;
        const bool isx = (fl & F_EXTRA) != 0;
        const bool paired = ((fl & F_PAIRED) != 0) & !isx;
        const Win *wp = &wins[lib < (unsigned)kWLibs ? lib : 0u];
        const uint4 w4 = *reinterpret_cast<const uint4 *>(&wp->Lk);          /* Lk hist_off hist_len flags */
        const bool fast = paired & (lib < (unsigned)kWLibs) & ((w4.w & 1u) != 0u);
        const uint4 w0 = *reinterpret_cast<const uint4 *>(&wp->altA_lo);
        const uint4 w2 = *reinterpret_cast<const uint4 *>(&wp->rAa_lo);
        const uint4 w3 = *reinterpret_cast<const uint4 *>(&wp->rBa_lo);
        const int st = (fl >> 2) & 3, o12 = (smeta >> 2) & 3;

        /* ---- is_ref_seq hits (parsers.py:801-816) ---- */
        int hitA, hitB;
        const bool same = SVGT_USE_SAME && s0.x == s0.y;    /* both breakends on one contig */
#if SVGT_DIAG & 1
        hitA = lo.x & 1; hitB = lo.w & 1;
#else
        if (common && same) {
            hits_chain<true>(lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, fl, s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, hitA, hitB);
        } else if (common) {
            hits_chain<false>(lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, fl, s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, hitA, hitB);
        } else {
            const bool okA = (smeta >> 8) & 1, okB = (smeta >> 9) & 1;
            const bool ea = hi.x == s0.x, eb = hi.x == s0.y, fa = hi.y == s0.x, fb = hi.y == s0.y;
            hitA = ((fl & F_HAS_A) != 0) && ((ea && okA && lo.x <= s0.z && lo.y >= s0.w) ||
                                             (eb && okB && lo.x <= s1.x && lo.y >= s1.y));
            hitB = ((fl & F_HAS_B) != 0) && ((fa && okA && lo.z <= s0.z && lo.w >= s0.w) ||
                                             (fb && okB && lo.z <= s1.x && lo.w >= s1.y));
        }
#endif
        /* EXTRA interval rows feed the next main row's MULTI slots (evidence.py) */
        const unsigned XM = __ballot_sync(full, (fl & (F_EXTRA | F_MULTI_A | F_MULTI_B | F_CONT)) != 0);
        unsigned nm = vm;
        if (XM != 0u || ((carryA | carryB) >> g) & 1u) {
#ifdef SVGT_MARK
            asm volatile("membar.cta;" ::: "memory");
#endif
            const bool cA = (carryA >> g) & 1u, cB = (carryB >> g) & 1u;
            const unsigned E = __ballot_sync(full, isx);
            const unsigned HA = __ballot_sync(full, isx && hitA), HB = __ballot_sync(full, isx && hitB);
            const unsigned below = (1u << lane) - 1u;
            const unsigned z = ~E & below;
            unsigned runm;
            bool reach0;
            if (z == 0u) { runm = below; reach0 = true; }
            else { const int pz = 31 - __clz(z); runm = below & ~((2u << pz) - 1u); reach0 = false; }
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
            carryA = (carryA & ~(1u << g)) | ((unsigned)nA << g);
            carryB = (carryB & ~(1u << g)) | ((unsigned)nB << g);
            nm = __ballot_sync(full, rv && !(fl & (F_CONT | F_EXTRA)));
            if (isx) { hitA = 0; hitB = 0; }
#ifdef SVGT_MARK
            asm volatile("membar.cta;" ::: "memory");
#endif
        }

        /* ---- paired-end evidence -> LUT indices ---- */
        /* weights as prob_mapq LUT indices: entry 0 is exactly 0.0, entries 256.. are halved */
        auto weights = [&](bool alt, bool refA, bool refB, bool pc, int &ia, int &ir, int &irB) {
            const bool both = refA & refB;
            const bool ref_on = (refA | refB) & (!both | is_del) & pc;
            const bool alt_on = alt & !(is_del & pc);
            ia = alt_on ? mqA : 0; ir = ref_on ? mqA : 0; irB = mqB + (both ? 0 : 256);
        };
        int idx_alt, idx_ref, idx_refB, tie = 0;
#if SVGT_DIAG & 1
        idx_alt = mqA & st; idx_ref = mqB & o12; idx_refB = mqB;
        if (false) {
#else
        if (common && same) {
#endif
            int alt_i;
            pe_chain<true>(lo.x, lo.w, hi.x, hi.y, st, (int)fast, s0.x, s0.y, o12, (int)is_del, w0, w2, w3, w4.x,
                           w4.y, w4.z, hist, mqA, mqB, idx_alt, idx_ref, idx_refB, tie, alt_i);
        } else if (common) {
            int alt_i;
            pe_chain<false>(lo.x, lo.w, hi.x, hi.y, st, (int)fast, s0.x, s0.y, o12, (int)is_del, w0, w2, w3, w4.x,
                            w4.y, w4.z, hist, mqA, mqB, idx_alt, idx_ref, idx_refB, tie, alt_i);
        } else {
            const bool ea = hi.x == s0.x, eb = hi.x == s0.y, fa = hi.y == s0.x, fb = hi.y == s0.y;
            const bool ab = ea & fb & fast;
            bool alt = ab & (st == o12) & in_win(lo.x, w0.x, w0.y) & in_win(lo.w, w0.z, w0.w);
            if (svtype == SV_INV) {
                const uint4 w1 = *reinterpret_cast<const uint4 *>(&wp->recA_lo);
                alt |= ab & (st == (o12 ^ 3)) & in_win(lo.x, w1.x, w1.y) & in_win(lo.w, w1.z, w1.w);
            }
            const bool fr = (st == 2) & fast;
            const bool refA = fr & ea & fa & in_win(lo.x, w2.x, w2.y) & in_win(lo.w, w2.z, w2.w);
            const bool refB = fr & eb & fb & in_win(lo.x, w3.x, w3.y) & in_win(lo.w, w3.z, w3.w);
            const unsigned o = __sad(lo.w, lo.x, 0u);
            const unsigned k2 = o - w4.x;
            const unsigned h1 = (fast & (o < w4.z)) ? hist[w4.y + o] : 0u;
            const unsigned h2 = (fast & (k2 < w4.z)) ? hist[w4.y + k2] : 0u;
            const unsigned l19 = 19u * h1;
            tie = (l19 == h2) & (h2 != 0u);
            weights(alt, refA, refB, l19 > h2, idx_alt, idx_ref, idx_refB);
        }
        const bool slow = paired & !fast;
        if (__any_sync(full, slow | (tie != 0))) {
            if (slow | (tie != 0)) {
                bool alt, refA, refB, pc;
                slow_row(p, t, S, lo, hi, s_lib, m, err, alt, refA, refB, pc);
                weights(alt, refA, refB, pc, idx_alt, idx_ref, idx_refB);
            }
        }
        /* singlesample.py:254-259: a = pm[A] if readA covers a breakend; :305-350: p_alt, p_ref.
         * 0.0 * x = 0.0 and the exact halving keep these bit-identical to the reference forms */
        const double va = s_pm[hitA ? mqA : 0];
        const double vb = s_pm[hitB ? mqB : 0];
        const double pmB = s_pm[mqB];
        const double p_alt = __dmul_rn(s_pm[idx_alt], pmB);
        const double p_ref = __dmul_rn(s_pm[idx_ref], s_pm[idx_refB]);
    FragOut o;
    o.s = __dadd_rn(va, vb); o.p_ref = p_ref; o.p_alt = p_alt;
    o.ia = hitA ? mqA : 0; o.ib = hitB ? mqB : 0;
    o.lead = 0; o.need_idx = true;
    if (ASSOC == SVGT_ASSOC_SSO && nm != vm) {          /* the chunk has CONT / EXTRA rows (warp-uniform) */
        const unsigned NN = vm & ~nm;                   /* rows that continue a fragment            */
        o.lead = nm ? __ffs(nm) - 1 : (n < 32 ? n : 32);
        const bool nonnew = (NN >> lane) & 1u;
        const bool inner = nonnew && lane >= o.lead;    /* continues a fragment that starts in this chunk */
        /* distance to the fragment's first row in this chunk */
        const unsigned below = nm & ((1u << lane) - 1u);
        const int dist = inner ? lane - (31 - __clz(below)) : 0;
        /* a continuation row normally carries no paired-end weight; if one does, fold those too */
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
        /* the running sub-total lives in the LAST row of the fragment (within the chunk); earlier rows
         * park 0.0, so phase B's "acc += pend; pend = s" leaves the sub-total pending, un-flushed, in
         * case the fragment continues in the next chunk */
        const bool has_next = lane + 1 < 32 && ((NN >> (lane + 1 < 32 ? lane + 1 : 31)) & 1u);
        if (lane >= o.lead && has_next) {
            o.s = 0.0;
            if (pe_too) { o.p_ref = 0.0; o.p_alt = 0.0; }    /* folded forward above; otherwise every row keeps its own */
        }
    }
    return o;
}

/* park a scored fragment row for phase B: {s, LUT indices of a and b, p_ref, p_alt} (32 B) */
__device__ __forceinline__ void park_frag(void *dst32, const FragOut &o)
{
    double2 *d = reinterpret_cast<double2 *>(dst32);
    d[0] = make_double2(o.s, __hiloint2double(o.ib, o.ia));
    d[1] = make_double2(o.p_ref, o.p_alt);
}

/* phase B of one fragment chunk for chain c (0 ref_seq, 1 ref_span, 2 alt_span) of one site */
template <int ASSOC>
__device__ __forceinline__ void replay_frag(const void *base, int c, int cnt, int lead, const double *s_pm, double &acc,
                                            double &pend)
{
    const double *px = reinterpret_cast<const double *>(base) + (c == 0 ? 0 : c + 1);
    const int2 *pi = reinterpret_cast<const int2 *>(reinterpret_cast<const char *>(base) + 8);   /* .x = ia, .y = ib */
    if (cnt <= 0) return;                               /* no chunk of this site in the super-step: `lead` is stale */
    lead = lead < cnt ? lead : cnt;
    if (ASSOC == SVGT_ASSOC_CLASSIC) {
        /* classic.py:306-311,339-408: every read goes straight into the site sum */
        if (c == 0) {
            for (int j = 0; j < cnt; ++j) {
                const int2 ix = pi[j * 4];
                acc = __dadd_rn(__dadd_rn(acc, s_pm[ix.x]), s_pm[ix.y]);
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) acc = __dadd_rn(acc, px[j * 4]);
        }
        lead = cnt;                                     /* nothing left for the sso loops below */
    }
    for (int j = 0; j < (ASSOC == SVGT_ASSOC_CLASSIC ? 0 : lead); ++j) {                    /* rows continuing the previous chunk's last fragment */
        if (c == 0) {
            const int2 ix = pi[j * 4];
            pend = __dadd_rn(__dadd_rn(pend, s_pm[ix.x]), s_pm[ix.y]);
        } else {
            pend = __dadd_rn(pend, px[j * 4]);
        }
    }
#pragma unroll 4
    for (int j = lead; j < cnt; ++j) {
        acc = __dadd_rn(acc, pend);
        pend = px[j * 4];
    }
}

/* one 32-row split chunk (parsers.py:1122-1215, singlesample.py:262-274): parks {alt_seq, alt_clip}
 * addends; rows that are not the FIRST split of their fragment are folded into the first one's
 * sub-totals exactly as above */
struct SplitOut { double vseq, vclip; int lead; };

template <int ASSOC>
__device__ __forceinline__ SplitOut score_split_chunk(const SiteS &S, const double *s_pm, const int lane, const int n,
                                                      const int slop, const int4 q0, const int4 q1)
{
    const unsigned full = 0xffffffffu;
    const bool rv = lane < n;
    const unsigned vm = n >= 32 ? full : ((1u << n) - 1u);
    /* arrange breakends left to right, parsers.py:1143-1161 */
    const int o1 = (S.meta >> 2) & 1, o2 = (S.meta >> 3) & 1, svtype = S.meta & 3;
    const bool swap = (S.tA != S.tB) || (S.posA > S.posB);
    const int tL = swap ? S.tB : S.tA, tR = swap ? S.tA : S.tB;
    const int pL = swap ? S.posB : S.posA, pR = swap ? S.posA : S.posB;
    const int rL = swap ? o2 : o1, rR = swap ? o1 : o2;
    const int sfl = rv ? ((q1.z >> 16) & 0xFFFF) : S_FIRST;
    const bool soft = sfl & S_SOFT_CLIP;
    const int cl = rL ? q0.y : q0.z, cr = rR ? q0.y : q0.z;       /* left piece vs L / R side */
    const int dl = rL ? q1.x : q1.y, dr = rR ? q1.x : q1.y;       /* right piece vs L / R side */
    const bool lL = (q0.x == tL) & ((unsigned)(cl - (pL - slop)) <= (unsigned)(2 * slop));
    const bool lR = (q0.x == tR) & ((unsigned)(cr - (pR - slop)) <= (unsigned)(2 * slop));
    const bool rLs = (q0.w == tL) & ((unsigned)(dl - (pL - slop)) <= (unsigned)(2 * slop));
    const bool rRs = (q0.w == tR) & ((unsigned)(dr - (pR - slop)) <= (unsigned)(2 * slop));
    const bool plain = !soft | (svtype == SV_DEL);
    const bool dup = soft & (svtype == SV_DUP), inv = soft & (svtype == SV_INV);
    const bool Ls = rv & ((plain & lL) | (dup & lR) | (inv & (lL | lR)));
    const bool Rs = rv & ((plain & rRs) | (dup & rLs) | (inv & (rLs | rRs)));
    const double x = s_pm[Ls ? (q1.z & 0xFF) : 0];
    const double y = s_pm[Rs ? ((q1.z >> 8) & 0xFF) : 0];
    const double p_alt = __dmul_rn(__dadd_rn(x, y), 0.5);       /* (.. + ..) / 2.0, exact either way */
    SplitOut o;
    o.vseq = soft ? 0.0 : p_alt; o.vclip = soft ? p_alt : 0.0; o.lead = 0;
    const unsigned nm = __ballot_sync(full, rv && (sfl & S_FIRST));
    if (ASSOC == SVGT_ASSOC_SSO && nm != vm) {
        const unsigned NN = vm & ~nm;
        o.lead = nm ? __ffs(nm) - 1 : (n < 32 ? n : 32);
        const bool nonnew = (NN >> lane) & 1u;
        const bool inner = nonnew && lane >= o.lead;
        const unsigned below = nm & ((1u << lane) - 1u);
        const int dist = inner ? lane - (31 - __clz(below)) : 0;
        const double vs0 = o.vseq, vc0 = o.vclip;
        for (int k = 1; k < 32; ++k) {
            if (!__any_sync(full, dist >= k)) break;
            const double us = __shfl_up_sync(full, o.vseq, 1), uc = __shfl_up_sync(full, o.vclip, 1);
            if (dist == k) { o.vseq = __dadd_rn(us, vs0); o.vclip = __dadd_rn(uc, vc0); }
        }
        const bool has_next = lane + 1 < 32 && ((NN >> (lane + 1 < 32 ? lane + 1 : 31)) & 1u);
        if (lane >= o.lead && has_next) { o.vseq = 0.0; o.vclip = 0.0; }
    }
    return o;
}

/* phase B of one split chunk for chain c (0 alt_seq, 1 alt_clip) */
template <int ASSOC>
__device__ __forceinline__ void replay_split(const void *base, int c, int cnt, int lead, double &acc, double &pend)
{
    const double *px = reinterpret_cast<const double *>(base) + c;
    if (cnt <= 0) return;
    lead = lead < cnt ? lead : cnt;
    if (ASSOC == SVGT_ASSOC_CLASSIC) {
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) acc = __dadd_rn(acc, px[j * 4]);
        lead = cnt;
    }
    for (int j = 0; j < (ASSOC == SVGT_ASSOC_CLASSIC ? 0 : lead); ++j) pend = __dadd_rn(pend, px[j * 4]);
#pragma unroll 4
    for (int j = lead; j < cnt; ++j) {
        acc = __dadd_rn(acc, pend);
        pend = px[j * 4];
    }
}

/* sums parked in the site's output row between the two launches */
struct ParkedSums { double ref_seq, alt_seq, alt_clip, ref_span, alt_span; };

}  // namespace
