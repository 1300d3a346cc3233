/*
 * svgt_coop.cu -- warp-cooperative scoring kernels (variants 2 and 3, the default path).
 *
 * Same arithmetic and results as the thread-per-site kernels in svgt_kernels.cu (reference
 * citations there); different mapping.  The only order-sensitive part of the path is the fp64
 * accumulation of per-fragment weights in sorted(qname) order; everything that
 * produces those weights -- is_ref_seq, is_pair_straddle x4, p_concordant, prob_mapq products
 * (svtyper/parsers.py:801-882, singlesample.py:246-353) -- is independent per row.  So the
 * path is two launches:
 *
 * svgt_tally_kernel<G>  (tally_variant_read_fragments, singlesample.py:355-404)
 *   work unit = G consecutive sites of the launch order, pulled from an atomic cursor.
 *   phase A  one ROW per lane: a warp reads 32 consecutive 32-byte rows of one site with
 *            coalesced 128-bit loads (the next chunk's rows are already in flight in registers),
 *            scores them branch-free against warp-uniform site constants and per-(site, library)
 *            integer windows kept in shared memory, and parks 4 doubles per row
 *            {a, b, p_ref, p_alt} in shared memory.  Cross-row schema state (EXTRA/MULTI interval
 *            rows) is resolved with ballots.
 *   phase B  one CHAIN per lane: lane 4g+c replays chain c (ref_seq, ref_span, alt_span |
 *            alt_seq, alt_clip) of site g over that site's 32 parked rows IN ORDER, so the serial
 *            adds of 3G chains issue together instead of one chain per instruction.
 *   The five sums of each site are parked in that site's 80-byte output row.
 *
 * svgt_call_kernel  (zeroing rules + bayesian_genotype / bayes_gt / log_choose,
 *                    singlesample.py:382-473, statistics.py:9-37)
 *   one site per thread, identity order: reads the five sums back, writes the final row.
 *
 * A site is walked 32 rows per warp step instead of one row per thread step, so the longest
 * site no longer bounds the launch and DRAM sees 1 KB sequential bursts.
 */
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

/* ordered replay of one chain over `cnt` parked rows (phase B).
 * SSO:     per row   if NEW: acc += pend, pend = 0;   pend = (pend + x) + y
 * CLASSIC: per row   acc = (acc + x) + y
 * NEW rows dominate, so the all-NEW case is a 2-add loop. */
template <int ASSOC>
__device__ __forceinline__ void replay_chain(const double *px, const double *py, int ystride, int cnt, unsigned newm,
                                             bool all_new, double &acc, double &pend)
{
    if (ASSOC == SVGT_ASSOC_CLASSIC) {
#pragma unroll 4
        for (int j = 0; j < cnt; ++j)
            acc = __dadd_rn(__dadd_rn(acc, px[j * 4]), py[j * ystride]);
    } else if (all_new) {
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const double tsum = __dadd_rn(px[j * 4], py[j * ystride]);
            acc = __dadd_rn(acc, pend);
            pend = tsum;
        }
    } else {
        for (int j = 0; j < cnt; ++j) {
            const bool nw = (newm >> j) & 1u;
            const double u = nw ? pend : 0.0;
            const double t0 = nw ? 0.0 : pend;
            acc = __dadd_rn(acc, u);
            pend = __dadd_rn(__dadd_rn(t0, px[j * 4]), py[j * ystride]);
        }
    }
}

/* sums parked in the site's output row between the two launches */
struct ParkedSums { double ref_seq, alt_seq, alt_clip, ref_span, alt_span; };

template <int G, int ASSOC>
__global__ void __launch_bounds__(SVGT_COOP_THREADS, (G >= 8 ? 2 : 3)) svgt_tally_kernel(const SvgtParams p)
{
    typedef WarpSmem<G> WS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_pm = reinterpret_cast<double *>(smem_raw);
    LibK *s_lib = reinterpret_cast<LibK *>(s_pm + 512);     /* pm[0..255], then pm[q] / 2 */
    size_t off = 512 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    off = (off + 127) & ~(size_t)127;
    WS *s_warp = reinterpret_cast<WS *>(smem_raw + off);
    off += sizeof(WS) * kCoopWarps;
    unsigned *s_hist = reinterpret_cast<unsigned *>(smem_raw + off);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    int err = 0;
    const int nl = p.n_lib < SVGT_SMEM_LIBS ? p.n_lib : SVGT_SMEM_LIBS;
    for (int i = tid; i < 256; i += SVGT_COOP_THREADS) {
        const double v = p.pm[i];
        s_pm[i] = v;
        s_pm[256 + i] = __dmul_rn(v, 0.5);
    }
    for (int i = tid; i < nl; i += SVGT_COOP_THREADS) s_lib[i] = derive_lib(p, i, &err);
    /* the 32-bit form of 19*h1 > h2 needs every histogram count below 2^26 */
    int big = 0;
    for (long long i = tid; i < p.n_hist; i += SVGT_COOP_THREADS) {
        const unsigned v = p.hist[i];
        if (p.hist_in_smem) s_hist[i] = v;
        big |= v >= (1u << 26);
    }
    WS &ws = s_warp[warp];
    if (lane < 2) ws.zero[lane] = 0.0;
    const bool small_counts = __syncthreads_or(big) == 0;

    Tables t;
    t.pm = s_pm; t.libs = s_lib; t.hist = p.hist_in_smem ? s_hist : p.hist;
    t.conc = p.consts[C_CONC]; t.disc = p.consts[C_DISC];
    const unsigned *hist = t.hist;
    const int m = p.min_aligned, slop = p.split_slop;

    /* phase-B role of this lane: chain c of interleaved site gb */
    const int gb = lane >> 2, c = lane & 3;
    const long long n_units = (p.n_sites + G - 1) / G;

    for (;;) {
        long long unit = 0;
        if (lane == 0) unit = (long long)atomicAdd(reinterpret_cast<unsigned *>(p.status + 1), 1u);
        unit = __shfl_sync(full, unit, 0);
        if (unit >= n_units) break;

        /* ---- lanes 0..G-1 read their site row and publish the scalars ---- */
        {
            const long long idx = unit * G + lane;
            const bool valid = lane < G && idx < p.n_sites;
            long long site = 0;
            if (valid) site = p.order ? (long long)p.order[idx] : idx;
            int4 a = make_int4(0, 0, 0, 0), b = a, cc = a, d = a;
            if (valid) {
                const int4 *sp = p.sites + site * 4;
                a = ldg4(sp); b = ldg4(sp + 1); cc = ldg4(sp + 2); d = ldg4(sp + 3);
            }
            const int meta = cc.y;
            const bool ranged = site_fields_in_range(a, b, m, slop);
            const bool run = valid && !(meta & SITE_SKIP) && ranged;
            const long long foff = ((long long)(unsigned)cc.z) | ((long long)cc.w << 32);
            const long long soff = ((long long)(unsigned)d.y) | ((long long)d.z << 32);
            int nf = run ? d.x : 0, ns = run ? d.w : 0;
            if (nf < 0 || foff < 0 || foff + nf > p.n_frag) { nf = 0; err = SVGT_ERR_ARG; }
            if (ns < 0 || soff < 0 || soff + ns > p.n_split) { ns = 0; err = SVGT_ERR_ARG; }
            if (lane < G) {
                SiteS &S = ws.site[lane];
                S.tA = b.z; S.tB = b.w;
                S.wA0 = a.x - m; S.wA1 = a.x + m; S.wB0 = a.y - m; S.wB1 = a.y + m;
                S.meta = (meta & 15) | ((a.x - m >= 0) << 8) | ((a.y - m >= 0) << 9);
                S.var_length = cc.x;
                S.posA = a.x; S.posB = a.y; S.ciA0 = a.z; S.ciA1 = a.w; S.ciB0 = b.x; S.ciB1 = b.y;
                S.dAB = a.y - a.x; S.nf = nf; S.foff = foff; S.soff = soff; S.ns = ns;
                S.slot = valid ? 1 : 0;
                /* the call kernel treats a site without rows as all-zero sums: nothing to park */
            }
            __syncwarp();
            for (int i = lane; i < G * kWLibs; i += 32) {
                const int g = i / kWLibs, l = i % kWLibs;
                if (l < nl) { if (ws.site[g].nf) ws.win[g][l] = make_win(ws.site[g], s_lib[l], m, small_counts); }
                else ws.win[g][l].flags = 0u;
            }
            __syncwarp();
        }

        /* pull the first 1 KB of every site's fragment and split rows towards L2 now (lane g, one
         * 128-byte line per instruction); later steps prefetch one super-step ahead of the loads */
        if (lane < G) {
            const SiteS &S = ws.site[lane];
            const char *fp = reinterpret_cast<const char *>(p.frags + 2 * S.foff);
            const char *sp = reinterpret_cast<const char *>(p.splits + 2 * S.soff);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k * 4 < S.nf) prefetch_l2(fp + k * 128);
                if (k * 4 < S.ns) prefetch_l2(sp + k * 128);
            }
        }

        double sum_frag = 0.0;      /* lane 4g+c: chain c of the fragment rows of site g */
        double sum_split = 0.0;     /* lane 4g+c: chain c of the split rows of site g    */

        /* ================= fragment rows ================= */
        {
            double acc = 0.0, pend = 0.0;
            unsigned carryA = 0u, carryB = 0u;      /* EXTRA-run hits carried into the next step, bit g */
            bool all_new = true;
            const int my_nf = lane < G ? ws.site[lane].nf : 0;

            /* chunk iterator (warp-uniform): step-major over the sites that still have rows */
            int it_step = -1;
            unsigned it_mask = 0u;
            auto advance = [&](int &st, int &g) -> bool {
                if (it_mask == 0u) {
                    ++it_step;
                    it_mask = __ballot_sync(full, my_nf > it_step * 32);
                    if (it_mask == 0u) return false;
                }
                g = __ffs(it_mask) - 1;
                it_mask &= it_mask - 1u;
                st = it_step;
                return true;
            };
            auto load_rows = [&](int st, int g, int4 &lo, int4 &hi) {
                const int n = ws.site[g].nf - st * 32;
                lo = make_int4(0, 0, 0, 0); hi = lo;
                if (lane < n) {
                    const int4 *rp = p.frags + 2 * (ws.site[g].foff + (long long)st * 32 + lane);
#if SVGT_DIAG & 8
                    lo = make_int4(lane * st, g, n, lane + 100); hi = make_int4(0, 0, 60 | (60 << 8), 0x1b);
                    (void)rp;
#else
                    lo = ldg4(rp); hi = ldg4(rp + 1);
#endif
                }
            };
            /* score one 32-row chunk (phase A); `flush` = last chunk of its super-step (phase B follows) */
            auto process = [&](const int step, const int g, const int4 lo, const int4 hi, const bool flush) {
#ifdef SVGT_MARK
                asm volatile("membar.cta;" ::: "memory");
#endif
                const int4 s0 = *reinterpret_cast<const int4 *>(&ws.site[g].tA);   /* tA tB wA0 wA1 */
                const int4 s1 = *reinterpret_cast<const int4 *>(&ws.site[g].wB0);  /* wB0 wB1 meta var_length */
                const int n = ws.site[g].nf - step * 32;
                const bool rv = lane < n;
                if (lane + 32 < n) prefetch_l2(p.frags + 2 * (ws.site[g].foff + (long long)step * 32 + 32 + lane));
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
                const Win *wp = &ws.win[g][lib < (unsigned)kWLibs ? lib : 0u];
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
                    all_new = all_new && (nm == vm);
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
                        slow_row(p, t, ws.site[g], lo, hi, s_lib, m, err, alt, refA, refB, pc);
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
                double4 *dst = reinterpret_cast<double4 *>(&ws.contrib[g][lane][0]);
                *dst = make_double4(va, vb, p_ref, p_alt);
                if (lane == 0) ws.newmask[g] = nm;
#ifdef SVGT_MARK
                asm volatile("membar.cta;" ::: "memory");
#endif

                /* ---------------- phase B at the end of each super-step ---------------- */
                if (flush) {
                    __syncwarp();
                    if (gb < G && c < 3) {
                        int cnt = ws.site[gb].nf - step * 32;
                        cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
                        const double *row0 = &ws.contrib[gb][0][0];
                        const double *px = row0 + (c == 0 ? 0 : c + 1);
                        const double *py = (c == 0) ? row0 + 1 : ws.zero;
#if !(SVGT_DIAG & 2)
                        replay_chain<ASSOC>(px, py, c == 0 ? 4 : 0, cnt, ws.newmask[gb], all_new, acc, pend);
#else
                        acc += px[0] + py[0] + cnt;
#endif
                    }
                    __syncwarp();
                    all_new = true;
                }
            };
#if SVGT_ROWBUFS == 3
            /* rotating register queue: while one chunk is scored the next TWO are in flight; the
             * rotation is 16 register moves (FMA pipe, otherwise idle) instead of a third code copy */
            int cs0 = 0, cg0 = 0, cs1 = 0, cg1 = 0, cs2 = 0, cg2 = 0;
            int4 l0, h0, l1, h1, l2, h2;
            bool k0 = advance(cs0, cg0);
            if (k0) load_rows(cs0, cg0, l0, h0);
            bool k1 = k0 && advance(cs1, cg1);
            if (k1) load_rows(cs1, cg1, l1, h1);
            while (k0) {
                const bool k2 = k1 && advance(cs2, cg2);
                if (k2) load_rows(cs2, cg2, l2, h2);
                process(cs0, cg0, l0, h0, !k1 || cs1 != cs0);
                cs0 = cs1; cg0 = cg1; l0 = l1; h0 = h1; k0 = k1;
                cs1 = cs2; cg1 = cg2; l1 = l2; h1 = h2; k1 = k2;
            }
#else
            /* two row buffers in registers: the next chunk is always in flight while one is scored */
            int st0 = 0, g0 = 0;
            int4 r0lo, r0hi, r1lo, r1hi;
            bool more = advance(st0, g0);
            if (more) load_rows(st0, g0, r0lo, r0hi);
            while (more) {
                int st1 = 0, g1 = 0;
                const bool m1 = advance(st1, g1);
                if (m1) load_rows(st1, g1, r1lo, r1hi);
                process(st0, g0, r0lo, r0hi, !m1 || st1 != st0);
                if (!m1) break;
                more = advance(st0, g0);
                if (more) load_rows(st0, g0, r0lo, r0hi);
                process(st1, g1, r1lo, r1hi, !more || st0 != st1);
            }
#endif
            if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
            sum_frag = acc;
        }

        /* ================= split rows ================= */
        {
            double acc = 0.0, pend = 0.0;
            const int my_ns = lane < G ? ws.site[lane].ns : 0;
            int sp_step = -1;
            unsigned sp_mask = 0u;
            auto advance_split = [&](int &st, int &g) -> bool {
                if (sp_mask == 0u) {
                    ++sp_step;
                    sp_mask = __ballot_sync(full, my_ns > sp_step * 32);
                    if (sp_mask == 0u || (SVGT_DIAG & 4)) return false;
                }
                g = __ffs(sp_mask) - 1;
                sp_mask &= sp_mask - 1u;
                st = sp_step;
                return true;
            };
            auto load_split = [&](int step, int g, int4 &q0, int4 &q1) {
                q0 = make_int4(0, 0, 0, 0); q1 = q0;
                const int n0 = ws.site[g].ns - step * 32;
                if (lane < n0) {
                    const int4 *rp = p.splits + 2 * (ws.site[g].soff + (long long)step * 32 + lane);
                    q0 = ldg4(rp); q1 = ldg4(rp + 1);
                }
            };
            bool all_new = true;
            int ss0 = 0, sg0 = 0, ss1 = 0, sg1 = 0, ss2 = 0, sg2 = 0;
            int4 a0, b0, a1, b1, a2, b2;
            bool e0 = advance_split(ss0, sg0);
            if (e0) load_split(ss0, sg0, a0, b0);
            bool e1 = e0 && advance_split(ss1, sg1);
            if (e1) load_split(ss1, sg1, a1, b1);
            while (e0) {
                const bool e2 = e1 && advance_split(ss2, sg2);
                if (e2) load_split(ss2, sg2, a2, b2);          /* two chunks in flight while one is scored */
                {
                    const int step = ss0, g = sg0;
                    const int4 q0 = a0, q1 = b0;
                    const SiteS &S = ws.site[g];
                    const int n = min(32, S.ns - step * 32);
                    const bool rv = lane < n;
                    /* arrange breakends left to right, parsers.py:1143-1161 */
                    const int o1 = (S.meta >> 2) & 1, o2 = (S.meta >> 3) & 1, svtype = S.meta & 3;
                    const bool swap = (S.tA != S.tB) || (S.posA > S.posB);
                    const int tL = swap ? S.tB : S.tA, tR = swap ? S.tA : S.tB;
                    const int pL = swap ? S.posB : S.posA, pR = swap ? S.posA : S.posB;
                    const int rL = swap ? o2 : o1, rR = swap ? o1 : o2;
                    const int sfl = (q1.z >> 16) & 0xFFFF;
                    const bool soft = sfl & S_SOFT_CLIP;
                    const int cl = rL ? q0.y : q0.z, cr = rR ? q0.y : q0.z;       /* left piece vs L / R side */
                    const int dl = rL ? q1.x : q1.y, dr = rR ? q1.x : q1.y;       /* right piece vs L / R side */
                    const bool lL = (q0.x == tL) & ((unsigned)(cl - (pL - slop)) <= (unsigned)(2 * slop));
                    const bool lR = (q0.x == tR) & ((unsigned)(cr - (pR - slop)) <= (unsigned)(2 * slop));
                    const bool rLs = (q0.w == tL) & ((unsigned)(dl - (pL - slop)) <= (unsigned)(2 * slop));
                    const bool rRs = (q0.w == tR) & ((unsigned)(dr - (pR - slop)) <= (unsigned)(2 * slop));
                    const bool plain = !soft | (svtype == SV_DEL);
                    const bool dup = soft & (svtype == SV_DUP), inv = soft & (svtype == SV_INV);
                    const bool Ls = (plain & lL) | (dup & lR) | (inv & (lL | lR));
                    const bool Rs = (plain & rRs) | (dup & rLs) | (inv & (rLs | rRs));
                    const double x = Ls ? s_pm[q1.z & 0xFF] : 0.0;
                    const double y = Rs ? s_pm[(q1.z >> 8) & 0xFF] : 0.0;
                    double p_alt = __dmul_rn(__dadd_rn(x, y), 0.5);
                    if (!rv) p_alt = 0.0;
                    const unsigned nm = __ballot_sync(full, rv && (sfl & S_FIRST));
                    const unsigned vm2 = n == 32 ? full : ((1u << n) - 1u);
                    all_new = all_new && (nm == vm2);
                    double2 *dst = reinterpret_cast<double2 *>(&ws.contrib[g][lane][0]);
                    *dst = make_double2(soft ? 0.0 : p_alt, soft ? p_alt : 0.0);
                    if (lane == 0) ws.newmask[g] = nm;
                }
                if (!e1 || ss1 != ss0) {                        /* last chunk of its super-step: phase B */
                    __syncwarp();
                    if (gb < G && c < 2) {
                        int cnt = ws.site[gb].ns - ss0 * 32;
                        cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
                        const double *px = &ws.contrib[gb][0][0] + c;
                        replay_chain<ASSOC>(px, ws.zero, 0, cnt, ws.newmask[gb], all_new, acc, pend);
                    }
                    __syncwarp();
                    all_new = true;
                }
                ss0 = ss1; sg0 = sg1; a0 = a1; b0 = b1; e0 = e1;
                ss1 = ss2; sg1 = sg2; a1 = a2; b1 = b2; e1 = e2;
            }
            if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
            sum_split = acc;
        }

        /* ---- park the five sums in the site's output row (lane 4g+c holds chain c of site g) ---- */
        if (gb < G && c < 3) {
            const long long idx = unit * G + gb;
            if (idx < p.n_sites && (ws.site[gb].nf | ws.site[gb].ns)) {
                const long long site = p.order ? (long long)p.order[idx] : idx;
                double *row = reinterpret_cast<double *>(p.out + site);
                /* ParkedSums: ref_seq, alt_seq, alt_clip, ref_span, alt_span */
                if (c == 0) { row[0] = sum_frag; row[1] = sum_split; }
                else if (c == 1) { row[3] = sum_frag; row[2] = sum_split; }
                else row[4] = sum_frag;
            }
        }
        __syncwarp();
    }
    if (err) {
        atomicCAS(p.status, 0, err);
        atomicAdd(p.status + 2, 1);
    }
}

/* one site per thread: zeroing rules + genotype call on the parked sums */
__global__ void __launch_bounds__(256) svgt_call_kernel(const SvgtParams p)
{
    const long long site = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (site >= p.n_sites) return;
    const int4 *sp = p.sites + site * 4;
    const int4 a = ldg4(sp), b = ldg4(sp + 1), cc = ldg4(sp + 2), d = ldg4(sp + 3);
    const int meta = cc.y;
    int err = 0;
    svgt_out_row_t o;
    o.gl[0] = o.gl[1] = o.gl[2] = 0.0; o.sq = 0.0;
    o.gt = 0; o.gq = 0; o.dp = 0; o.ro = 0; o.ao = 0; o.qr = 0; o.qa = 0;
    o.rs = 0; o.as_ = 0; o.asc = 0; o.rp = 0; o.ap = 0;
    if (meta & SITE_SKIP) { o.gt = SVGT_GT_SKIPPED; o.gq = -1; }
    else if (!site_fields_in_range(a, b, p.min_aligned, p.split_slop)) { o.gt = SVGT_GT_BLANK; o.gq = -1; err = SVGT_ERR_RANGE; }
    else {
        ParkedSums s = {0.0, 0.0, 0.0, 0.0, 0.0};
        const long long foff = ((long long)(unsigned)cc.z) | ((long long)cc.w << 32);
        const long long soff = ((long long)(unsigned)d.y) | ((long long)d.z << 32);
        const bool fok = !(d.x < 0 || foff < 0 || foff + d.x > p.n_frag);
        const bool sok = !(d.w < 0 || soff < 0 || soff + d.w > p.n_split);
        if ((fok && d.x > 0) || (sok && d.w > 0)) {
            const double *row = reinterpret_cast<const double *>(p.out + site);
            s.ref_seq = row[0]; s.alt_seq = row[1]; s.alt_clip = row[2]; s.ref_span = row[3]; s.alt_span = row[4];
        }
        Tables t;
        t.pm = p.pm; t.libs = nullptr; t.hist = p.hist;
        t.conc = 0.0; t.disc = 0.0;
        call_site(p, t, meta & 3, s.ref_seq, s.alt_seq, s.alt_clip, s.ref_span, s.alt_span, o, err);
    }
    int4 *dst = reinterpret_cast<int4 *>(p.out + site);
    const int4 *src = reinterpret_cast<const int4 *>(&o);
#pragma unroll
    for (int i = 0; i < 5; ++i) dst[i] = src[i];
    if (err) {
        atomicCAS(p.status, 0, err);
        atomicAdd(p.status + 2, 1);
    }
}

template <int G>
size_t coop_smem_bytes(const SvgtParams &p)
{
    size_t off = 512 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    off = (off + 127) & ~(size_t)127;
    off += sizeof(WarpSmem<G>) * kCoopWarps;
    if (p.hist_in_smem) off += (size_t)p.n_hist * sizeof(unsigned);
    return off;
}

template <int G>
int launch_coop(const SvgtParams &p, cudaStream_t stream)
{
    auto kern = (p.assoc_mode == SVGT_ASSOC_CLASSIC) ? svgt_tally_kernel<G, SVGT_ASSOC_CLASSIC>
                                                     : svgt_tally_kernel<G, SVGT_ASSOC_SSO>;
    const size_t smem = coop_smem_bytes<G>(p);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SVGT_COOP_THREADS, smem)) != cudaSuccess)
        return (int)e;
    if (per_sm < 1) per_sm = 1;
    const long long units = (p.n_sites + G - 1) / G;
    const long long want = (units + kCoopWarps - 1) / kCoopWarps;
    const long long cap = (long long)sms * per_sm;      /* persistent: one resident wave */
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, SVGT_COOP_THREADS, smem, stream>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    const int cgrid = (int)((p.n_sites + 255) / 256);
    svgt_call_kernel<<<cgrid, 256, 0, stream>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace

int svgt_launch_coop(const SvgtParams &p, int variant, cudaStream_t stream)
{
    if (variant == SVGT_VAR_COOP4) return launch_coop<4>(p, stream);
    return launch_coop<8>(p, stream);
}
