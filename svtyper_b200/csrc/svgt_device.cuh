/*
 * svgt_device.cuh -- device functions shared by the scoring kernels (thread-per-site variants in
 * svgt_kernels.cu, warp-cooperative variant in svgt_coop.cu): schema constants, per-library
 * constants, the literal fp64 helpers, p_concordant(), the per-row scoring of the thread-per-site
 * path and the genotype call.  Reference citations are in svgt_kernels.cu's header.
 */
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "svgt_kernels.cuh"

namespace {

/* ---- schema constants (svtyper_b200/evidence.py) ---- */
enum { SV_DEL = 0, SV_DUP = 1, SV_INV = 2, SV_BND = 3 };
enum { SITE_O1_REV = 1 << 2, SITE_O2_REV = 1 << 3, SITE_SKIP = 1 << 4 };
enum {
    F_HAS_A = 1 << 0, F_HAS_B = 1 << 1, F_REV_A = 1 << 2, F_REV_B = 1 << 3, F_PAIRED = 1 << 4,
    F_CONT = 1 << 5, F_EXTRA = 1 << 6, F_MULTI_A = 1 << 7, F_MULTI_B = 1 << 8
};
enum { S_SOFT_CLIP = 1 << 0, S_FIRST = 1 << 1 };
enum { C_NONDUP_ALT = 0, C_NONDUP_REF = 3, C_DUP_ALT = 6, C_DUP_REF = 9, C_CONC = 12, C_DISC = 13,
       C_POW10_MIN_X = 14 };

constexpr int kRange = 1 << 30;       /* site coordinates must lie in (-2^30, 2^30)   */
constexpr int kCiRange = 1 << 28;
constexpr int kStages = 3;            /* bulk ring depth                               */
constexpr int kSlotBytes = SVGT_STAGE_ROWS * 32 + 16;   /* padded: conflict-free LDS.128 */

/* per-library constants derived once per CTA */
struct LibK {
    double flank, two_sd, N;
    int FL;          /* floor(flank) when safe                                  */
    int ceil2sd;     /* d < two_sd  <=>  d < ceil2sd   for int d                */
    int hist_off, hist_len, nondel_L;
    int safe;        /* integer rewrites are exact for this library             */
};

struct Tables {
    const double *pm;        /* shared memory copy of the prob_mapq LUT          */
    const LibK *libs;        /* shared memory, first min(n_lib, SVGT_SMEM_LIBS)  */
    const unsigned *hist;    /* shared or global                                 */
    double conc, disc;
};

struct SiteK {
    int tA, tB;
    int LA, HA, LB, HB;      /* alt windows on a_start / b_end before the flank  */
    int wA0, wA1, wB0, wB1;  /* pos -/+ min_aligned                              */
    int dAB;                 /* posB - posA (post-increment)                     */
    int var_length;
    int meta;                /* svtype | O1_REV | O2_REV | okA<<8 | okB<<9       */
    /* literal path needs the raw fields */
    int posA, posB, ciA0, ciA1, ciB0, ciB1;
};

struct Acc {
    double ref_seq, sub_ref, ref_span, alt_span;
    int pend;
    int err;
};

#ifndef SVGT_LDG_MODE
#define SVGT_LDG_MODE 2
#endif
/* 128-bit row load: 0 = read-only (nc) path, 1 = L2 only (.cg), 2 = streaming (.cs), 3 = plain */
__device__ __forceinline__ int4 ldg4(const int4 *p)
{
#if SVGT_LDG_MODE == 1
    return __ldcg(p);
#elif SVGT_LDG_MODE == 2
    return __ldcs(p);
#elif SVGT_LDG_MODE == 3
    return *p;
#else
    return __ldg(p);
#endif
}

__device__ __forceinline__ LibK derive_lib(const SvgtParams &p, int l, int *err)
{
    LibK k;
    k.flank = p.lib_f64[4 * l + 0];
    k.two_sd = p.lib_f64[4 * l + 1];
    k.N = p.lib_f64[4 * l + 2];
    const int4 li = p.lib_i32[l];
    k.hist_off = li.x; k.hist_len = li.y; k.nondel_L = li.z;
    if (li.x < 0 || li.y < 0 || (long long)li.x + li.y > p.n_hist) { k.hist_len = 0; *err = SVGT_ERR_ARG; }
    const double fl = floor(k.flank);
    const double frac = k.flank - fl;      /* exact */
    const double eps = 9.5367431640625e-07; /* 2^-20 */
    k.safe = (fabs(fl) <= (double)kCiRange) && (frac == 0.0 || (frac >= eps && frac <= 1.0 - eps)) &&
             (k.N > 0.0) && (k.N < CUDART_INF);
    k.FL = k.safe ? (int)fl : 0;
    const double c2 = ceil(k.two_sd);
    if (!(k.two_sd == k.two_sd)) k.ceil2sd = INT_MIN;          /* NaN: comparison is false */
    else if (c2 >= 2147483647.0) k.ceil2sd = INT_MAX;
    else if (c2 <= -2147483648.0) k.ceil2sd = INT_MIN;
    else k.ceil2sd = (int)c2;
    return k;
}

/* ---------------- literal (oracle-shaped) helpers: used for unsafe libraries ---------------- */

__device__ __noinline__ bool straddle_literal(int4 lo, int4 hi, int tA, long long posA, long long ciA0,
                                              long long ciA1, int tB, long long posB, long long ciB0,
                                              long long ciB1, int o1rev, int o2rev, int m, double flank)
{
    /* parsers.py:821-857 */
    const int fl = hi.w;
    if (!(fl & F_PAIRED)) return false;
    if ((!!(fl & F_REV_A)) != o1rev) return false;
    if ((!!(fl & F_REV_B)) != o2rev) return false;
    if (hi.x != tA || hi.y != tB) return false;
    const long long i0 = (long long)lo.x + m;
    const long long i1 = (long long)lo.w - m - 1;
    if (!o1rev && (i0 > posA + ciA1 || (double)i0 < __dsub_rn((double)(posA + ciA0), flank))) return false;
    if (o1rev && (i0 < posA + ciA0 || (double)i0 > __dadd_rn((double)(posA + ciA1), flank))) return false;
    if (!o2rev && (i1 > posB + ciB1 || (double)i1 < __dsub_rn((double)(posB + ciB0), flank))) return false;
    if (o2rev && (i1 < posB + ciB0 || (double)i1 > __dadd_rn((double)(posB + ciB1), flank))) return false;
    return true;
}

__device__ __forceinline__ unsigned hist_at(const Tables &t, const LibK &L, long long key)
{
    if (key < 0 || key >= L.hist_len) return 0u;   /* Counter: missing key -> 0 */
    return t.hist[L.hist_off + (int)key];
}

/* parsers.py:861-882 evaluated literally in fp64 */
__device__ __noinline__ bool p_conc_literal(const Tables &t, const LibK &L, unsigned h1, unsigned h2)
{
    const double d1 = __ddiv_rn((double)h1, L.N);
    const double d2 = __ddiv_rn((double)h2, L.N);
    const double num = __dmul_rn(d1, t.conc);
    const double den = __dadd_rn(__dmul_rn(t.conc, d1), __dmul_rn(t.disc, d2));
    if (den == 0.0) return false;                  /* ZeroDivisionError -> None > 0.5 -> False */
    return __ddiv_rn(num, den) > 0.5;
}

__device__ __forceinline__ bool p_concordant(const Tables &t, const LibK &L, int a_start, int b_end,
                                             bool is_del, int var_length)
{
    const long long o64 = (long long)b_end - (long long)a_start;
    const long long o = o64 < 0 ? -o64 : o64;
    const unsigned h1 = hist_at(t, L, o);
    unsigned h2 = 0u;
    if (is_del) h2 = hist_at(t, L, o - (long long)var_length);
    else if (L.nondel_L >= 0) h2 = hist_at(t, L, o - (long long)L.nondel_L);
    if (L.safe) {
        const unsigned long long l19 = 19ull * h1;
        if (l19 != (unsigned long long)h2) return l19 > (unsigned long long)h2;
        if (h1 == 0u) return false;
    }
    return p_conc_literal(t, L, h1, h2);
}

__device__ __forceinline__ bool in_range(int v, int lo, int hi) { return v >= lo && v <= hi; }

/* ---------------- one fragment row (32 B) ---------------- */
template <int ASSOC>
__device__ __forceinline__ void frag_row(const SvgtParams &p, const Tables &t, const SiteK &s, const int4 lo,
                                         const int4 hi, Acc &acc)
{
    const int fl = hi.w;
    const bool okA = (s.meta >> 8) & 1, okB = (s.meta >> 9) & 1;
    /* parsers.py:801-816 on the gap-free interval [start, end) of each slot */
    bool hitA = false, hitB = false;
    if (fl & F_HAS_A)
        hitA = (hi.x == s.tA && okA && lo.x <= s.wA0 && lo.y >= s.wA1) ||
               (hi.x == s.tB && okB && lo.x <= s.wB0 && lo.y >= s.wB1);
    if (fl & F_HAS_B)
        hitB = (hi.y == s.tA && okA && lo.z <= s.wA0 && lo.w >= s.wA1) ||
               (hi.y == s.tB && okB && lo.z <= s.wB0 && lo.w >= s.wB1);
    if (fl & F_EXTRA) { acc.pend |= (int)hitA | ((int)hitB << 1); return; }
    if (fl & F_MULTI_A) hitA = acc.pend & 1;
    if (fl & F_MULTI_B) hitB = (acc.pend >> 1) & 1;
    acc.pend = 0;

    const double pmA = t.pm[hi.z & 0xFF], pmB = t.pm[(hi.z >> 8) & 0xFF];
    const int lib = (hi.z >> 16) & 0xFFFF;
    const double a = ((fl & F_HAS_A) && hitA) ? pmA : 0.0;
    const double b = ((fl & F_HAS_B) && hitB) ? pmB : 0.0;
    if (ASSOC == SVGT_ASSOC_SSO) {
        /* singlesample.py:254-259,367: per-fragment sub-total, then into the site sum */
        if (!(fl & F_CONT)) { acc.ref_seq = __dadd_rn(acc.ref_seq, acc.sub_ref); acc.sub_ref = 0.0; }
        acc.sub_ref = __dadd_rn(__dadd_rn(acc.sub_ref, a), b);
    } else {
        /* classic.py:306-311: every read straight into the site sum */
        acc.ref_seq = __dadd_rn(__dadd_rn(acc.ref_seq, a), b);
    }

    if (!(fl & F_PAIRED)) return;
    if (lib >= p.n_lib) { acc.err = SVGT_ERR_LIB_INDEX; return; }
    LibK Ls;
    if (lib >= SVGT_SMEM_LIBS) { int e = 0; Ls = derive_lib(p, lib, &e); }
    const LibK &L = (lib < SVGT_SMEM_LIBS) ? t.libs[lib] : Ls;

    const int svtype = s.meta & 3;
    const bool is_del = svtype == SV_DEL;
    const int o1 = (s.meta >> 2) & 1, o2 = (s.meta >> 3) & 1;
    const int rA = (fl >> 2) & 1, rB = (fl >> 3) & 1;
    bool alt, recip = false, refA, refB;
    if (L.safe) {
        /* singlesample.py:289,328: small deletions carry no paired-end evidence */
        const bool small_del = is_del && (s.dAB < L.ceil2sd);
        const int FL = L.FL;
        const bool tids = (hi.x == s.tA) && (hi.y == s.tB);
        alt = !small_del && tids && rA == o1 && rB == o2 &&
              in_range(lo.x, s.LA - (o1 ? 0 : FL), s.HA + (o1 ? FL : 0)) &&
              in_range(lo.w, s.LB - (o2 ? 0 : FL), s.HB + (o2 ? FL : 0));
        if (svtype == SV_INV)
            recip = tids && rA != o1 && rB != o2 &&
                    in_range(lo.x, s.LA - (o1 ? FL : 0), s.HA + (o1 ? 0 : FL)) &&
                    in_range(lo.w, s.LB - (o2 ? FL : 0), s.HB + (o2 ? 0 : FL));
        const bool fr = !small_del && !rA && rB;
        refA = fr && hi.x == s.tA && hi.y == s.tA && in_range(lo.x, s.wA0 - FL, s.wA0) &&
               in_range(lo.w, s.wA1 + 1, s.wA1 + 1 + FL);
        refB = fr && hi.x == s.tB && hi.y == s.tB && in_range(lo.x, s.wB0 - FL, s.wB0) &&
               in_range(lo.w, s.wB1 + 1, s.wB1 + 1 + FL);
    } else {
        const int m = p.min_aligned;
        const bool small_del = is_del && ((double)((long long)s.posB - s.posA) < L.two_sd);
        alt = !small_del && straddle_literal(lo, hi, s.tA, s.posA, s.ciA0, s.ciA1, s.tB, s.posB, s.ciB0,
                                             s.ciB1, o1, o2, m, L.flank);
        if (svtype == SV_INV)
            recip = straddle_literal(lo, hi, s.tA, s.posA, s.ciA0, s.ciA1, s.tB, s.posB, s.ciB0, s.ciB1,
                                     !o1, !o2, m, L.flank);
        refA = !small_del && straddle_literal(lo, hi, s.tA, s.posA, 0, 0, s.tA, s.posA, 0, 0, 0, 1, m, L.flank);
        refB = !small_del && straddle_literal(lo, hi, s.tB, s.posB, 0, 0, s.tB, s.posB, 0, 0, 0, 1, m, L.flank);
    }
    const bool is_alt = alt || recip;
    const bool use_ref = (refA || refB) && (!(refA && refB) || is_del);
    if (!(is_alt || use_ref)) return;

    bool pc = false;
    if ((is_alt && is_del) || use_ref) pc = p_concordant(t, L, lo.x, lo.w, is_del, s.var_length);
    const double prod = __dmul_rn(pmA, pmB);
    /* singlesample.py:305-318: DEL alt weight is (1 - p_conc) with p_conc a boolean */
    const double p_alt = is_alt ? ((is_del && pc) ? 0.0 : prod) : 0.0;
    /* singlesample.py:336-350: (refA + refB) * p_ref / 2  (k = 2 -> p_ref, k = 1 -> p_ref / 2, both exact) */
    double p_ref = (use_ref && pc) ? prod : 0.0;
    if (!(refA && refB)) p_ref = __dmul_rn(p_ref, 0.5);
    acc.alt_span = __dadd_rn(acc.alt_span, p_alt);
    acc.ref_span = __dadd_rn(acc.ref_span, p_ref);
}

/* ---------------- one split row (32 B) ---------------- */
struct SplitK {
    int tL, tR, loL, hiL, loR, hiR, rL, rR, svtype;
};
struct SAcc { double alt_seq, alt_clip, sub_seq, sub_clip; };

__device__ __forceinline__ bool split_support(int tid, int start, int end, int site_tid, int lo, int hi, int rev)
{
    /* parsers.py:1122-1134 */
    const int coord = rev ? start : end;
    return tid == site_tid && coord >= lo && coord <= hi;
}

template <int ASSOC>
__device__ __forceinline__ void split_row(const Tables &t, const SplitK &k, const int4 q0, const int4 q1, SAcc &a)
{
    const int sfl = (q1.z >> 16) & 0xFFFF;
    const bool soft = sfl & S_SOFT_CLIP;
    /* q0 = l_tid, l_start, l_end, r_tid ; q1 = r_start, r_end, meta, ordinal */
    bool L = false, R = false;
    const bool lL = split_support(q0.x, q0.y, q0.z, k.tL, k.loL, k.hiL, k.rL);
    const bool lR = split_support(q0.x, q0.y, q0.z, k.tR, k.loR, k.hiR, k.rR);
    const bool rL = split_support(q0.w, q1.x, q1.y, k.tL, k.loL, k.hiL, k.rL);
    const bool rR = split_support(q0.w, q1.x, q1.y, k.tR, k.loR, k.hiR, k.rR);
    /* parsers.py:1163-1213 */
    if (!soft || k.svtype == SV_DEL) { L = lL; R = rR; }
    else if (k.svtype == SV_DUP) { L = lR; R = rL; }
    else if (k.svtype == SV_INV) { L = lL || lR; R = rL || rR; }
    const double x = L ? t.pm[q1.z & 0xFF] : 0.0;
    const double y = R ? t.pm[(q1.z >> 8) & 0xFF] : 0.0;
    const double p_alt = __dmul_rn(__dadd_rn(x, y), 0.5);      /* (.. + ..) / 2.0 is exact either way */
    if (ASSOC == SVGT_ASSOC_SSO) {
        if (sfl & S_FIRST) {
            a.alt_seq = __dadd_rn(a.alt_seq, a.sub_seq); a.alt_clip = __dadd_rn(a.alt_clip, a.sub_clip);
            a.sub_seq = 0.0; a.sub_clip = 0.0;
        }
        if (soft) a.sub_clip = __dadd_rn(a.sub_clip, p_alt); else a.sub_seq = __dadd_rn(a.sub_seq, p_alt);
    } else {
        if (soft) a.alt_clip = __dadd_rn(a.alt_clip, p_alt); else a.alt_seq = __dadd_rn(a.alt_seq, p_alt);
    }
}

/* ---------------- genotype call, singlesample.py:382-473 + statistics.py:9-37 ---------------- */
template <bool PIPE = false>   /* PIPE: log_choose's LUT values are fetched eight steps ahead (32 more registers) */
__device__ __forceinline__ void call_site(const SvgtParams &p, const Tables &t, int svtype, double ref_seq,
                                          double alt_seq, double alt_clip, double ref_span, double alt_span,
                                          svgt_out_row_t &o, int &err)
{
    /* zeroing rules, applied in order on already-modified values (singlesample.py:382-393) */
    if (__dadd_rn(alt_seq, alt_clip) < 0.5 && alt_span >= 1.0) { alt_seq = 0.0; alt_clip = 0.0; ref_seq = 0.0; }
    if (alt_span < 0.5 && __dadd_rn(alt_seq, alt_clip) >= 1.0) { alt_span = 0.0; ref_span = 0.0; }
    if (__dadd_rn(alt_span, alt_seq) == 0.0 && alt_clip > 0.0) alt_clip = 0.0;

    const double s1 = __dadd_rn(ref_seq, alt_seq);
    const double s2 = __dadd_rn(s1, ref_span);
    const double s3 = __dadd_rn(s2, alt_span);
    if (__dadd_rn(s3, alt_clip) == 0.0) { o.gt = SVGT_GT_BLANK; o.gq = -1; return; }

    const bool is_dup = svtype == SV_DUP;
    const double alt_splitters = __dadd_rn(alt_seq, alt_clip);
    const long long QR = __double2ll_rz(__dmul_rn(p.split_weight, ref_seq)) +
                         __double2ll_rz(__dmul_rn(p.disc_weight, ref_span));
    const long long QA = __double2ll_rz(__dmul_rn(p.split_weight, alt_splitters)) +
                         __double2ll_rz(__dmul_rn(p.disc_weight, alt_span));
    if (QR < 0 || QA < 0) { err = SVGT_ERR_ARG; o.gt = SVGT_GT_BLANK; o.gq = -1; return; }
    if (QR + QA >= p.n_log) { err = SVGT_ERR_LOG_TABLE; o.gt = SVGT_GT_BLANK; o.gq = -1; return; }

    /* log_choose(QR + QA, QA): the add/sub chain replayed in order on the host LUT */
    long long n = QR + QA, k = QA;
    if (k * 2 > n) k = n - k;
    double lc = 0.0;
    {
        /* r += log(n, 10); r -= log(d, 10); n -= 1 (statistics.py:14-18): 2 k dependent fp64 adds (8 cycles each on a
         * B200, scripts/ub/ub_fp64_latency.cu); the only freedom is when the LUT values are fetched -- eight steps
         * ahead of the adds that use them, so the chain never waits for a load */
        const double *ln = p.logt + n;
        const double *ld = p.logt + 1;
        long long d = 0;
        if (PIPE && k >= 16) {
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = __ldg(ln - i); b[i] = __ldg(ld + i); }
            for (; d + 16 <= k; d += 8) {
                double na[8], nb[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { na[i] = __ldg(ln - (d + 8 + i)); nb[i] = __ldg(ld + (d + 8 + i)); }
#pragma unroll
                for (int i = 0; i < 8; ++i) { lc = __dadd_rn(lc, a[i]); lc = __dsub_rn(lc, b[i]); }
#pragma unroll
                for (int i = 0; i < 8; ++i) { a[i] = na[i]; b[i] = nb[i]; }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { lc = __dadd_rn(lc, a[i]); lc = __dsub_rn(lc, b[i]); }
            d += 8;
        }
        for (; d < k; ++d) {
            lc = __dadd_rn(lc, __ldg(ln - d));
            lc = __dsub_rn(lc, __ldg(ld + d));
        }
    }
    const double *ca = p.consts + (is_dup ? C_DUP_ALT : C_NONDUP_ALT);
    const double *cr = p.consts + (is_dup ? C_DUP_REF : C_NONDUP_REF);
    double gl[3];
#pragma unroll
    for (int g = 0; g < 3; ++g)
        gl[g] = __dadd_rn(__dadd_rn(lc, __dmul_rn((double)QA, __ldg(ca + g))), __dmul_rn((double)QR, __ldg(cr + g)));

    /* stable descending sort: ties keep the lower index first */
    int best = 0;
    if (gl[1] > gl[best]) best = 1;
    if (gl[2] > gl[best]) best = 2;
    int second = -1;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        if (g == best) continue;
        if (second < 0 || gl[g] > gl[second]) second = g;
    }
    o.gl[0] = gl[0]; o.gl[1] = gl[1]; o.gl[2] = gl[2];
    o.dp = (int)__double2ll_rz(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(ref_seq, alt_seq), alt_clip), ref_span), alt_span));
    o.ro = (int)__double2ll_rz(__dadd_rn(ref_seq, ref_span));
    o.ao = (int)__double2ll_rz(__dadd_rn(__dadd_rn(alt_seq, alt_clip), alt_span));
    o.qr = (int)QR;
    o.qa = (int)QA;
    o.rs = (int)__double2ll_rz(ref_seq);
    o.as_ = (int)__double2ll_rz(alt_seq);
    o.asc = (int)__double2ll_rz(alt_clip);
    o.rp = (int)__double2ll_rz(ref_span);
    o.ap = (int)__double2ll_rz(alt_span);

    /* 10**gl underflows to 0.0 below the host-probed threshold */
    const double minx = __ldg(p.consts + C_POW10_MIN_X);
    if (gl[0] >= minx || gl[1] >= minx || gl[2] >= minx) {
        double gt_sum = 0.0;
#pragma unroll
        for (int g = 0; g < 3; ++g) gt_sum += (gl[g] >= minx) ? pow(10.0, gl[g]) : 0.0;
        if (!(gt_sum > 0.0)) gt_sum = 4.9406564584124654e-324;
        const double gt_sum_log = log(gt_sum) / 2.302585092994046;   /* math.log(x, 10) */
        o.sq = fabs(-10.0 * (gl[0] - gt_sum_log));
        double phred = __dmul_rn(-10.0, __dsub_rn(gl[second], gl[best]));
        if (phred > 200.0) phred = 200.0;
        o.gq = (int)__double2ll_rz(phred);
        o.gt = best;
    } else {
        o.gq = -1; o.sq = 0.0; o.gt = SVGT_GT_UNDERFLOW;
    }
}


/* coordinates the integer window arithmetic is exact for (SVGT_ERR_RANGE otherwise) */
__device__ __forceinline__ bool site_fields_in_range(const int4 a, const int4 b, int m, int slop)
{
    auto ok = [](int v, int lim) { return v > -lim && v < lim; };
    return ok(a.x, kRange) && ok(a.y, kRange) && ok(a.z, kCiRange) && ok(a.w, kCiRange) && ok(b.x, kCiRange) &&
           ok(b.y, kCiRange) && m >= 0 && m < (1 << 20) && slop >= 0 && slop < (1 << 20);
}

}  // namespace
