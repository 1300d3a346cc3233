/*
 * svgt_oracle.c -- CPU restatement of SVTyper's per-breakpoint scoring path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * svtyper_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load it; the product path never does (and has no CPU fallback).
 *
 * It restates, on the array layout of svtyper_b200/evidence.py, the algorithm of
 * hall-lab/svtyper v0.7.1 (all citations relative to /root/reference):
 *
 *   prob_mapq                      svtyper/utils.py:74-75
 *   SamFragment.is_ref_seq         svtyper/parsers.py:801-816
 *   SamFragment.is_pair_straddle   svtyper/parsers.py:821-857
 *   SamFragment.p_concordant       svtyper/parsers.py:861-882
 *   SplitRead.check_split_support  svtyper/parsers.py:1122-1134
 *   SplitRead.is_split_straddle    svtyper/parsers.py:1136-1215
 *   gather_split_read_evidence     svtyper/singlesample.py:246-276 (classic.py:306-332)
 *   gather_paired_end_evidence     svtyper/singlesample.py:278-353 (classic.py:339-408)
 *   tally_variant_read_fragments   svtyper/singlesample.py:355-404 (classic.py:286-435)
 *   bayesian_genotype              svtyper/singlesample.py:406-473 (classic.py:437-495)
 *   log_choose / bayes_gt          svtyper/statistics.py:9-37
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this restatement against the
 * per-site outputs of the reference itself on the reference's own fixture
 * (tests/data, 211 breakpoints; fixtures under tests/golden/) and against the
 * reference run live on synthetic batches when oracle/_ref is present.
 *
 * Transcendentals use the host libm exactly like CPython does
 * (math.log(x, 10) == log(x) / log(10.0); 10 ** y == pow(10.0, y)); everything else
 * is IEEE-754 double add/sub/mul/div in the reference's evaluation order.  Build
 * with -ffp-contract=off so the compiler never fuses a*b+c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define SITE_WORDS 16
#define FRAG_WORDS 8
#define SPLIT_WORDS 8

enum { SV_DEL = 0, SV_DUP = 1, SV_INV = 2, SV_BND = 3 };
enum { SITE_O1_REV = 1 << 2, SITE_O2_REV = 1 << 3, SITE_SKIP = 1 << 4 };
enum {
    F_HAS_A = 1 << 0, F_HAS_B = 1 << 1, F_REV_A = 1 << 2, F_REV_B = 1 << 3, F_PAIRED = 1 << 4,
    F_CONT = 1 << 5, F_EXTRA = 1 << 6, F_MULTI_A = 1 << 7, F_MULTI_B = 1 << 8
};
enum { S_SOFT_CLIP = 1 << 0, S_FIRST = 1 << 1 };
enum { GT_UNDERFLOW = -1, GT_BLANK = -2, GT_SKIPPED = -3 };
enum { ASSOC_SSO = 0, ASSOC_CLASSIC = 1 };

typedef struct {
    double gl[3];
    double sq;
    int32_t gt, gq, dp, ro, ao, qr, qa, rs, as_, asc, rp, ap;
} out_row_t;

typedef struct {
    const double *lib_f64;   /* [n_lib][4] flank, two_sd, N, mean */
    const int32_t *lib_i32;  /* [n_lib][4] hist_off, hist_len, nondel_L, 0 */
    const uint32_t *hist;
    int n_lib;
    double pm[256];
    double conc_prior, disc_prior;
    double c_alt[2][3], c_ref[2][3];
    int min_aligned, split_slop, assoc;
    double split_weight, disc_weight;
} ctx_t;

static double log10_py(double x) { return log(x) / log(10.0); }  /* math.log(x, 10) */

/* statistics.py:9-20 */
double svgt_oracle_log_choose(int64_t n, int64_t k)
{
    double r = 0.0;
    if (k * 2 > n) k = n - k;
    for (int64_t d = 1; d <= k; ++d) {
        r += log10_py((double)n);
        r -= log10_py((double)d);
        n -= 1;
    }
    return r;
}

/* statistics.py:23-37 */
void svgt_oracle_bayes_gt(int64_t ref, int64_t alt, int is_dup, double *lp)
{
    const double p_nondup[3] = {1e-3, 0.5, 0.9};
    const double p_dup[3] = {1e-2, 0.2, 1 / 3.0};
    const double *p = is_dup ? p_dup : p_nondup;
    double lc = svgt_oracle_log_choose(ref + alt, alt);
    for (int g = 0; g < 3; ++g) {
        double a = (double)alt * log10_py(p[g]);
        double b = (double)ref * log10_py(1 - p[g]);
        lp[g] = (lc + a) + b;
    }
}

/* utils.py:74-75 */
double svgt_oracle_prob_mapq(int q) { return 1 - pow(10.0, -(double)q / 10.0); }

/* parsers.py:801-816 for one gap-free aligned interval [s, e) of a read on `tid` */
static int ref_seq_hit(int32_t tid, int32_t s, int32_t e, int32_t site_tid, int64_t pos, int m)
{
    if (tid != site_tid) return 0;
    int64_t w0 = pos - m, w1 = pos + m;
    if (w0 < 0) return 0; /* max(0, pos-m) shortens the window below 2m */
    return (int64_t)s <= w0 && (int64_t)e >= w1;
}

/* parsers.py:821-857 */
static int pair_straddle(const int32_t *f, int32_t tA, int64_t posA, int64_t ciA0, int64_t ciA1,
                         int32_t tB, int64_t posB, int64_t ciB0, int64_t ciB1, int o1rev, int o2rev,
                         int m, double flank)
{
    int fl = f[7];
    if (!(fl & F_PAIRED)) return 0;
    if ((!!(fl & F_REV_A)) != o1rev) return 0;
    if ((!!(fl & F_REV_B)) != o2rev) return 0;
    if (f[4] != tA) return 0;
    if (f[5] != tB) return 0;
    int64_t i0 = (int64_t)f[0] + m;
    int64_t i1 = (int64_t)f[3] - m - 1;
    if (!o1rev && (i0 > posA + ciA1 || (double)i0 < (double)(posA + ciA0) - flank)) return 0;
    if (o1rev && (i0 < posA + ciA0 || (double)i0 > (double)(posA + ciA1) + flank)) return 0;
    if (!o2rev && (i1 > posB + ciB1 || (double)i1 < (double)(posB + ciB0) - flank)) return 0;
    if (o2rev && (i1 < posB + ciB0 || (double)i1 > (double)(posB + ciB1) + flank)) return 0;
    return 1;
}

static double dens(const ctx_t *c, int lib, int64_t key)
{
    const int32_t *li = c->lib_i32 + 4 * lib;
    if (key < 0 || key >= li[1]) return 0.0; /* Counter: missing key -> 0 */
    return (double)c->hist[li[0] + key] / c->lib_f64[4 * lib + 2];
}

/* parsers.py:861-882; returns the boolean `p > 0.5` (False on ZeroDivisionError) */
static int p_concordant(const ctx_t *c, const int32_t *f, int lib, int is_del, int64_t var_length)
{
    int64_t o = (int64_t)f[3] - (int64_t)f[0];
    if (o < 0) o = -o;
    double d1 = dens(c, lib, o), d2;
    if (is_del) {
        d2 = dens(c, lib, o - var_length);
    } else {
        int64_t L = c->lib_i32[4 * lib + 2]; /* integral mean+3sd, or -1: float key never hits */
        d2 = (L >= 0) ? dens(c, lib, o - L) : 0.0;
    }
    double num = d1 * c->conc_prior;
    double den = c->conc_prior * d1 + c->disc_prior * d2;
    if (den == 0.0) return 0;
    return (num / den) > 0.5;
}

/* parsers.py:1122-1134 */
static int split_support(int32_t tid, int32_t start, int32_t end, int32_t site_tid, int64_t pos,
                         int is_rev, int slop)
{
    if (tid != site_tid) return 0;
    int64_t coord = is_rev ? start : end;
    if (coord > pos + slop || coord < pos - slop) return 0;
    return 1;
}

static void score_site(const ctx_t *c, const int32_t *s, const int32_t *frags, const int32_t *splits,
                       out_row_t *o)
{
    memset(o, 0, sizeof(*o));
    const int meta = s[9];
    if (meta & SITE_SKIP) { o->gt = GT_SKIPPED; o->gq = -1; return; }
    const int svtype = meta & 3;
    const int o1rev = !!(meta & SITE_O1_REV), o2rev = !!(meta & SITE_O2_REV);
    const int64_t posA = s[0], posB = s[1], ciA0 = s[2], ciA1 = s[3], ciB0 = s[4], ciB1 = s[5];
    const int32_t tA = s[6], tB = s[7];
    const int64_t var_length = s[8];
    int64_t foff, soff;
    memcpy(&foff, s + 10, 8);
    memcpy(&soff, s + 13, 8);
    const int nf = s[12], ns = s[15];
    const int m = c->min_aligned;
    const int is_del = svtype == SV_DEL;

    double ref_seq = 0, alt_seq = 0, alt_clip = 0, ref_span = 0, alt_span = 0;

    /* ---- fragment rows: ref_seq (split-read reference support) + paired-end evidence ---- */
    double sub_ref = 0.0;
    int pendA = 0, pendB = 0;
    for (int j = 0; j < nf; ++j) {
        const int32_t *f = frags + (foff + j) * FRAG_WORDS;
        const int fl = f[7];
        int hitA = 0, hitB = 0;
        if (fl & F_HAS_A)
            hitA = ref_seq_hit(f[4], f[0], f[1], tA, posA, m) || ref_seq_hit(f[4], f[0], f[1], tB, posB, m);
        if (fl & F_HAS_B)
            hitB = ref_seq_hit(f[5], f[2], f[3], tA, posA, m) || ref_seq_hit(f[5], f[2], f[3], tB, posB, m);
        if (fl & F_EXTRA) { pendA |= hitA; pendB |= hitB; continue; }
        if (fl & F_MULTI_A) hitA = pendA;
        if (fl & F_MULTI_B) hitB = pendB;
        pendA = pendB = 0;
        const double pmA = c->pm[f[6] & 0xFF], pmB = c->pm[(f[6] >> 8) & 0xFF];
        const int lib = (f[6] >> 16) & 0xFFFF;

        /* singlesample.py:254-259 per-fragment sub-total; classic.py:306-311 adds straight in */
        if (c->assoc == ASSOC_SSO) {
            if (!(fl & F_CONT)) { ref_seq += sub_ref; sub_ref = 0.0; }
            if ((fl & F_HAS_A) && hitA) sub_ref += pmA;
            if ((fl & F_HAS_B) && hitB) sub_ref += pmB;
        } else {
            if ((fl & F_HAS_A) && hitA) ref_seq += pmA;
            if ((fl & F_HAS_B) && hitB) ref_seq += pmB;
        }

        if (!(fl & F_PAIRED) || lib >= c->n_lib) continue;
        const double flank = c->lib_f64[4 * lib + 0], two_sd = c->lib_f64[4 * lib + 1];
        /* singlesample.py:289 / :328 -- positions are post-increment here */
        const int small_del = is_del && ((double)(posB - posA) < two_sd);
        int alt = 0, recip = 0;
        if (!small_del)
            alt = pair_straddle(f, tA, posA, ciA0, ciA1, tB, posB, ciB0, ciB1, o1rev, o2rev, m, flank);
        if (svtype == SV_INV)
            recip = pair_straddle(f, tA, posA, ciA0, ciA1, tB, posB, ciB0, ciB1, !o1rev, !o2rev, m, flank);
        if (alt || recip) {
            if (is_del) {
                int pc = p_concordant(c, f, lib, 1, var_length);
                double p_alt = pc ? 0.0 : (pmA * pmB); /* (1 - p_conc) * pmA * pmB */
                alt_span += p_alt;
            } else {
                alt_span += pmA * pmB;
            }
        }
        int refA = 0, refB = 0;
        if (!small_del) {
            refA = pair_straddle(f, tA, posA, 0, 0, tA, posA, 0, 0, 0, 1, m, flank);
            refB = pair_straddle(f, tB, posB, 0, 0, tB, posB, 0, 0, 0, 1, m, flank);
        }
        if (refA || refB) {
            if (!(refA && refB) || is_del) {
                int pc = p_concordant(c, f, lib, is_del, var_length);
                double p_ref = pc ? (pmA * pmB) : 0.0; /* p_conc * pmA * pmB */
                ref_span += (double)(refA + refB) * p_ref / 2;
            }
        }
    }
    ref_seq += sub_ref; /* adding 0.0 is a no-op in classic mode */

    /* ---- split rows (singlesample.py:262-274, classic.py:317-328) ---- */
    {
        /* arrange breakends left to right, parsers.py:1143-1161 */
        int32_t tL, tR; int64_t pL, pR; int rL, rR;
        if (tA != tB || posA > posB) { tL = tB; pL = posB; rL = o2rev; tR = tA; pR = posA; rR = o1rev; }
        else { tL = tA; pL = posA; rL = o1rev; tR = tB; pR = posB; rR = o2rev; }
        double sub_seq = 0.0, sub_clip = 0.0;
        const int slop = c->split_slop;
        for (int j = 0; j < ns; ++j) {
            const int32_t *q = splits + (soff + j) * SPLIT_WORDS;
            const int sfl = (q[6] >> 16) & 0xFFFF;
            const int soft = sfl & S_SOFT_CLIP;
            int L = 0, R = 0;
            if (!soft || svtype == SV_DEL) {
                L = split_support(q[0], q[1], q[2], tL, pL, rL, slop);
                R = split_support(q[3], q[4], q[5], tR, pR, rR, slop);
            } else if (svtype == SV_DUP) {
                L = split_support(q[0], q[1], q[2], tR, pR, rR, slop);
                R = split_support(q[3], q[4], q[5], tL, pL, rL, slop);
            } else if (svtype == SV_INV) {
                L = split_support(q[0], q[1], q[2], tL, pL, rL, slop) || split_support(q[0], q[1], q[2], tR, pR, rR, slop);
                R = split_support(q[3], q[4], q[5], tL, pL, rL, slop) || split_support(q[3], q[4], q[5], tR, pR, rR, slop);
            } /* soft-clipped BND: (False, False) */
            double a = L ? c->pm[q[6] & 0xFF] : 0.0;
            double b = R ? c->pm[(q[6] >> 8) & 0xFF] : 0.0;
            double p_alt = (a + b) / 2.0;
            if (c->assoc == ASSOC_SSO) {
                if (sfl & S_FIRST) { alt_seq += sub_seq; alt_clip += sub_clip; sub_seq = sub_clip = 0.0; }
                if (soft) sub_clip += p_alt; else sub_seq += p_alt;
            } else {
                if (soft) alt_clip += p_alt; else alt_seq += p_alt;
            }
        }
        alt_seq += sub_seq;
        alt_clip += sub_clip;
    }

    /* ---- zeroing rules, singlesample.py:382-393 ---- */
    if ((alt_seq + alt_clip) < 0.5 && alt_span >= 1) { alt_seq = 0; alt_clip = 0; ref_seq = 0; }
    if (alt_span < 0.5 && (alt_seq + alt_clip) >= 1) { alt_span = 0; ref_span = 0; }
    if (alt_span + alt_seq == 0 && alt_clip > 0) alt_clip = 0;

    /* singlesample.py:494-496 / classic.py:437,496 */
    if (ref_seq + alt_seq + ref_span + alt_span + alt_clip == 0) { o->gt = GT_BLANK; o->gq = -1; return; }

    /* ---- bayesian_genotype, singlesample.py:406-473 ---- */
    const int is_dup = svtype == SV_DUP;
    const double alt_splitters = alt_seq + alt_clip;
    const int64_t QR = (int64_t)(c->split_weight * ref_seq) + (int64_t)(c->disc_weight * ref_span);
    const int64_t QA = (int64_t)(c->split_weight * alt_splitters) + (int64_t)(c->disc_weight * alt_span);
    double lc = svgt_oracle_log_choose(QR + QA, QA);
    for (int g = 0; g < 3; ++g)
        o->gl[g] = (lc + (double)QA * c->c_alt[is_dup][g]) + (double)QR * c->c_ref[is_dup][g];
    /* stable descending sort -> best / second best (ties keep the lower index first) */
    int best = 0;
    for (int g = 1; g < 3; ++g) if (o->gl[g] > o->gl[best]) best = g;
    int second = -1;
    for (int g = 0; g < 3; ++g) {
        if (g == best) continue;
        if (second < 0 || o->gl[g] > o->gl[second]) second = g;
    }
    o->dp = (int32_t)(ref_seq + alt_seq + alt_clip + ref_span + alt_span);
    o->ro = (int32_t)(ref_seq + ref_span);
    o->ao = (int32_t)(alt_seq + alt_clip + alt_span);
    o->qr = (int32_t)QR;
    o->qa = (int32_t)QA;
    o->rs = (int32_t)ref_seq;
    o->as_ = (int32_t)alt_seq;
    o->asc = (int32_t)alt_clip;
    o->rp = (int32_t)ref_span;
    o->ap = (int32_t)alt_span;
    double gt_sum = 0;
    for (int g = 0; g < 3; ++g) gt_sum += pow(10.0, o->gl[g]);
    if (gt_sum > 0) {
        double gt_sum_log = log10_py(gt_sum);
        o->sq = fabs(-10 * (o->gl[0] - gt_sum_log));
        double phred = -10 * (o->gl[second] - o->gl[best]);
        if (phred > 200) phred = 200;
        o->gq = (int32_t)phred;
        o->gt = best;
    } else {
        o->gq = -1;
        o->sq = 0.0;
        o->gt = GT_UNDERFLOW;
    }
}

typedef struct {
    const ctx_t *c;
    const int32_t *sites, *frags, *splits;
    out_row_t *out;
    int64_t n_sites;
    int64_t next; /* atomic work cursor */
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    const int64_t chunk = 32;
    for (;;) {
        int64_t lo = __atomic_fetch_add(&j->next, chunk, __ATOMIC_RELAXED);
        if (lo >= j->n_sites) break;
        int64_t hi = lo + chunk < j->n_sites ? lo + chunk : j->n_sites;
        for (int64_t i = lo; i < hi; ++i)
            score_site(j->c, j->sites + i * SITE_WORDS, j->frags, j->splits, j->out + i);
    }
    return NULL;
}

int svgt_oracle_score(const int32_t *sites, int64_t n_sites, const int32_t *frags, const int32_t *splits,
                      const double *lib_f64, const int32_t *lib_i32, int n_lib, const uint32_t *hist,
                      int min_aligned, int split_slop, double split_weight, double disc_weight,
                      int assoc_mode, void *out_rows, int n_threads)
{
    ctx_t c;
    memset(&c, 0, sizeof(c));
    c.lib_f64 = lib_f64; c.lib_i32 = lib_i32; c.hist = hist; c.n_lib = n_lib;
    for (int q = 0; q < 256; ++q) c.pm[q] = svgt_oracle_prob_mapq(q);
    c.disc_prior = 0.05;
    c.conc_prior = 1 - c.disc_prior;
    {
        const double p_nondup[3] = {1e-3, 0.5, 0.9};
        const double p_dup[3] = {1e-2, 0.2, 1 / 3.0};
        for (int g = 0; g < 3; ++g) {
            c.c_alt[0][g] = log10_py(p_nondup[g]); c.c_ref[0][g] = log10_py(1 - p_nondup[g]);
            c.c_alt[1][g] = log10_py(p_dup[g]);    c.c_ref[1][g] = log10_py(1 - p_dup[g]);
        }
    }
    c.min_aligned = min_aligned; c.split_slop = split_slop; c.assoc = assoc_mode;
    c.split_weight = split_weight; c.disc_weight = disc_weight;
    out_row_t *out = (out_row_t *)out_rows;
    if (sizeof(out_row_t) != 80) return -1;
    job_t job = {&c, sites, frags, splits, out, n_sites, 0};
    if (n_threads <= 1) { worker(&job); return 0; }
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < n_threads - 1; ++t)
        if (pthread_create(&th[started], NULL, worker, &job) == 0) ++started;
    worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    return 0;
}

int svgt_oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
