/*
 * svgt_coop.cu -- warp-cooperative scoring kernel (variant 2, the default).
 *
 * Same arithmetic and results as the thread-per-site kernels in svgt_kernels.cu (reference
 * citations there); different mapping.  The only order-sensitive part of the path is the fp64
 * accumulation of per-fragment weights in sorted(qname) order; everything that
 * produces those weights -- is_ref_seq, is_pair_straddle x4, p_concordant, prob_mapq products
 * (svtyper/parsers.py:801-882, singlesample.py:246-353) -- is independent per row.  So:
 *
 *   phase A  one ROW per lane: a warp reads 32 consecutive 32-byte rows of one site with fully
 *            coalesced 128-bit loads (1 KB per request pair), scores them with warp-uniform site
 *            constants (per-(site, library) integer windows precomputed in shared memory), and
 *            parks 4 doubles per row {a, b, p_ref, p_alt} in shared memory.  Cross-row state of
 *            the schema (EXTRA/MULTI interval rows) is resolved with ballots.
 *   phase B  one CHAIN per lane: G sites are interleaved per warp, so lane 4g+c replays chain c
 *            (ref_seq, ref_span, alt_span | alt_seq, alt_clip) of site g over that site's 32
 *            parked rows in order -- the serial adds of G*3 chains issue together instead of
 *            one per instruction.
 *   epilogue the five sums of 32 sites (4 rounds of G = 8) are shuffled back to one lane per
 *            site and bayesian_genotype / bayes_gt / log_choose (singlesample.py:406-473,
 *            statistics.py:9-37) run lane-parallel.
 *
 * A site is thus walked at 32 rows per warp step instead of one row per thread step: the
 * longest site no longer bounds the launch, and DRAM sees 1 KB sequential bursts.
 */
#include "svgt_device.cuh"

namespace {

constexpr int kG = 8;               /* sites interleaved per warp                          */
constexpr int kWLibs = 4;           /* libraries with per-site windows cached in smem      */
constexpr int kCoopWarps = SVGT_COOP_THREADS / 32;

/* per-(warp, g) site scalars, read warp-uniformly */
struct SiteS {
    int tA, tB, wA0, wA1, wB0, wB1, meta, var_length;
    int posA, posB, ciA0, ciA1, ciB0, ciB1, dAB, nf;
    long long foff, soff;
    int ns, pad0, pad1, pad2;
};  /* 96 B */

/* per-(warp, g, library) windows as (lo, width+1): pass iff (unsigned)(v - lo) < w1 */
struct Win {
    unsigned altA_lo, altA_w1, altB_lo, altB_w1;
    unsigned recA_lo, recA_w1, recB_lo, recB_w1;
    unsigned rAa_lo, rAa_w1, rAb_lo, rAb_w1;
    unsigned rBa_lo, rBa_w1, rBb_lo, rBb_w1;
};  /* 64 B */

struct WarpSmem {
    SiteS site[kG];
    Win win[kG][kWLibs];
    double contrib[kG][32][4];
    unsigned newmask[kG];
    double zero[2];
};

__device__ __forceinline__ void set_win(unsigned &lo_out, unsigned &w1_out, int lo, int hi, bool enable)
{
    lo_out = (unsigned)lo;
    w1_out = (enable && hi >= lo) ? (unsigned)(hi - lo) + 1u : 0u;
}

__device__ __forceinline__ Win make_win(const SiteS &S, const LibK &L, int m)
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
    return w;
}

__device__ __forceinline__ bool in_win(int v, unsigned lo, unsigned w1) { return ((unsigned)v - lo) < w1; }

/* p_concordant() on the integer path (safe library); falls to the literal expression on a tie */
__device__ __forceinline__ bool p_conc_fast(const Tables &t, const LibK &L, int a_start, int b_end, bool is_del,
                                            int var_length)
{
    const unsigned o = b_end >= a_start ? (unsigned)b_end - (unsigned)a_start : (unsigned)a_start - (unsigned)b_end;
    const unsigned hl = (unsigned)L.hist_len;
    const unsigned h1 = o < hl ? t.hist[L.hist_off + o] : 0u;
    const int Lk = is_del ? var_length : L.nondel_L;
    unsigned h2 = 0u;
    if (is_del || Lk >= 0) {
        if (Lk >= 0) {
            const unsigned k2 = o - (unsigned)Lk;
            if (o >= (unsigned)Lk && k2 < hl) h2 = t.hist[L.hist_off + k2];
        } else {
            const long long k2 = (long long)o - (long long)Lk;
            if (k2 < (long long)hl) h2 = t.hist[L.hist_off + (int)k2];
        }
    }
    const unsigned long long l19 = 19ull * h1;
    if (l19 != (unsigned long long)h2) return l19 > (unsigned long long)h2;
    if (h1 == 0u) return false;
    return p_conc_literal(t, L, h1, h2);
}

/* ordered replay of one chain over `cnt` parked rows (phase B).
 * SSO:     per row   if NEW: acc += pend, pend = 0;   pend = (pend + x) + y
 * CLASSIC: per row   acc = (acc + x) + y
 * NEW rows dominate, so the all-NEW case is a 2-add loop. */
template <int ASSOC>
__device__ __forceinline__ void replay_chain(const double *px, const double *py, int xstride, int ystride, int cnt,
                                             unsigned newm, bool all_new, double &acc, double &pend)
{
    if (ASSOC == SVGT_ASSOC_CLASSIC) {
        for (int j = 0; j < cnt; ++j)
            acc = __dadd_rn(__dadd_rn(acc, px[j * xstride]), py[j * ystride]);
    } else if (all_new) {
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const double tsum = __dadd_rn(px[j * xstride], py[j * ystride]);
            acc = __dadd_rn(acc, pend);
            pend = tsum;
        }
    } else {
        for (int j = 0; j < cnt; ++j) {
            const bool nw = (newm >> j) & 1u;
            const double u = nw ? pend : 0.0;
            const double t0 = nw ? 0.0 : pend;
            acc = __dadd_rn(acc, u);
            pend = __dadd_rn(__dadd_rn(t0, px[j * xstride]), py[j * ystride]);
        }
    }
}

template <int ASSOC>
__global__ void __launch_bounds__(SVGT_COOP_THREADS, 2) svgt_coop_kernel(const SvgtParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_pm = reinterpret_cast<double *>(smem_raw);
    LibK *s_lib = reinterpret_cast<LibK *>(s_pm + 256);
    size_t off = 256 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    off = (off + 127) & ~(size_t)127;
    WarpSmem *s_warp = reinterpret_cast<WarpSmem *>(smem_raw + off);
    off += sizeof(WarpSmem) * kCoopWarps;
    unsigned *s_hist = reinterpret_cast<unsigned *>(smem_raw + off);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    int err = 0;
    const int nl = p.n_lib < SVGT_SMEM_LIBS ? p.n_lib : SVGT_SMEM_LIBS;
    for (int i = tid; i < 256; i += SVGT_COOP_THREADS) s_pm[i] = p.pm[i];
    for (int i = tid; i < nl; i += SVGT_COOP_THREADS) s_lib[i] = derive_lib(p, i, &err);
    if (p.hist_in_smem)
        for (int i = tid; i < (int)p.n_hist; i += SVGT_COOP_THREADS) s_hist[i] = p.hist[i];
    WarpSmem &ws = s_warp[warp];
    if (lane < 2) ws.zero[lane] = 0.0;
    __syncthreads();

    Tables t;
    t.pm = s_pm; t.libs = s_lib; t.hist = p.hist_in_smem ? s_hist : p.hist;
    t.conc = p.consts[C_CONC]; t.disc = p.consts[C_DISC];
    const int m = p.min_aligned, slop = p.split_slop;
    /* libraries 0..31 whose integer rewrites are exact */
    unsigned safemask = 0u;
    for (int l = 0; l < nl && l < 32; ++l) safemask |= (s_lib[l].safe ? 1u : 0u) << l;

    /* phase-B role of this lane: chain c of interleaved site gb */
    const int gb = lane >> 2, c = lane & 3;

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(p.status + 1, 1);
        tile = __shfl_sync(full, tile, 0);
        if (tile >= p.n_tiles) break;
        const long long idx = (long long)tile * 32 + lane;
        const bool valid = idx < p.n_sites;
        long long site = 0;
        if (valid) site = p.order ? (long long)p.order[idx] : idx;
        int4 a = make_int4(0, 0, 0, 0), b = a, cc = a, d = a;
        if (valid) {
            const int4 *sp = p.sites + site * 4;
            a = ldg4(sp); b = ldg4(sp + 1); cc = ldg4(sp + 2); d = ldg4(sp + 3);
        }
        const int meta = cc.y;
        const bool skip = !valid || (meta & SITE_SKIP);
        bool ranged = (a.x > -kRange && a.x < kRange && a.y > -kRange && a.y < kRange && a.z > -kCiRange &&
                       a.z < kCiRange && a.w > -kCiRange && a.w < kCiRange && b.x > -kCiRange && b.x < kCiRange &&
                       b.y > -kCiRange && b.y < kCiRange && m >= 0 && m < (1 << 20) && slop >= 0 && slop < (1 << 20));
        const bool run = !skip && ranged;
        long long foff = ((long long)(unsigned)cc.z) | ((long long)cc.w << 32);
        long long soff = ((long long)(unsigned)d.y) | ((long long)d.z << 32);
        int nf = run ? d.x : 0, ns = run ? d.w : 0;
        if (nf < 0 || foff < 0 || foff + nf > p.n_frag) { nf = 0; err = SVGT_ERR_ARG; }
        if (ns < 0 || soff < 0 || soff + ns > p.n_split) { ns = 0; err = SVGT_ERR_ARG; }

        double r_ref_seq = 0.0, r_ref_span = 0.0, r_alt_span = 0.0, r_alt_seq = 0.0, r_alt_clip = 0.0;

        for (int round = 0; round < 32 / kG; ++round) {
            const int base_lane = round * kG;
            /* skip rounds with no work at all */
            const unsigned work = __ballot_sync(full, (nf | ns) != 0);
            if (((work >> base_lane) & ((1u << kG) - 1u)) == 0u) continue;

            /* ---- publish site scalars and per-(site, library) windows ---- */
            if (lane >= base_lane && lane < base_lane + kG) {
                SiteS &S = ws.site[lane - base_lane];
                S.tA = b.z; S.tB = b.w;
                S.wA0 = a.x - m; S.wA1 = a.x + m; S.wB0 = a.y - m; S.wB1 = a.y + m;
                S.meta = (meta & 15) | ((a.x - m >= 0) << 8) | ((a.y - m >= 0) << 9);
                S.var_length = cc.x;
                S.posA = a.x; S.posB = a.y; S.ciA0 = a.z; S.ciA1 = a.w; S.ciB0 = b.x; S.ciB1 = b.y;
                S.dAB = a.y - a.x; S.nf = nf; S.foff = foff; S.soff = soff; S.ns = ns;
            }
            __syncwarp();
            for (int i = lane; i < kG * kWLibs; i += 32) {
                const int g = i / kWLibs, l = i % kWLibs;
                if (l < nl) ws.win[g][l] = make_win(ws.site[g], s_lib[l], m);
            }
            __syncwarp();

            /* ================= fragment rows ================= */
            int nfmax = 0;
#pragma unroll
            for (int g = 0; g < kG; ++g) nfmax = max(nfmax, ws.site[g].nf);
            double acc = 0.0, pend = 0.0;
            unsigned carryA = 0u, carryB = 0u;          /* EXTRA-run hits carried into the next step, bit g */
            for (int step = 0; step * 32 < nfmax; ++step) {
                bool all_new = true;
                for (int g = 0; g < kG; ++g) {
                    const SiteS &S = ws.site[g];
                    const int n = min(32, S.nf - step * 32);
                    if (n <= 0) continue;                       /* warp-uniform */
                    const bool rv = lane < n;
                    int4 lo = make_int4(0, 0, 0, 0), hi = lo;
                    if (rv) {
                        const int4 *rp = p.frags + 2 * (S.foff + (long long)step * 32 + lane);
                        lo = ldg4(rp); hi = ldg4(rp + 1);
                    }
                    const int fl = rv ? hi.w : 0;
                    const bool okA = (S.meta >> 8) & 1, okB = (S.meta >> 9) & 1;
                    const bool ea = hi.x == S.tA, eb = hi.x == S.tB, fa = hi.y == S.tA, fb = hi.y == S.tB;
                    bool hitA = (fl & F_HAS_A) && ((ea && okA && lo.x <= S.wA0 && lo.y >= S.wA1) ||
                                                   (eb && okB && lo.x <= S.wB0 && lo.y >= S.wB1));
                    bool hitB = (fl & F_HAS_B) && ((fa && okA && lo.z <= S.wA0 && lo.w >= S.wA1) ||
                                                   (fb && okB && lo.z <= S.wB0 && lo.w >= S.wB1));
                    /* EXTRA interval rows feed the next main row's MULTI slots (evidence.py) */
                    const bool isx = (fl & F_EXTRA) != 0;
                    const unsigned E = __ballot_sync(full, isx);
                    const bool cA = (carryA >> g) & 1u, cB = (carryB >> g) & 1u;
                    if (E != 0u || cA || cB) {
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
                        const unsigned vm = n == 32 ? full : ((1u << n) - 1u);
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
                    } else {
                        if (fl & F_MULTI_A) hitA = false;
                        if (fl & F_MULTI_B) hitB = false;
                    }
                    const unsigned nm = __ballot_sync(full, rv && !(fl & (F_CONT | F_EXTRA)));
                    const unsigned vm2 = n == 32 ? full : ((1u << n) - 1u);
                    all_new = all_new && (nm == vm2);

                    const double pmA = t.pm[hi.z & 0xFF], pmB = t.pm[(hi.z >> 8) & 0xFF];
                    const int lib = (hi.z >> 16) & 0xFFFF;
                    double va = 0.0, vb = 0.0, p_ref = 0.0, p_alt = 0.0;
                    if (!isx) {
                        va = ((fl & F_HAS_A) && hitA) ? pmA : 0.0;
                        vb = ((fl & F_HAS_B) && hitB) ? pmB : 0.0;
                    }
                    if ((fl & F_PAIRED) && !isx) {
                        if (lib >= p.n_lib) err = SVGT_ERR_LIB_INDEX;
                        else {
                            LibK Ls;
                            if (lib >= SVGT_SMEM_LIBS) { int e = 0; Ls = derive_lib(p, lib, &e); }
                            const LibK &L = (lib < SVGT_SMEM_LIBS) ? s_lib[lib] : Ls;
                            const bool safe = lib < 32 ? ((safemask >> lib) & 1u) : (L.safe != 0);
                            const int svtype = S.meta & 3;
                            const bool is_del = svtype == SV_DEL;
                            bool alt, recip = false, refA, refB;
                            if (safe) {
                                Win wl;
                                if (lib >= kWLibs) wl = make_win(S, L, m);
                                const Win &w = (lib < kWLibs) ? ws.win[g][lib] : wl;
                                const int st = (fl >> 2) & 3, o12 = (S.meta >> 2) & 3;
                                const bool ab = ea && fb;
                                alt = ab && st == o12 && in_win(lo.x, w.altA_lo, w.altA_w1) && in_win(lo.w, w.altB_lo, w.altB_w1);
                                recip = ab && st == (o12 ^ 3) && in_win(lo.x, w.recA_lo, w.recA_w1) &&
                                        in_win(lo.w, w.recB_lo, w.recB_w1);
                                const bool fr = st == 2;        /* readA forward, readB reverse */
                                refA = fr && ea && fa && in_win(lo.x, w.rAa_lo, w.rAa_w1) && in_win(lo.w, w.rAb_lo, w.rAb_w1);
                                refB = fr && eb && fb && in_win(lo.x, w.rBa_lo, w.rBa_w1) && in_win(lo.w, w.rBb_lo, w.rBb_w1);
                            } else {
                                const int o1 = (S.meta >> 2) & 1, o2 = (S.meta >> 3) & 1;
                                const bool small_del = is_del && ((double)((long long)S.posB - S.posA) < L.two_sd);
                                alt = !small_del && straddle_literal(lo, hi, S.tA, S.posA, S.ciA0, S.ciA1, S.tB, S.posB,
                                                                     S.ciB0, S.ciB1, o1, o2, m, L.flank);
                                if (svtype == SV_INV)
                                    recip = straddle_literal(lo, hi, S.tA, S.posA, S.ciA0, S.ciA1, S.tB, S.posB, S.ciB0,
                                                             S.ciB1, !o1, !o2, m, L.flank);
                                refA = !small_del && straddle_literal(lo, hi, S.tA, S.posA, 0, 0, S.tA, S.posA, 0, 0, 0, 1, m, L.flank);
                                refB = !small_del && straddle_literal(lo, hi, S.tB, S.posB, 0, 0, S.tB, S.posB, 0, 0, 0, 1, m, L.flank);
                            }
                            const bool is_alt = alt || recip;
                            const bool use_ref = (refA || refB) && (!(refA && refB) || is_del);
                            if (is_alt || use_ref) {
                                bool pc = false;
                                if ((is_alt && is_del) || use_ref)
                                    pc = safe ? p_conc_fast(t, L, lo.x, lo.w, is_del, S.var_length)
                                              : p_concordant(t, L, lo.x, lo.w, is_del, S.var_length);
                                const double prod = __dmul_rn(pmA, pmB);
                                p_alt = is_alt ? ((is_del && pc) ? 0.0 : prod) : 0.0;
                                p_ref = (use_ref && pc) ? prod : 0.0;
                                if (!(refA && refB)) p_ref = __dmul_rn(p_ref, 0.5);
                            }
                        }
                    }
                    double4 *dst = reinterpret_cast<double4 *>(&ws.contrib[g][lane][0]);
                    *dst = make_double4(va, vb, p_ref, p_alt);
                    if (lane == 0) ws.newmask[g] = nm;
                }
                __syncwarp();
                /* ---- phase B: lane 4g+c replays chain c of site g ---- */
                if (gb < kG && c < 3) {
                    int cnt = ws.site[gb].nf - step * 32;
                    cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
                    const double *row0 = &ws.contrib[gb][0][0];
                    const double *px = row0 + (c == 0 ? 0 : c + 1);
                    const double *py = (c == 0) ? row0 + 1 : ws.zero;
                    replay_chain<ASSOC>(px, py, 4, c == 0 ? 4 : 0, cnt, ws.newmask[gb], all_new, acc, pend);
                }
                __syncwarp();
            }
            if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
            {
                /* hand the three sums to the site's own lane */
                const int mine = lane - base_lane;
                const bool take = mine >= 0 && mine < kG;
                const int src = take ? 4 * mine : 0;
                const double v0 = __shfl_sync(full, acc, src), v1 = __shfl_sync(full, acc, src + 1),
                             v2 = __shfl_sync(full, acc, src + 2);
                if (take) { r_ref_seq = v0; r_ref_span = v1; r_alt_span = v2; }
            }

            /* ================= split rows ================= */
            int nsmax = 0;
#pragma unroll
            for (int g = 0; g < kG; ++g) nsmax = max(nsmax, ws.site[g].ns);
            acc = 0.0; pend = 0.0;
            for (int step = 0; step * 32 < nsmax; ++step) {
                bool all_new = true;
                for (int g = 0; g < kG; ++g) {
                    const SiteS &S = ws.site[g];
                    const int n = min(32, S.ns - step * 32);
                    if (n <= 0) continue;
                    const bool rv = lane < n;
                    int4 q0 = make_int4(0, 0, 0, 0), q1 = q0;
                    if (rv) {
                        const int4 *rp = p.splits + 2 * (S.soff + (long long)step * 32 + lane);
                        q0 = ldg4(rp); q1 = ldg4(rp + 1);
                    }
                    /* arrange breakends left to right, parsers.py:1143-1161 */
                    const int o1 = (S.meta >> 2) & 1, o2 = (S.meta >> 3) & 1, svtype = S.meta & 3;
                    const bool swap = (S.tA != S.tB) || (S.posA > S.posB);
                    const int tL = swap ? S.tB : S.tA, tR = swap ? S.tA : S.tB;
                    const int pL = swap ? S.posB : S.posA, pR = swap ? S.posA : S.posB;
                    const int rL = swap ? o2 : o1, rR = swap ? o1 : o2;
                    const int sfl = (q1.z >> 16) & 0xFFFF;
                    const bool soft = sfl & S_SOFT_CLIP;
                    const bool lL = split_support(q0.x, q0.y, q0.z, tL, pL - slop, pL + slop, rL);
                    const bool lR = split_support(q0.x, q0.y, q0.z, tR, pR - slop, pR + slop, rR);
                    const bool rLs = split_support(q0.w, q1.x, q1.y, tL, pL - slop, pL + slop, rL);
                    const bool rRs = split_support(q0.w, q1.x, q1.y, tR, pR - slop, pR + slop, rR);
                    bool Ls = false, Rs = false;
                    if (!soft || svtype == SV_DEL) { Ls = lL; Rs = rRs; }
                    else if (svtype == SV_DUP) { Ls = lR; Rs = rLs; }
                    else if (svtype == SV_INV) { Ls = lL || lR; Rs = rLs || rRs; }
                    const double x = Ls ? t.pm[q1.z & 0xFF] : 0.0;
                    const double y = Rs ? t.pm[(q1.z >> 8) & 0xFF] : 0.0;
                    double p_alt = __dmul_rn(__dadd_rn(x, y), 0.5);
                    if (!rv) p_alt = 0.0;
                    const unsigned nm = __ballot_sync(full, rv && (sfl & S_FIRST));
                    const unsigned vm2 = n == 32 ? full : ((1u << n) - 1u);
                    all_new = all_new && (nm == vm2);
                    double2 *dst = reinterpret_cast<double2 *>(&ws.contrib[g][lane][0]);
                    *dst = make_double2(soft ? 0.0 : p_alt, soft ? p_alt : 0.0);
                    if (lane == 0) ws.newmask[g] = nm;
                }
                __syncwarp();
                if (gb < kG && c < 2) {
                    int cnt = ws.site[gb].ns - step * 32;
                    cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
                    const double *px = &ws.contrib[gb][0][0] + c;
                    replay_chain<ASSOC>(px, ws.zero, 4, 0, cnt, ws.newmask[gb], all_new, acc, pend);
                }
                __syncwarp();
            }
            if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
            {
                const int mine = lane - base_lane;
                const bool take = mine >= 0 && mine < kG;
                const int src = take ? 4 * mine : 0;
                const double v0 = __shfl_sync(full, acc, src), v1 = __shfl_sync(full, acc, src + 1);
                if (take) { r_alt_seq = v0; r_alt_clip = v1; }
            }
            __syncwarp();
        }

        /* ================= lane-parallel genotype call ================= */
        if (valid) {
            svgt_out_row_t o;
            o.gl[0] = o.gl[1] = o.gl[2] = 0.0; o.sq = 0.0;
            o.gt = 0; o.gq = 0; o.dp = 0; o.ro = 0; o.ao = 0; o.qr = 0; o.qa = 0;
            o.rs = 0; o.as_ = 0; o.asc = 0; o.rp = 0; o.ap = 0;
            if (meta & SITE_SKIP) { o.gt = SVGT_GT_SKIPPED; o.gq = -1; }
            else if (!ranged) { o.gt = SVGT_GT_BLANK; o.gq = -1; err = SVGT_ERR_RANGE; }
            else call_site(p, t, meta & 3, r_ref_seq, r_alt_seq, r_alt_clip, r_ref_span, r_alt_span, o, err);
            int4 *dst = reinterpret_cast<int4 *>(p.out + site);
            const int4 *src = reinterpret_cast<const int4 *>(&o);
#pragma unroll
            for (int i = 0; i < 5; ++i) dst[i] = src[i];
        }
    }
    if (err) {
        atomicCAS(p.status, 0, err);
        atomicAdd(p.status + 2, 1);
    }
}

size_t coop_smem_bytes(const SvgtParams &p)
{
    size_t off = 256 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    off = (off + 127) & ~(size_t)127;
    off += sizeof(WarpSmem) * kCoopWarps;
    if (p.hist_in_smem) off += (size_t)p.n_hist * sizeof(unsigned);
    return off;
}

}  // namespace

int svgt_launch_coop(const SvgtParams &p, cudaStream_t stream)
{
    auto kern = (p.assoc_mode == SVGT_ASSOC_CLASSIC) ? svgt_coop_kernel<SVGT_ASSOC_CLASSIC>
                                                     : svgt_coop_kernel<SVGT_ASSOC_SSO>;
    const size_t smem = coop_smem_bytes(p);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SVGT_COOP_THREADS, smem)) != cudaSuccess)
        return (int)e;
    if (per_sm < 1) per_sm = 1;
    const long long want = ((long long)p.n_tiles + kCoopWarps - 1) / kCoopWarps;
    const long long cap = (long long)sms * per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, SVGT_COOP_THREADS, smem, stream>>>(p);
    return (int)cudaGetLastError();
}
