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
#include "svgt_coop.cuh"

namespace {


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
                const int n = ws.site[g].nf - step * 32;
                const FragOut fo = score_frag_chunk<ASSOC>(p, t, ws.site[g], &ws.win[g][0], s_pm, s_lib, hist, lane, n, g, m,
                                                           lo, hi, carryA, carryB, err);
                park_frag(&ws.contrib[g][lane][0], fo);
                if (lane == 0) ws.newmask[g] = (unsigned)fo.lead;
#ifdef SVGT_MARK
                asm volatile("membar.cta;" ::: "memory");
#endif

                /* ---------------- phase B at the end of each super-step ---------------- */
                if (flush) {
                    __syncwarp();
                    if (gb < G && c < 3) {
                        int cnt = ws.site[gb].nf - step * 32;
                        cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
#if !(SVGT_DIAG & 2)
                        replay_frag<ASSOC>(&ws.contrib[gb][0][0], c, cnt, (int)ws.newmask[gb], s_pm, acc, pend);
#else
                        acc += ws.contrib[gb][0][c] + cnt;
#endif
                    }
                    __syncwarp();
                }
            };
#if SVGT_ROWBUFS == 1
            /* no register prefetch: the other warps of the SM cover the load */
            int cs0 = 0, cg0 = 0;
            bool k0 = advance(cs0, cg0);
            while (k0) {
                int4 l0, h0;
                load_rows(cs0, cg0, l0, h0);
                int cs1 = 0, cg1 = 0;
                const bool k1 = advance(cs1, cg1);
                process(cs0, cg0, l0, h0, !k1 || cs1 != cs0);
                cs0 = cs1; cg0 = cg1; k0 = k1;
            }
#elif SVGT_ROWBUFS == 4
            /* rotating register queue, three chunks in flight */
            int cs0 = 0, cg0 = 0, cs1 = 0, cg1 = 0, cs2 = 0, cg2 = 0, cs3 = 0, cg3 = 0;
            int4 l0, h0, l1, h1, l2, h2, l3, h3;
            bool k0 = advance(cs0, cg0);
            if (k0) load_rows(cs0, cg0, l0, h0);
            bool k1 = k0 && advance(cs1, cg1);
            if (k1) load_rows(cs1, cg1, l1, h1);
            bool k2 = k1 && advance(cs2, cg2);
            if (k2) load_rows(cs2, cg2, l2, h2);
            while (k0) {
                const bool k3 = k2 && advance(cs3, cg3);
                if (k3) load_rows(cs3, cg3, l3, h3);
                process(cs0, cg0, l0, h0, !k1 || cs1 != cs0);
                cs0 = cs1; cg0 = cg1; l0 = l1; h0 = h1; k0 = k1;
                cs1 = cs2; cg1 = cg2; l1 = l2; h1 = h2; k1 = k2;
                cs2 = cs3; cg2 = cg3; l2 = l3; h2 = h3; k2 = k3;
            }
#elif SVGT_ROWBUFS == 3
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
#if SVGT_SPLIT_ROT
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
                    const SplitOut so = score_split_chunk<ASSOC>(S, s_pm, lane, n, slop, q0, q1);
                    *reinterpret_cast<double2 *>(&ws.contrib[g][lane][0]) = make_double2(so.vseq, so.vclip);
                    if (lane == 0) ws.newmask[g] = (unsigned)so.lead;
                }
                if (!e1 || ss1 != ss0) {                        /* last chunk of its super-step: phase B */
                    __syncwarp();
                    if (gb < G && c < 2) {
                        int cnt = ws.site[gb].ns - ss0 * 32;
                        cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
                        replay_split<ASSOC>(&ws.contrib[gb][0][0], c, cnt, (int)ws.newmask[gb], acc, pend);
                    }
                    __syncwarp();
                }
                ss0 = ss1; sg0 = sg1; a0 = a1; b0 = b1; e0 = e1;
                ss1 = ss2; sg1 = sg2; a1 = a2; b1 = b2; e1 = e2;
            }
#else
            int ss0 = 0, sg0 = 0;
            bool e0 = advance_split(ss0, sg0);
            while (e0) {
                int4 q0, q1;
                load_split(ss0, sg0, q0, q1);
                const SiteS &S = ws.site[sg0];
                const int n = min(32, S.ns - ss0 * 32);
                const SplitOut so = score_split_chunk<ASSOC>(S, s_pm, lane, n, slop, q0, q1);
                *reinterpret_cast<double2 *>(&ws.contrib[sg0][lane][0]) = make_double2(so.vseq, so.vclip);
                if (lane == 0) ws.newmask[sg0] = (unsigned)so.lead;
                int ss1 = 0, sg1 = 0;
                const bool e1 = advance_split(ss1, sg1);
                if (!e1 || ss1 != ss0) {                        /* last chunk of its super-step: phase B */
                    __syncwarp();
                    if (gb < G && c < 2) {
                        int cnt = ws.site[gb].ns - ss0 * 32;
                        cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
                        replay_split<ASSOC>(&ws.contrib[gb][0][0], c, cnt, (int)ws.newmask[gb], acc, pend);
                    }
                    __syncwarp();
                }
                ss0 = ss1; sg0 = sg1; e0 = e1;
            }
#endif
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
    return svgt_launch_call(p, stream);
}

}  // namespace

int svgt_launch_call(const SvgtParams &p, cudaStream_t stream)
{
    const int cgrid = (int)((p.n_sites + 255) / 256);
    svgt_call_kernel<<<cgrid, 256, 0, stream>>>(p);
    return (int)cudaGetLastError();
}

int svgt_launch_coop(const SvgtParams &p, int variant, cudaStream_t stream)
{
    if (variant == SVGT_VAR_COOP4) return launch_coop<4>(p, stream);
    return launch_coop<8>(p, stream);
}
