/*
 * svgt_lean.cu -- the default tally kernel (variant 5): the warp-cooperative mapping of svgt_coop.cu
 * (phase A one evidence row per lane, phase B one ordered fp64 chain per lane; reference
 * singlesample.py:355-404) with the lean row scorer of svgt_lean.cuh.
 *
 * What differs from svgt_tally_kernel (svgt_coop.cu), all of it aimed at the limits ncu showed there --
 * instruction issue and shared-memory wavefronts, not HBM:
 *   - rows travel HBM -> shared memory through a per-warp cp.async ring (each lane copies and later reads
 *     its own row: no registers or scoreboards held by rows in flight), one chunk stream per unit:
 *     fragment chunks step-major over the live sites, then split chunks;
 *   - per site the kernel pre-digests SiteF (two uniform LDS.128 per chunk; three for the general chain),
 *     SplitF, the row streams, and per (site, library) WinF (two LDS.128 per row instead of four).  Sites
 *     with both is_ref_seq windows valid take fast_row (one contig, not an inversion) or fast_row_gen
 *     (two contigs and / or inversion); a breakend within min_aligned of the contig start builds the
 *     cooperative kernel's windows on demand and takes its scorer, so every batch is still covered;
 *   - the insert-size histograms of the first four libraries live in shared memory with a zero
 *     sentinel behind each, so a look-up is an index clamp; the literal fp64 path reads the global copy;
 *   - a row parks 24 bytes {a + b, p_ref, p_alt} structure-of-arrays (three conflict-free STS.64); the
 *     two LUT indices replace a + b only where phase B needs a and b separately;
 *   - one CTA of 16 warps per SM (the per-CTA tables are paid once), work units claimed one ahead, a
 *     1-2-4-8 ramp of unit sizes for batches too small to amortise an 8-site unit of the heaviest sites.
 * The genotype call is the same second launch (svgt_call_kernel, svgt_coop.cu); running it inside this
 * kernel is a build switch that measured slower (SVGT_LEAN_FUSE_CALL).  Other measured switches:
 * SVGT_LEAN_TMA (cp.async.bulk per chunk), SVGT_LEAN_DEPTH, SVGT_LEAN_CP, SVGT_LEAN_BIG_RAMP
 * (profiles/README.md has the numbers).
 */
#include "svgt_lean.cuh"

namespace {

#ifndef SVGT_LEAN_THREADS
#define SVGT_LEAN_THREADS 512       /* one CTA of 16 warps per SM: the per-CTA tables are paid once */
#endif
#ifndef SVGT_LEAN_MINB
#define SVGT_LEAN_MINB 1
#endif
#ifndef SVGT_LEAN_DEPTH
#define SVGT_LEAN_DEPTH 3           /* chunks in flight per warp (1 KB each) */
#endif
constexpr int kD = SVGT_LEAN_DEPTH;
#ifndef SVGT_LEAN_FUSE_CALL
#define SVGT_LEAN_FUSE_CALL 0       /* 1: the genotype call runs at the end of each work unit (one launch): measured +26 % kernel time at 1M
                                       sites (log_choose is a latency-bound loop; in svgt_call_kernel a million threads hide it), -16 % at 10k */
#endif
#ifndef SVGT_LEAN_BIG_RAMP
#define SVGT_LEAN_BIG_RAMP 0        /* unit ramp for batches above ~600k sites: 0 none, 1 full, 2 single-site head only */
#endif
#ifndef SVGT_LEAN_CP
#define SVGT_LEAN_CP "cp.async.cg.shared.global"        /* A/B: .ca (through L1), .L2::128B / .L2::256B prefetch hints */
#endif
#ifndef SVGT_LEAN_TMA
#define SVGT_LEAN_TMA 0             /* 1: one cp.async.bulk (TMA 1-D) per chunk on an mbarrier instead of per-lane cp.async */
#endif
constexpr int kLeanWarps = SVGT_LEAN_THREADS / 32;
constexpr int kHistPad = 8;         /* sentinels behind the cached libraries' counts */
constexpr int kLeanHistWords = 4864;/* shared-memory budget for the cached counts (the CTA uses ~211 KB besides) */

/* one row stream of a site: where its rows start and how many there are (one uniform LDS.128 per chunk) */
struct alignas(16) Strm { const int4 *rows; int n; int pad; };

template <int G>
struct alignas(128) LeanSmem {
    unsigned char ring[kD][1024];   /* cp.async targets: 32 x 16 B low halves, then 32 x 16 B high halves */
    unsigned long long bar[kD];     /* SVGT_LEAN_TMA: one mbarrier per ring slot */
    Strm strm[2][8];                /* row streams: [0] fragment rows, [1] split rows (a chunk's low 4 bits index it) */
    SiteS site[G];
    SiteF sf[G];
    SplitF spf[G];
    WinF wf[G][kWLibs + 1];
    Win gwin[kWLibs];               /* windows of the non-fast site being scored */
    Parked park[G];                 /* phase A -> phase B */
    double zero[2];
};

/*
 * Work units.  The launch order is work-descending, so fixed 8-site units put the eight heaviest sites
 * on one warp (the critical path of small or heavy-tailed batches).  With the ramp, the first W units
 * (W = resident warps) hold one site each, the next W two, the next W four, the rest G = 8: the head of
 * the order is spread one site per warp, the bulk keeps full 8-site interleaving for the ordered replay.
 */
struct UnitRange { long long base; int count; };

/* ramp: 0 none, 1 full (1-2-4-8), 2 short (W single-site units, then 8): large batches */
template <int G>
__host__ __device__ __forceinline__ UnitRange lean_unit_range(long long unit, long long W, int ramp)
{
    UnitRange r;
    if (!ramp || G < 8) { r.base = unit * G; r.count = G; return r; }
    if (unit < W) { r.base = unit; r.count = 1; }
    else if (ramp == 2) { r.base = W + (unit - W) * 8; r.count = 8; }
    else if (unit < 2 * W) { r.base = W + (unit - W) * 2; r.count = 2; }
    else if (unit < 3 * W) { r.base = 3 * W + (unit - 2 * W) * 4; r.count = 4; }
    else { r.base = 7 * W + (unit - 3 * W) * 8; r.count = 8; }
    return r;
}

template <int G>
__host__ __device__ __forceinline__ long long lean_n_units(long long n_sites, long long W, int ramp)
{
    if (!ramp || G < 8) return (n_sites + G - 1) / G;
    if (n_sites <= W) return n_sites;
    if (ramp == 2) return W + (n_sites - W + 7) / 8;
    if (n_sites <= 3 * W) return W + (n_sites - W + 1) / 2;
    if (n_sites <= 7 * W) return 2 * W + (n_sites - 3 * W + 3) / 4;
    return 3 * W + (n_sites - 7 * W + 7) / 8;
}

__device__ __forceinline__ unsigned lean_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__host__ __device__ __forceinline__ long long lean_hist_words(long long n_hist)
{
    return (n_hist < kLeanHistWords ? n_hist : kLeanHistWords) + kHistPad;
}

#if SVGT_LEAN_FUSE_CALL
/* the genotype call of one site on its five sums (what svgt_call_kernel does per thread), out of line */
__device__ __noinline__ int finish_site(const SvgtParams &p, int site, int status, int svtype, double ref_seq,
                                        double alt_seq, double alt_clip, double ref_span, double alt_span)
{
    svgt_out_row_t o;
    o.gl[0] = o.gl[1] = o.gl[2] = 0.0; o.sq = 0.0;
    o.gt = 0; o.gq = 0; o.dp = 0; o.ro = 0; o.ao = 0; o.qr = 0; o.qa = 0;
    o.rs = 0; o.as_ = 0; o.asc = 0; o.rp = 0; o.ap = 0;
    int e = 0;
    if (status == 1) { o.gt = SVGT_GT_SKIPPED; o.gq = -1; }
    else if (status == 2) { o.gt = SVGT_GT_BLANK; o.gq = -1; e = SVGT_ERR_RANGE; }
    else {
        Tables t;
        t.pm = p.pm; t.libs = nullptr; t.hist = p.hist; t.conc = 0.0; t.disc = 0.0;
        call_site(p, t, svtype, ref_seq, alt_seq, alt_clip, ref_span, alt_span, o, e);
    }
    int4 *dst = reinterpret_cast<int4 *>(p.out + site);
    const int4 *src = reinterpret_cast<const int4 *>(&o);
#pragma unroll
    for (int i = 0; i < 5; ++i) dst[i] = src[i];
    return e;
}
#endif

/* non-fast sites: the cooperative kernel's scorer, out of line so the hot loop stays small; everything
 * the hot loop keeps in registers travels by value */
struct GenericOut { FragOut fo; unsigned carryA, carryB; int err; };

#ifndef SVGT_LEAN_GENERIC_INLINE
#define SVGT_LEAN_GENERIC_INLINE 0
#endif
template <int ASSOC>
#if SVGT_LEAN_GENERIC_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
GenericOut generic_frag_chunk(const SvgtParams &p, const Tables &t, const SiteS &S, const Win *wins, const double *s_pm,
                              const LibK *s_lib, const int lane, const int n, const int g, const int m, const int4 lo,
                              const int4 hi, unsigned carryA, unsigned carryB, int err)
{
    GenericOut r;
    r.fo = score_frag_chunk<ASSOC>(p, t, S, wins, s_pm, s_lib, p.hist, lane, n, g, m, lo, hi, carryA, carryB, err);
    r.carryA = carryA; r.carryB = carryB; r.err = err;
    return r;
}

template <int G, int ASSOC>
__global__ void __launch_bounds__(SVGT_LEAN_THREADS, SVGT_LEAN_MINB) svgt_lean_kernel(const SvgtParams p)
{
    typedef LeanSmem<G> WS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_pm = reinterpret_cast<double *>(smem_raw);
    LibK *s_lib = reinterpret_cast<LibK *>(s_pm + 512);     /* pm[0..255], then pm[q] / 2 */
    size_t off = 512 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    LibF *s_libf = reinterpret_cast<LibF *>(smem_raw + off);
    off += (kWLibs + 1) * sizeof(LibF);
    off = (off + 127) & ~(size_t)127;
    WS *s_warp = reinterpret_cast<WS *>(smem_raw + off);
    off += sizeof(WS) * kLeanWarps;
    unsigned *s_hist = reinterpret_cast<unsigned *>(smem_raw + off);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    int err = 0;
    const int nl = p.n_lib < SVGT_SMEM_LIBS ? p.n_lib : SVGT_SMEM_LIBS;
    for (int i = tid; i < 256; i += SVGT_LEAN_THREADS) {
        const double v = p.pm[i];
        s_pm[i] = v;
        s_pm[256 + i] = __dmul_rn(v, 0.5);
    }
    for (int i = tid; i < nl; i += SVGT_LEAN_THREADS) s_lib[i] = derive_lib(p, i, &err);
    /* the 32-bit form of 19*h1 > h2 needs every histogram count below 2^26 */
    int big = 0;
    for (long long i = tid; i < p.n_hist; i += SVGT_LEAN_THREADS) big |= p.hist[i] >= (1u << 26);
    WS &ws = s_warp[warp];
    if (lane < 2) ws.zero[lane] = 0.0;
#if SVGT_LEAN_TMA
    if (lane < kD) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lean_smem_addr(&ws.bar[lane])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned cc = 0u;                                       /* chunks consumed by this warp so far (slot phase) */
#endif
    const bool small_counts = __syncthreads_or(big) == 0;
    /* lean copies of the first kWLibs histograms, each followed by a zero sentinel */
    if (tid == 0) {
        const long long cap = lean_hist_words(p.n_hist);
        long long base = 0;
        for (int l = 0; l <= kWLibs; ++l) {
            LibF f; f.addr = 0u; f.len = 0; f.ok = 0; f.pad = 0;
            if (l < kWLibs && l < nl) {
                const LibK &L = s_lib[l];
                const bool fits = L.hist_len < (1 << kHistLenBits) && base + L.hist_len + 1 <= cap;
                if (L.safe && small_counts && fits) {
                    f.addr = lean_smem_addr(s_hist + base); f.len = L.hist_len; f.ok = 1;
                    base += L.hist_len + 1;
                }
            }
            s_libf[l] = f;
        }
    }
    __syncthreads();
    for (int l = 0; l < kWLibs; ++l) {
        const LibF f = s_libf[l];
        if (!f.ok) continue;
        unsigned *dst = s_hist + ((f.addr - lean_smem_addr(s_hist)) >> 2);
        const unsigned *src = p.hist + s_lib[l].hist_off;
        for (int i = tid; i <= f.len; i += SVGT_LEAN_THREADS) dst[i] = i < f.len ? src[i] : 0u;
    }
    __syncthreads();

    Tables t;
    t.pm = s_pm; t.libs = s_lib; t.hist = p.hist;           /* literal paths read the global counts */
    t.conc = p.consts[C_CONC]; t.disc = p.consts[C_DISC];
    const int m = p.min_aligned, slop = p.split_slop;
    const unsigned zero_addr = lean_smem_addr(&ws.zero[0]);

    /* phase-B role of this lane: chain c of interleaved site gb */
    const int gb = lane >> 2, c = lane & 3;
    const int ramp = p.n_tiles;                             /* SvgtParams::n_tiles doubles as the ramp mode here */
    const long long W = (long long)gridDim.x * kLeanWarps;
    const long long n_units = lean_n_units<G>(p.n_sites, W, ramp);

    unsigned head = 0u;                                     /* ring slot of the chunk being scored; never reset, so
                                                               slot s is always chunk number == s (mod kD) of this warp */
    long long unit = 0, unit_next = 0;
    if (lane == 0) unit = (long long)atomicAdd(reinterpret_cast<unsigned *>(p.status + 1), 1u);
    unit = __shfl_sync(full, unit, 0);
    for (; unit < n_units; unit = __shfl_sync(full, unit_next, 0)) {

        /* ---- lanes 0..G-1 read their site row and publish the scalars ---- */
        {
            const UnitRange ur = lean_unit_range<G>(unit, W, ramp);
            const long long idx = ur.base + lane;
            const bool valid = lane < ur.count && idx < p.n_sites;
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
                S.slot = valid ? (int)site : -1;        /* where the sums go (order[] is int32) */
                S.pad1 = (meta & SITE_SKIP) ? 1 : (!ranged ? 2 : 0);     /* fused call: skipped / out of range */
                SiteF &F = ws.sf[lane];
                const int svtype = meta & 3;
                F.tA = b.z; F.wA0 = a.x - m; F.wA1 = a.x + m; F.wB0 = a.y - m; F.wB1 = a.y + m;
                F.pat = (meta & (SITE_O1_REV | SITE_O2_REV)) | F_PAIRED;   /* site bits 2,3 line up with F_REV_A/B */
                F.del = svtype == SV_DEL;
                const bool okwin = (a.x - m >= 0) && (a.y - m >= 0);      /* both is_ref_seq windows valid */
                F.fast = !okwin ? 0 : (svtype != SV_INV && b.z == b.w) ? 1 : 2;
                F.tB = b.w; F.inv = svtype == SV_INV;
                /* make_win(): the reciprocal window starts FL after the alt window for a forward breakend
                 * (alt [L - FL, H] vs reciprocal [L, H + FL]), FL before it for a reverse one; same width */
                F.sgnA = (meta & SITE_O1_REV) ? 1 : -1; F.sgnB = (meta & SITE_O2_REV) ? 1 : -1;
                ws.spf[lane] = make_splitf(S, slop);
                Strm q;
                q.pad = 0;
                q.rows = p.frags + 2 * foff; q.n = nf; ws.strm[0][lane] = q;
                q.rows = p.splits + 2 * soff; q.n = ns; ws.strm[1][lane] = q;
            }
            __syncwarp();
            for (int i = lane; i < G * (kWLibs + 1); i += 32) {
                const int g = i / (kWLibs + 1), l = i % (kWLibs + 1);
                if (ws.site[g].nf && ws.sf[g].fast)
                    ws.wf[g][l] = make_winf(ws.site[g], s_lib[l < nl ? l : 0], s_libf[l < nl ? l : kWLibs], m, zero_addr);
            }
            __syncwarp();
        }

        /* claim the next unit now: the atomic's round trip is hidden behind this unit's rows */
        if (lane == 0) unit_next = (long long)atomicAdd(reinterpret_cast<unsigned *>(p.status + 1), 1u);

        double sum_frag = 0.0;      /* lane 4g+c: chain c of the fragment rows of site g */
        double sum_split = 0.0;     /* lane 4g+c: chain c of the split rows of site g    */
        {
            /*
             * One chunk stream per unit: fragment chunks step-major over the sites that still have rows,
             * then split chunks the same way (both are 32-byte rows).  A chunk is described by one int,
             * phase << 27 | step << 3 | g.  Rows travel HBM -> shared memory by cp.async (LDGSTS, 16 bytes
             * per lane twice, zero-filled beyond the site's last row), kD chunks ahead of the one being
             * scored; a lane reads back exactly the row it copied, so completion is the lane's own
             * cp.async.wait_group and no registers or scoreboards are held by rows in flight.
             */
            double acc = 0.0, pend = 0.0;
            unsigned carryA = 0u, carryB = 0u;      /* EXTRA-run hits carried into the next step, bit g */
            unsigned long long leads = 0ull;        /* `lead` of each site's chunk in this super-step, 8 bits per site */
            const int my_nf = lane < G ? ws.site[lane].nf : 0;
            const int my_ns = lane < G ? ws.site[lane].ns : 0;
            /* a chunk is one int: step << 4 | phase << 3 | g  (its low four bits index ws.strm) */
            int it_phase = 0, it_step = -1;
            unsigned it_mask = 0u;
            auto next_chunk = [&]() -> int {
                while (it_mask == 0u) {
                    ++it_step;
                    it_mask = __ballot_sync(full, (it_phase ? my_ns : my_nf) > it_step * 32);
                    if (it_mask == 0u) {
                        if (it_phase) return -1;
                        it_phase = 8; it_step = -1;
                    }
                }
                const int g = __ffs(it_mask) - 1;
                it_mask &= it_mask - 1u;
                return (it_step << 4) | it_phase | g;
            };
            const Strm *strm = &ws.strm[0][0];
#if SVGT_LEAN_TMA
            const unsigned ring = lean_smem_addr(&ws.ring[0][0]) + lane * 32;
            const unsigned bars = lean_smem_addr(&ws.bar[0]);
            auto issue = [&](const int d, const unsigned slot) {
                if (d >= 0) {
                    const Strm q = strm[d & 15];
                    const int row0 = (d >> 4) * 32;
                    const int n = min(32, q.n - row0);
                    const unsigned dst = ring + slot * 1024u;
                    if (lane == 0) {
                        const unsigned bytes = (unsigned)n * 32u;
                        const unsigned bar = bars + slot * 8u;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(dst), "l"(q.rows + 2 * (long long)row0), "r"(bytes), "r"(bar) : "memory");
                    }
                    if (lane >= n) {                        /* rows beyond the last one read as zeros */
                        asm volatile("st.shared.v4.s32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
                        asm volatile("st.shared.v4.s32 [%0], {%1, %1, %1, %1};" ::"r"(dst + 16u), "r"(0) : "memory");
                    }
                }
            };
#else
            const unsigned ring = lean_smem_addr(&ws.ring[0][0]) + lane * 16;
            auto issue = [&](const int d, const unsigned slot) {
                if (d >= 0) {
                    const Strm q = strm[d & 15];
                    const int row = (d >> 4) * 32 + lane;
                    /* rows beyond the last one are zero-filled (0 source bytes); the address stays in range */
                    const int4 *src = q.rows + 2 * (long long)min(row, q.n - 1);
                    const unsigned bytes = row < q.n ? 16u : 0u;
                    const unsigned dst = ring + slot * 1024u;
                    asm volatile(SVGT_LEAN_CP " [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
                    asm volatile(SVGT_LEAN_CP " [%0], [%1], 16, %2;" ::"r"(dst + 512u), "l"(src + 1), "r"(bytes) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
#endif
            /* descriptors of the chunk being scored (dq[0]) and of the kD chunks in flight behind it */
            int dq[kD + 1];
#pragma unroll
            for (int i = 0; i < kD; ++i) {
                dq[i] = next_chunk();
                issue(dq[i], (head + (unsigned)i) % (unsigned)kD);
            }
            while (dq[0] >= 0) {
                const int d = dq[0];
                const unsigned src = ring + head * 1024u;
                int4 lo, hi;
#if SVGT_LEAN_TMA
                {
                    const unsigned bar = bars + head * 8u, parity = (cc / (unsigned)kD) & 1u;
                    unsigned ok;
                    do {
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                    } while (!ok);
                    ++cc;
                }
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(src));
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                             : "r"(src + 16u));
#else
                asm volatile("cp.async.wait_group %0;" ::"n"(kD - 1) : "memory");
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(src));
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                             : "r"(src + 512u));
#endif
                /* refill the slot just read with the chunk kD ahead */
                dq[kD] = next_chunk();
                issue(dq[kD], head);
                head = head + 1u == (unsigned)kD ? 0u : head + 1u;

                const int g = d & 7, step = d >> 4;
                const bool sp = (d & 8) != 0;
                if (!sp) {
                    /* ---- phase A, fragment rows ---- */
                    FragOut fo;
                    if (ws.sf[g].fast) {
                        fo = score_frag_chunk_fast<ASSOC>(p, t, ws.site[g], ws.sf[g], &ws.wf[g][0], s_pm, s_lib, lane, step, g,
                                                          m, lo, hi, carryA, carryB, err);
                    } else {
                        __syncwarp();
                        if (lane < kWLibs) {
                            if (lane < nl) ws.gwin[lane] = make_win(ws.site[g], s_lib[lane], m, small_counts);
                            else ws.gwin[lane].flags = 0u;
                        }
                        __syncwarp();
                        const GenericOut r = generic_frag_chunk<ASSOC>(p, t, ws.site[g], &ws.gwin[0], s_pm, s_lib, lane,
                                                                       ws.site[g].nf - step * 32, g, m, lo, hi, carryA, carryB,
                                                                       err);
                        fo = r.fo; carryA = r.carryA; carryB = r.carryB; err = r.err;
                    }
                    park_frag_soa<ASSOC>(ws.park[g], lane, fo);
                    if (fo.lead) leads |= (unsigned long long)fo.lead << (8 * g);
                } else {
                    /* ---- phase A, split rows ---- */
                    const SplitOut so = score_split_chunk_lean<ASSOC>(ws.spf[g], s_pm, lane, ws.strm[1][g].n - step * 32, lo, hi);
                    ws.park[g].ch[0][lane] = so.vseq; ws.park[g].ch[1][lane] = so.vclip;
                    if (so.lead) leads |= (unsigned long long)so.lead << (8 * g);
                }
                /* ---- phase B after the last chunk of a super-step (same phase and step) ---- */
                if ((dq[1] >> 3) != (d >> 3)) {
                    __syncwarp();
                    if (gb < G && c < (sp ? 2 : 3)) {
                        int cnt = ws.strm[sp ? 1 : 0][gb].n - step * 32;
                        cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
                        const int lead = (int)(leads >> (8 * gb)) & 0xFF;
                        if (!sp) replay_frag_soa<ASSOC>(ws.park[gb], c, cnt, lead, s_pm, acc, pend);
                        else replay_split_soa<ASSOC>(ws.park[gb], c, cnt, lead, acc, pend);
                    }
                    leads = 0ull;
                    __syncwarp();
                    if (!sp && (dq[1] < 0 || (dq[1] & 8) != 0)) {       /* the fragment rows are done */
                        if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
                        sum_frag = acc; acc = 0.0; pend = 0.0;
                    }
                }
#pragma unroll
                for (int i = 0; i < kD; ++i) dq[i] = dq[i + 1];
            }
            if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
            sum_split = acc;
        }

#if SVGT_LEAN_FUSE_CALL
        /* ---- zeroing rules + bayesian_genotype (singlesample.py:382-473) for the unit's sites: lane 4g gathers
         *      the five sums of site g and writes the final 80-byte row ---- */
        {
            const double ref_span = __shfl_down_sync(full, sum_frag, 1), alt_span = __shfl_down_sync(full, sum_frag, 2);
            const double alt_clip = __shfl_down_sync(full, sum_split, 1);
            if (gb < G && c == 0) {
                const int site = ws.site[gb].slot;
                if (site >= 0) {
                    const int e = finish_site(p, site, ws.site[gb].pad1, ws.site[gb].meta & 3, sum_frag, sum_split, alt_clip,
                                              ref_span, alt_span);
                    if (e) err = e;
                }
            }
        }
#else
        /* ---- park the five sums in the site's output row (lane 4g+c holds chain c of site g) ---- */
        if (gb < G && c < 3) {
            const int site = ws.site[gb].slot;
            if (site >= 0 && (ws.site[gb].nf | ws.site[gb].ns)) {
                double *row = reinterpret_cast<double *>(p.out + site);
                /* ParkedSums: ref_seq, alt_seq, alt_clip, ref_span, alt_span */
                if (c == 0) { row[0] = sum_frag; row[1] = sum_split; }
                else if (c == 1) { row[3] = sum_frag; row[2] = sum_split; }
                else row[4] = sum_frag;
            }
        }
#endif
        __syncwarp();
    }
    if (err) {
        atomicCAS(p.status, 0, err);
        atomicAdd(p.status + 2, 1);
    }
}

template <int G>
size_t lean_smem_bytes(const SvgtParams &p)
{
    size_t off = 512 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK) + (kWLibs + 1) * sizeof(LibF);
    off = (off + 127) & ~(size_t)127;
    off += sizeof(LeanSmem<G>) * kLeanWarps;
    off += (size_t)lean_hist_words(p.n_hist) * sizeof(unsigned);
    return off;
}

/* per (kernel instantiation, device): opt-in shared memory and the resident-CTA count, queried once */
struct LeanLaunchInfo { int ready[16]; int per_sm[16]; int sms[16]; size_t smem_set[16]; };

template <int G>
int launch_lean(const SvgtParams &p, int ramp, cudaStream_t stream)
{
    static LeanLaunchInfo info[2] = {};
    const int a = p.assoc_mode == SVGT_ASSOC_CLASSIC ? 1 : 0;
    auto kern = a ? svgt_lean_kernel<G, SVGT_ASSOC_CLASSIC> : svgt_lean_kernel<G, SVGT_ASSOC_SSO>;
    const size_t smem = lean_smem_bytes<G>(p);
    cudaError_t e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    LeanLaunchInfo &li = info[a];
    const int di = dev & 15;
    if (!li.ready[di] || li.smem_set[di] < smem) {
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return (int)e;
        if ((e = cudaDeviceGetAttribute(&li.sms[di], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&li.per_sm[di], kern, SVGT_LEAN_THREADS, smem)) !=
            cudaSuccess)
            return (int)e;
        if (li.per_sm[di] < 1) li.per_sm[di] = 1;
        li.smem_set[di] = smem; li.ready[di] = 1;
    }
    const long long cap = (long long)li.sms[di] * li.per_sm[di];      /* persistent: one resident wave */
    /* with the ramp a unit is one site at first, so the grid is sized by sites */
    const long long units = ramp ? p.n_sites : (p.n_sites + G - 1) / G;
    const long long want = (units + kLeanWarps - 1) / kLeanWarps;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    /* the ramp trades a little replay efficiency on the heaviest ~7 W sites for a short critical path:
     * worth it until the batch is large enough to amortise its longest unit (measured: +2 % kernel time
     * at 1M sites, -30 % at 200k heavy-tailed sites) */
    if (ramp == 1 && p.n_sites >= 256 * cap * kLeanWarps) ramp = SVGT_LEAN_BIG_RAMP;    /* ~600k sites on a B200 */
    SvgtParams q = p;
    q.n_tiles = ramp;
    kern<<<grid, SVGT_LEAN_THREADS, smem, stream>>>(q);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
#if SVGT_LEAN_FUSE_CALL
    return 0;
#else
    return svgt_launch_call(p, stream);
#endif
}

}  // namespace

int svgt_lean_launches(void) { return SVGT_LEAN_FUSE_CALL ? 1 : 2; }

#ifndef SVGT_LEAN_G
#define SVGT_LEAN_G 8
#endif
int svgt_launch_lean(const SvgtParams &p, int variant, cudaStream_t stream)
{
    if (variant == SVGT_VAR_LEAN2) return launch_lean<2>(p, 0, stream);
    return launch_lean<SVGT_LEAN_G>(p, variant != SVGT_VAR_LEAN8 ? 1 : 0, stream);
}
