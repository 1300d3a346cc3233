/*
 * svgt_compact.cu -- the default tally kernel: compact 16-byte evidence rows (svtyper_b200/compact.py)
 * delivered by TMA bulk copies, scored one row per lane, summed in order one chain per lane
 * (reference singlesample.py:355-404: tally_variant_read_fragments; classic.py:286-435).
 *
 * Mapping (one persistent CTA per SM, work units claimed from an atomic cursor):
 *   - a warp owns a unit of up to G = 6 sites of the work-descending launch order and walks them in
 *     lock step: super-step k covers rows [32k, 32k + 32) of every site that still has rows -- first the
 *     fragment rows of the sites, then their split rows;
 *   - row delivery: at the start of super-step k lanes 0..G-1 each issue ONE cp.async.bulk (TMA 1-D,
 *     <= 512 bytes, global -> shared) for THEIR site's rows of super-step k + 1 into the other half of a
 *     two-slot ring, all G copies signalling one mbarrier whose transaction count lane 0 armed with the
 *     summed byte count.  The warp polls that barrier once per super-step and each lane then reads its row
 *     with one conflict-free LDS.128.  No per-lane copy instructions, no registers or scoreboards held by
 *     rows in flight, a whole super-step (2-3 us of scoring) of prefetch distance;
 *   - phase A (score_cfrag_chunk / score_csplit_chunk): one row per lane, straight-line predicate chain;
 *     the three doubles a row yields {a + b, p_ref, p_alt} are parked structure-of-arrays -- p_ref / p_alt
 *     over the site's 512 bytes of the ring slot the rows were just read from, a + b beside it;
 *   - phase B after each super-step: lane 4g + c replays chain c of site g over the <= 32 parked rows in
 *     row order (the fp64 sums are order-sensitive: SURVEY.md H1), 3 G chains side by side;
 *   - the five sums of a site are parked in its 80-byte row of `out`; svgt_call_compact_kernel (one site
 *     per thread) applies the zeroing rules and bayesian_genotype (singlesample.py:382-473) and writes the
 *     final row -- to `out`, or to `out_final` when the caller supplies one (a peer-mapped buffer on
 *     another GPU: the multi-GPU gather without a collective).
 */
#include "svgt_compact.cuh"

#include <stdlib.h>

namespace {

#ifndef SVGT_C_G
#define SVGT_C_G 6                  /* sites per work unit.  Measured on the benchmark shape (1M sites): G = 6 with 20
                                       warps 1.55 ms, G = 8 with 15 warps 1.59 ms, G = 4 with 28 / 24 / 20 warps
                                       1.73 / 1.76 / 1.67 ms: fewer sites per unit buy resident warps (shared memory) but
                                       thin out the 3 G chains of the ordered replay and the per-super-step overheads */
#endif
#ifndef SVGT_C_THREADS
#define SVGT_C_THREADS 640          /* 20 warps x 10 KB of per-warp state + the per-CTA tables fill the 227 KB */
#endif
#ifndef SVGT_C_MINB
#define SVGT_C_MINB 1
#endif
#ifndef SVGT_C_RAMP_PER_WARP_DEFAULT
#define SVGT_C_RAMP_PER_WARP_DEFAULT 256
#endif
#ifndef SVGT_C_PREFETCH
#define SVGT_C_PREFETCH 0
#endif
#ifndef SVGT_C_UNROLL
#define SVGT_C_UNROLL 1
#endif
#ifndef SVGT_C_DEPTH
#define SVGT_C_DEPTH 2              /* super-step slots in the ring */
#endif
constexpr int kCD = SVGT_C_DEPTH;
constexpr int kCUnroll = SVGT_C_UNROLL;
constexpr int kCWarps = SVGT_C_THREADS / 32;
constexpr int kCHistPad = 8;

template <int G>
struct alignas(128) CWarpSmem {
    int4 ring[kCD][G][33];          /* TMA destinations: [slot][site][row].  Once the warp has read a site's rows, the
                                       same 528 bytes park p_ref[33] | p_alt[33] (or alt_seq | alt_clip) for phase B */
    double spark[G][33];            /* parked a + b (or the two LUT indices where phase B needs a and b apart) */
    unsigned long long bar[kCD];    /* one mbarrier per slot */
    int cnt[2][8];                  /* [0] fragment rows, [1] split rows of each site (uniform reads) */
    int lead[8];                    /* rows at the head of a site's chunk that continue the previous chunk's fragment */
    SiteS site[G];
    CSiteF sf[G];
    CSplitF spf[G];
    WinF wf[G][kWLibs + 1];
    double zero[2];
};

/* shared memory left for the cached histogram counts once the per-warp state and the per-CTA tables are placed */
constexpr size_t kCFixedBytes = 256 * sizeof(double) + ((SVGT_SMEM_LIBS * sizeof(LibK) + (kWLibs + 1) * sizeof(LibF) + 127) & ~(size_t)127) +
                                sizeof(CWarpSmem<SVGT_C_G>) * kCWarps;
constexpr size_t kCSmemLimit = 227 * 1024;
static_assert(kCFixedBytes + 4096 <= kCSmemLimit, "per-warp state does not fit the SM's shared memory");
constexpr long long kCHistWordsMax = (long long)((kCSmemLimit - kCFixedBytes) / 4) - kCHistPad;

__host__ __device__ __forceinline__ long long c_hist_words(long long n_hist)
{
    return (n_hist < kCHistWordsMax ? n_hist : kCHistWordsMax) + kCHistPad;
}

/* unit ramp: the launch order is work-descending, so for batches too small to amortise their longest unit the
 * first W units (W = resident warps) hold one site each, the next W two, the next W four, the rest G: the
 * heaviest sites are spread one per warp, the bulk keeps G-site interleaving for the ordered replay */
struct CUnit { long long base; int count; };

template <int G>
__host__ __device__ __forceinline__ CUnit c_unit_range(long long unit, long long W, int ramp)
{
    CUnit r;
    if (!ramp || G < 5) { r.base = unit * G; r.count = G; return r; }
    if (unit < W) { r.base = unit; r.count = 1; }
    else if (ramp == 2) { r.base = W + (unit - W) * G; r.count = G; }
    else if (unit < 2 * W) { r.base = W + (unit - W) * 2; r.count = 2; }
    else if (unit < 3 * W) { r.base = 3 * W + (unit - 2 * W) * 4; r.count = 4; }
    else { r.base = 7 * W + (unit - 3 * W) * G; r.count = G; }
    return r;
}

template <int G>
__host__ __device__ __forceinline__ long long c_n_units(long long n_sites, long long W, int ramp)
{
    if (!ramp || G < 5) return (n_sites + G - 1) / G;
    if (n_sites <= W) return n_sites;
    if (ramp == 2) return W + (n_sites - W + G - 1) / G;
    if (n_sites <= 3 * W) return W + (n_sites - W + 1) / 2;
    if (n_sites <= 7 * W) return 2 * W + (n_sites - 3 * W + 3) / 4;
    return 3 * W + (n_sites - 7 * W + G - 1) / G;
}

__device__ __forceinline__ unsigned c_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

/* phase B of one fragment chunk for chain c: `px` = the chain's 33 parked doubles (c = 0: a + b, or the LUT
 * index pairs `pi` where a and b must be added apart); the replay of replay_frag_soa() */
template <int ASSOC, int IMASK = -1>   /* IMASK: applied to the LUT indices (the replay kernel reads them from scratch memory) */
__device__ __forceinline__ void c_replay_frag(const double *px, const double *s, int c, int cnt, int lead, const double *s_pm,
                                              double &acc, double &pend)
{
    const int2 *pi = reinterpret_cast<const int2 *>(s);                 /* .x = ia, .y = ib */
    if (cnt <= 0) return;
    lead = lead < cnt ? lead : cnt;
    if (ASSOC == SVGT_ASSOC_CLASSIC) {
        if (c == 0) {
            for (int j = 0; j < cnt; ++j) {
                const int2 ix = pi[j];
                acc = __dadd_rn(__dadd_rn(acc, s_pm[ix.x & IMASK]), s_pm[ix.y & IMASK]);
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) acc = __dadd_rn(acc, px[j]);
        }
    } else {
        for (int j = 0; j < lead; ++j) {                /* rows continuing the previous chunk's last fragment */
            if (c == 0) {
                const int2 ix = pi[j];
                pend = __dadd_rn(__dadd_rn(pend, s_pm[ix.x & IMASK]), s_pm[ix.y & IMASK]);
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

/* phase B of one split chunk for one chain (alt_seq or alt_clip) */
template <int ASSOC>
__device__ __forceinline__ void c_replay_split(const double *px, int cnt, int lead, double &acc, double &pend)
{
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

/*
 * SEG: the launch list (cp.entries) names sites AND pieces of sites.  A piece is a run of whole 32-row chunks of
 * one site's fragment rows or split rows; the warp scores it like a site of its own (phase A is per chunk, so the
 * addends are the ones the unsplit site would park), but instead of summing them it copies each chunk's parked
 * addends to cp.scratch -- the order-sensitive sums of such a site are made afterwards, in row order, by
 * svgt_replay_pieces_kernel.  That takes a long site off the critical path of a small or heavy-tailed batch: its
 * chunks are scored by many warps at once, and the serial part left is one DADD per row.
 */
template <int G, int ASSOC, bool SEG>
__global__ void __launch_bounds__(SVGT_C_THREADS, SVGT_C_MINB) svgt_compact_kernel(const SvgtCompactParams cp)
{
    typedef CWarpSmem<G> WS;
    const SvgtParams &p = cp.base;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_pm[256];                            /* static: its shared address is a link-time constant */
    LibK *s_lib = reinterpret_cast<LibK *>(smem_raw);
    size_t off = (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    LibF *s_libf = reinterpret_cast<LibF *>(smem_raw + off);
    off += (kWLibs + 1) * sizeof(LibF);
    off = (off + 127) & ~(size_t)127;
    WS *s_warp = reinterpret_cast<WS *>(smem_raw + off);
    off += sizeof(WS) * kCWarps;
    unsigned *s_hist = reinterpret_cast<unsigned *>(smem_raw + off);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    int err = 0;
    const int nl = p.n_lib < SVGT_SMEM_LIBS ? p.n_lib : SVGT_SMEM_LIBS;
    for (int i = tid; i < 256; i += SVGT_C_THREADS) s_pm[i] = p.pm[i];
    for (int i = tid; i < nl; i += SVGT_C_THREADS) s_lib[i] = derive_lib(p, i, &err);
    WS &ws = s_warp[warp];
    if (lane < 2) ws.zero[lane] = 0.0;
    if (lane < 8) ws.lead[lane] = 0;
    const unsigned pm_addr = c_smem(s_pm);
    if (lane < kCD) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c_smem(&ws.bar[lane])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    /* the 32-bit form of 19 * h1 > h2 needs every histogram count below 2^26: the host states the largest
     * count (svgt_cbatch_t::hist_max); 0 = unknown, then the CTA looks for itself */
    bool small_counts;
    if (cp.hist_max != 0u) { small_counts = cp.hist_max < (1u << 26); __syncthreads(); }
    else {
        int big = 0;
        for (long long i = tid; i < p.n_hist; i += SVGT_C_THREADS) big |= p.hist[i] >= (1u << 26);
        small_counts = __syncthreads_or(big) == 0;
    }
    if (tid == 0) {
        const long long cap = c_hist_words(p.n_hist);
        long long base = 0;
        for (int l = 0; l <= kWLibs; ++l) {
            LibF f; f.addr = 0u; f.len = 0; f.ok = 0; f.pad = 0;
            if (l < kWLibs && l < nl) {
                const LibK &L = s_lib[l];
                const bool fits = L.hist_len < (1 << kHistLenBits) && base + L.hist_len + 1 <= cap;
                if (L.safe && small_counts && fits) {
                    f.addr = c_smem(s_hist + base); f.len = L.hist_len; f.ok = 1;
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
        unsigned *dst = s_hist + ((f.addr - c_smem(s_hist)) >> 2);
        const unsigned *src = p.hist + s_lib[l].hist_off;
        for (int i = tid; i <= f.len; i += SVGT_C_THREADS) dst[i] = i < f.len ? src[i] : 0u;
    }
    __syncthreads();

    Tables t;
    t.pm = s_pm; t.libs = s_lib; t.hist = p.hist;
    t.conc = p.consts[C_CONC]; t.disc = p.consts[C_DISC];
    const int m = p.min_aligned, slop = p.split_slop;
    const unsigned zero_addr = c_smem(&ws.zero[0]);
    const unsigned ring0 = c_smem(&ws.ring[0][0][0]);
    const unsigned bars = c_smem(&ws.bar[0]);

    const int gb = lane >> 2, c = lane & 3;                 /* phase-B role: chain c of site gb */
    const int ramp = cp.ramp;
    const long long W = (long long)gridDim.x * kCWarps;
    const long long n_entries = SEG ? cp.n_entries : p.n_sites;
    const long long n_units = c_n_units<G>(n_entries, W, ramp);
    unsigned tt = 0u;                                       /* super-steps consumed by this warp (slot / phase) */

    long long unit = 0, unit_next = 0;
    if (lane == 0) unit = (long long)atomicAdd(reinterpret_cast<unsigned *>(p.status + 1), 1u);
    unit = __shfl_sync(full, unit, 0);
    for (; unit < n_units; unit = __shfl_sync(full, unit_next, 0)) {
        /* ---- lanes 0..G-1 read their site row and publish the scalars ---- */
        int my_nf = 0, my_ns = 0;
        const int4 *my_rows = cp.rows;
        unsigned segmask = 0u;                              /* SEG: sites of this unit that are pieces */
        {
            const CUnit ur = c_unit_range<G>(unit, W, ramp);
            const long long idx = ur.base + lane;
            const bool valid = lane < ur.count && lane < G && idx < n_entries;
            long long site = 0;
            bool valid2 = valid;
            int4 pc = make_int4(0, 0, 0, 0);                /* SEG: the piece (site, first row, rows | split << 31, scratch chunk) */
            bool is_piece = false;
            if (valid) {
                if (SEG) {
                    const int e = cp.entries[idx];
                    if (e < 0) {
                        const long long k = (long long)~e;
                        if (k < cp.n_pieces) { pc = ldg4(cp.pieces + k); site = pc.x; is_piece = true; }
                        else site = -1;
                    } else site = e;
                } else site = p.order ? (long long)p.order[idx] : idx;
                if (site < 0 || site >= p.n_sites) { site = 0; valid2 = false; is_piece = false; err = SVGT_ERR_ARG; }   /* bad entry */
            }
            int4 a = make_int4(0, 0, 0, 0), b = a, d = a;
            if (valid2) {
                const int4 *sp = cp.sites + site * 3;
                a = ldg4(sp); b = ldg4(sp + 1); d = ldg4(sp + 2);
            }
            /* a = posA posB ciA0 ciA1; b = ciB0 ciB1 var_length meta; d = row_off(lo, hi) n_frag n_split */
            const int meta = b.w;
            const int4 cis = make_int4(b.x, b.y, 0, 0);
            const bool ranged = site_fields_in_range(a, cis, m, slop);
            const bool run = valid2 && !(meta & SITE_SKIP) && ranged;
            const long long roff = ((long long)(unsigned)d.x) | ((long long)d.y << 32);
            int nf = run ? d.z : 0, ns = run ? d.w : 0;
            if (nf < 0 || ns < 0 || roff < 0 || roff + nf + ns > cp.n_rows) { nf = 0; ns = 0; err = SVGT_ERR_ARG; }
            long long prow = roff;
            if (SEG && is_piece) {
                /* whole chunks of one part of the site: chunk boundaries are the unsplit site's */
                const int cnt = pc.z & 0x7fffffff;
                const bool psplit = pc.z < 0;
                const int part = psplit ? ns : nf;
                if (pc.y < 0 || (pc.y & 31) || cnt <= 0 || cnt > part - pc.y || pc.w < 0 ||
                    (long long)pc.w + ((cnt + 31) >> 5) > cp.scratch_chunks) {
                    nf = 0; ns = 0; is_piece = false; err = SVGT_ERR_ARG;
                } else {
                    prow = roff + (psplit ? nf : 0) + pc.y;
                    nf = psplit ? 0 : cnt; ns = psplit ? cnt : 0;
                }
            }
            if (SEG) segmask = __ballot_sync(full, is_piece) & ((1u << G) - 1u);
            my_nf = nf; my_ns = ns; my_rows = cp.rows + prow;
            if (lane < G) {
                SiteS &S = ws.site[lane];
                const int same = (meta >> 5) & 1;
                S.tA = 0; S.tB = same ? 0 : 1;            /* invented contig ids: rows carry class bits */
                S.wA0 = a.x - m; S.wA1 = a.x + m; S.wB0 = a.y - m; S.wB1 = a.y + m;
                S.meta = (meta & 15) | ((a.x - m >= 0) << 8) | ((a.y - m >= 0) << 9);
                S.var_length = b.z;
                S.posA = a.x; S.posB = a.y; S.ciA0 = a.z; S.ciA1 = a.w; S.ciB0 = b.x; S.ciB1 = b.y;
                S.dAB = a.y - a.x; S.nf = nf; S.foff = prow; S.soff = prow + nf; S.ns = ns;
                S.slot = (valid2 && !(SEG && is_piece)) ? (int)site : -1;   /* a piece's sums are made by the replay kernel */
                S.pad1 = pc.w; S.pad2 = 0;                                  /* pad1: the piece's first scratch chunk */
                CSiteF &F = ws.sf[lane];
                const int svtype = meta & 3;
                F.wA0 = a.x - m; F.wA1 = a.x + m; F.wB0 = a.y - m; F.wB1 = a.y + m;
                F.pat = (int)(CF_PAIRED | ((meta & SITE_O1_REV) ? CF_REV_A : 0u) | ((meta & SITE_O2_REV) ? CF_REV_B : 0u));
                F.del = svtype == SV_DEL;
                const bool okA = a.x - m >= 0, okB = a.y - m >= 0;      /* max(0, pos - m) cuts the window short: never covered */
                F.fast = (okA && okB && svtype != SV_INV && same) ? 1 : 2;
                F.m21 = 2 * m - 1;
                F.sgnA = (meta & SITE_O1_REV) ? 1 : -1; F.sgnB = (meta & SITE_O2_REV) ? 1 : -1;
                F.inv = svtype == SV_INV; F.same = same;
                F.mAA = okA ? CLS_A_ON_A : 0u; F.mAB = okB ? CLS_A_ON_B : 0u;
                F.mBA = okA ? CLS_B_ON_A : 0u; F.mBB = okB ? CLS_B_ON_B : 0u;
                ws.spf[lane] = make_csplitf(S, slop);
                ws.cnt[0][lane] = nf; ws.cnt[1][lane] = ns;
            }
            __syncwarp();
            for (int i = lane; i < G * (kWLibs + 1); i += 32) {
                const int g = i / (kWLibs + 1), l = i % (kWLibs + 1);
                if (ws.site[g].nf)
                    ws.wf[g][l] = make_winf(ws.site[g], s_lib[l < nl ? l : 0], s_libf[l < nl ? l : kWLibs], m, zero_addr);
            }
            __syncwarp();
        }
        /* claim the next unit now: the atomic's round trip is hidden behind this unit's rows */
        if (lane == 0) unit_next = (long long)atomicAdd(reinterpret_cast<unsigned *>(p.status + 1), 1u);

        /* super-steps: nsf over fragment rows, then nss over split rows */
        const int nsf = (__reduce_max_sync(full, my_nf) + 31) >> 5;
        const int nss = (__reduce_max_sync(full, my_ns) + 31) >> 5;
        const int T = nsf + nss;

        /* lane g issues the bulk copy of site g's rows of super-step k into ring slot `slot` */
        auto issue = [&](const int k, const unsigned slot) {
            const bool sp = k >= nsf;
            const int row0 = sp ? my_nf + (k - nsf) * 32 : k * 32;
            const int end = sp ? my_nf + my_ns : my_nf;
            int nrow = end - row0;
            nrow = nrow < 0 ? 0 : (nrow > 32 ? 32 : nrow);
            const unsigned bytes = (unsigned)nrow * 16u;
            const unsigned total = __reduce_add_sync(full, bytes);
            const unsigned bar = bars + slot * 8u;
            /* the slot was last written by this warp's own parked rows (generic proxy): order them before the
             * bulk copies (async proxy) that overwrite it */
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
            __syncwarp();
            if (bytes) {
                const unsigned dst = ring0 + (slot * G + (unsigned)lane) * 528u;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(my_rows + row0), "r"(bytes), "r"(bar) : "memory");
            }
        };

        double sum_frag = 0.0, sum_split = 0.0;
        double acc = 0.0, pend = 0.0;
#pragma unroll 1
        for (int k = 0; k < kCD - 1 && k < T; ++k) issue(k, (tt + (unsigned)k) % (unsigned)kCD);
#pragma unroll 1
        for (int k = 0; k < T; ++k, ++tt) {
            const unsigned slot = tt % (unsigned)kCD;
            if (k + kCD - 1 < T) issue(k + kCD - 1, (tt + (unsigned)(kCD - 1)) % (unsigned)kCD);
            {
                const unsigned bar = bars + slot * 8u, parity = (tt / (unsigned)kCD) & 1u;
                unsigned ok;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                } while (!ok);
            }
            const bool sp = k >= nsf;
            const int step = sp ? k - nsf : k;
            const int *cnts = &ws.cnt[sp ? 1 : 0][0];
            const unsigned slotaddr = ring0 + slot * (unsigned)(G * 528);
            auto load_row = [&](const int g, const int n) -> int4 {
                int4 r;
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                             : "r"(slotaddr + (unsigned)g * 528u + (unsigned)lane * 16u));
                /* the copy stopped at the site's last row: a row with zero flag / class words scores as nothing,
                 * whatever the stale coordinates are */
                if (lane >= n) { r.z = 0; r.w = 0; }
                return r;
            };
            auto park_frag = [&](const int g, const int4 r, const CRow &a, const FragOut &fo) {
                const unsigned pk = slotaddr + (unsigned)g * 528u + (unsigned)lane * 8u;
                ws.spark[g][lane] = ASSOC == SVGT_ASSOC_CLASSIC ? crow_lut_pair(r, a) : fo.s;
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(pk), "d"(fo.p_ref) : "memory");
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(pk + 264u), "d"(fo.p_alt) : "memory");
            };
            if (!sp) {
#if SVGT_C_PREFETCH
                /* the next site's row is fetched while this one is scored (parking only touches this site's bytes) */
                auto raw_row = [&](const int g) -> int4 {
                    int4 r;
                    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                                 : "r"(slotaddr + (unsigned)g * 528u + (unsigned)lane * 16u));
                    return r;
                };
                int4 rn = raw_row(0);
#endif
#pragma unroll kCUnroll
                for (int g = 0; g < G; ++g) {
#if SVGT_C_PREFETCH
                    int4 r = rn;
                    if (g + 1 < G) rn = raw_row(g + 1);
                    const int n = cnts[g] - step * 32;
                    if (n <= 0) continue;
                    if (lane >= n) { r.z = 0; r.w = 0; }
#else
                    const int n = cnts[g] - step * 32;
                    if (n <= 0) continue;
                    const int4 r = load_row(g, n);
#endif
                    __syncwarp();
                    const CSiteF &F = ws.sf[g];
                    CRow a;
                    crow_stage1(F, &ws.wf[g][0], pm_addr, r, F.fast != 1, a);
                    if (__any_sync(full, crow_is_rare(r, a))) {     /* MULTI / CONT rows, ties: fix-ups, folds, lead rows */
                        crow_stage2(p, t, ws.site[g], F, s_lib, lane, n, m, r, a, err);
                        const FragOut fo = crow_stage3<ASSOC, true>(lane, n, r, a);
                        park_frag(g, r, a, fo);
                        if (ASSOC != SVGT_ASSOC_CLASSIC && fo.lead != 0) {                      /* warp-uniform */
                            if (lane < fo.lead) ws.spark[g][lane] = crow_lut_pair(r, a);
                            if (lane == 0) ws.lead[g] = fo.lead;
                        }
                    } else {
                        park_frag(g, r, a, crow_stage3<ASSOC, false>(lane, n, r, a));
                    }
                }
            } else {
#pragma unroll 1
                for (int g = 0; g < G; ++g) {
                    const int n = cnts[g] - step * 32;
                    if (n <= 0) continue;
                    const int4 r = load_row(g, n);
                    __syncwarp();
                    const unsigned pk = slotaddr + (unsigned)g * 528u + (unsigned)lane * 8u;
                    const SplitOut so = score_csplit_chunk<ASSOC>(ws.spf[g], pm_addr, lane, n, r);
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(pk), "d"(so.vseq) : "memory");
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(pk + 264u), "d"(so.vclip) : "memory");
                    if (__builtin_expect(so.lead != 0, 0) && lane == 0) ws.lead[g] = so.lead;
                }
            }
            /* ---- phase B: the ordered replay of this super-step's parked rows ---- */
            __syncwarp();
            if (SEG && segmask) {                           /* pieces: the chunk's parked addends (and its lead count) go to scratch */
#pragma unroll 1
                for (int g = 0; g < G; ++g) {
                    if (!((segmask >> g) & 1u) || cnts[g] - step * 32 <= 0) continue;
                    const long long q = (long long)ws.site[g].pad1 + step;
                    double *dst = cp.scratch + q * 96;
                    const double *pr = reinterpret_cast<const double *>(&ws.ring[slot][g][0]);
                    dst[lane] = ws.spark[g][lane];
                    dst[32 + lane] = pr[lane];
                    dst[64 + lane] = pr[33 + lane];
                    if (lane == 0) cp.scratch_lead[q] = ws.lead[g];
                }
            }
            if (gb < G && c < (sp ? 2 : 3) && !(SEG && ((segmask >> gb) & 1u))) {
                int cnt = cnts[gb] - step * 32;
                cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
                const int lead = ws.lead[gb];
                const double *pr = reinterpret_cast<const double *>(&ws.ring[slot][gb][0]);
                if (!sp) c_replay_frag<ASSOC>(c == 0 ? &ws.spark[gb][0] : pr + (c - 1) * 33, &ws.spark[gb][0], c, cnt, lead, s_pm,
                                              acc, pend);
                else c_replay_split<ASSOC>(pr + c * 33, cnt, lead, acc, pend);
            }
            __syncwarp();
            if (lane < 8) ws.lead[lane] = 0;
            if (k == nsf - 1) {                             /* the fragment rows are done */
                if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
                sum_frag = acc; acc = 0.0; pend = 0.0;
            }
        }
        if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
        sum_split = acc;

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
        __syncwarp();
    }
    if (err) {
        atomicCAS(p.status, 0, err);
        atomicAdd(p.status + 2, 1);
    }
}

/* zeroing rules + genotype call of one site on its parked sums (reference singlesample.py:382-473): reads the
 * site's five sums from its row of base.out, writes the final 80-byte row (to out_final when given).
 */
template <bool PIPE>
__device__ __forceinline__ void c_call_one(const SvgtCompactParams &cp, const long long site, int &err)
{
    const SvgtParams &p = cp.base;
    const int4 *sp = cp.sites + site * 3;
    const int4 a = ldg4(sp), b = ldg4(sp + 1), d = ldg4(sp + 2);
    const int meta = b.w;
    svgt_out_row_t o;
    o.gl[0] = o.gl[1] = o.gl[2] = 0.0; o.sq = 0.0;
    o.gt = 0; o.gq = 0; o.dp = 0; o.ro = 0; o.ao = 0; o.qr = 0; o.qa = 0;
    o.rs = 0; o.as_ = 0; o.asc = 0; o.rp = 0; o.ap = 0;
    if (meta & SITE_SKIP) { o.gt = SVGT_GT_SKIPPED; o.gq = -1; }
    else if (!site_fields_in_range(a, make_int4(b.x, b.y, 0, 0), p.min_aligned, p.split_slop)) {
        o.gt = SVGT_GT_BLANK; o.gq = -1; err = SVGT_ERR_RANGE;
    } else {
        ParkedSums s = {0.0, 0.0, 0.0, 0.0, 0.0};
        const long long roff = ((long long)(unsigned)d.x) | ((long long)d.y << 32);
        const bool ok = !(d.z < 0 || d.w < 0 || roff < 0 || roff + d.z + d.w > cp.n_rows);
        if (ok && (d.z > 0 || d.w > 0)) {
            const double *row = reinterpret_cast<const double *>(p.out + site);
            s.ref_seq = row[0]; s.alt_seq = row[1]; s.alt_clip = row[2]; s.ref_span = row[3]; s.alt_span = row[4];
        }
        Tables t;
        t.pm = p.pm; t.libs = nullptr; t.hist = p.hist;
        t.conc = 0.0; t.disc = 0.0;
        call_site<PIPE>(p, t, meta & 3, s.ref_seq, s.alt_seq, s.alt_clip, s.ref_span, s.alt_span, o, err);
    }
    int4 *dst = reinterpret_cast<int4 *>((cp.out_final ? cp.out_final : p.out) + site);
    const int4 *src = reinterpret_cast<const int4 *>(&o);
#pragma unroll
    for (int i = 0; i < 5; ++i) dst[i] = src[i];
}

/* one site per thread.  SMALL (batches of up to kCallSmallMax sites, where the longest log_choose chain -- two
 * dependent fp64 adds per step, statistics.py:9-20 -- is a visible share of the step): the sites are walked in launch
 * order when there is one (work-descending: the long chains start first and share their warps with chains of similar
 * length; -10 % on the 10k-site shape) and the chain's LUT values are fetched ahead.  Larger batches keep the index
 * order, whose coalesced site rows matter more (+0.8 % at 1M sites otherwise), and the leaner loop (fewer registers:
 * more resident threads) */
constexpr long long kCallSmallMax = 1 << 18;

template <bool SMALL>
__global__ void __launch_bounds__(256) svgt_call_compact_kernel(const SvgtCompactParams cp)
{
    const SvgtParams &p = cp.base;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int err = 0;
    if (idx < p.n_sites) {
        const long long site = (SMALL && p.order) ? (long long)p.order[idx] : idx;
        if (site >= 0 && site < p.n_sites) c_call_one<SMALL>(cp, site, err);     /* a bad entry was flagged by the tally kernel */
    }
    if (err) {
        atomicCAS(p.status, 0, err);
        atomicAdd(p.status + 2, 1);
    }
    /* multi-GPU: the last CTA to finish tells the gathering rank this shard's rows have landed */
    if (cp.done_flag) {
        /* the CTA's rows are ordered before thread 0's system-scope fence by the barrier (fences are cumulative),
         * so ONE fence per CTA publishes them; the last CTA's thread 0 has then observed every other CTA's count */
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(p.status + 3, 1) == (int)gridDim.x - 1) {
                __threadfence_system();
                *reinterpret_cast<volatile int *>(cp.done_flag) = cp.done_value;
                __threadfence_system();
            }
        }
    }
}

/*
 * The ordered sums of the sites that were scored in pieces (SEG): one warp per such site walks the site's scratch
 * chunks in row order -- fragment chunks, then split chunks -- and replays them with the same c_replay_frag /
 * c_replay_split as phase B of the tally kernel (lane c = chain c), so the sums are the ones the unsplit site would
 * get (reference singlesample.py:364-378: the fp64 sums run over sorted(query_name) fragments one after another).
 * Each lane fetches its row of the next kRD chunks (coalesced 256-byte reads of L2-resident scratch) while the
 * chains of the current ones are replayed out of shared memory.
 */
constexpr int kRWarps = 4;          /* warps (= sites) per CTA */
constexpr int kRD = 4;              /* chunks fetched ahead */

template <int ASSOC>
__global__ void __launch_bounds__(kRWarps * 32) svgt_replay_pieces_kernel(const SvgtCompactParams cp)
{
    const SvgtParams &p = cp.base;
    __shared__ double s_pm[256];
    __shared__ double s_buf[kRWarps][2][kRD][3][33];
    __shared__ int s_lead[kRWarps][2][kRD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += kRWarps * 32) s_pm[i] = p.pm[i];
    __syncthreads();
    const long long h = (long long)blockIdx.x * kRWarps + warp;
    if (h >= cp.n_heavy) return;
    const int4 hv = ldg4(cp.heavy + h);                      /* site, first scratch chunk, fragment chunks, split chunks */
    const long long site = hv.x;
    int nf = 0, ns = 0;
    bool ok = site >= 0 && site < p.n_sites && hv.y >= 0 && hv.z >= 0 && hv.w >= 0 &&
              (long long)hv.y + hv.z + hv.w <= cp.scratch_chunks;
    if (ok) {
        const int4 d = ldg4(cp.sites + site * 3 + 2);
        nf = d.z; ns = d.w;
        ok = nf >= 0 && ns >= 0 && hv.z == ((nf + 31) >> 5) && hv.w == ((ns + 31) >> 5);
    }
    if (!ok) {
        if (lane == 0) { atomicCAS(p.status, 0, SVGT_ERR_ARG); atomicAdd(p.status + 2, 1); }
        return;
    }
    const int cf = hv.z, T = hv.z + hv.w;
    const double *src = cp.scratch + (long long)hv.y * 96;
    const int *lsrc = cp.scratch_lead + hv.y;
    double v[kRD][3];
    int ld = 0;
    auto fetch = [&](const int t0) {
#pragma unroll
        for (int i = 0; i < kRD; ++i) {
            const bool in = t0 + i < T;
            const double *q = src + (long long)(t0 + i) * 96;
            v[i][0] = in ? q[lane] : 0.0; v[i][1] = in ? q[32 + lane] : 0.0; v[i][2] = in ? q[64 + lane] : 0.0;
        }
        ld = (lane < kRD && t0 + lane < T) ? lsrc[t0 + lane] : 0;
    };
    double acc = 0.0, pend = 0.0, sum_frag = 0.0;
    fetch(0);
    for (int t0 = 0, b = 0; t0 < T; t0 += kRD, b ^= 1) {
#pragma unroll
        for (int i = 0; i < kRD; ++i) {
            s_buf[warp][b][i][0][lane] = v[i][0]; s_buf[warp][b][i][1][lane] = v[i][1]; s_buf[warp][b][i][2][lane] = v[i][2];
        }
        if (lane < kRD) s_lead[warp][b][lane] = ld;
        __syncwarp();
        if (t0 + kRD < T) fetch(t0 + kRD);
        if (lane < 3) {
            const int c = lane;
            for (int i = 0; i < kRD && t0 + i < T; ++i) {
                const int t = t0 + i;
                const bool sp = t >= cf;
                const int left = sp ? ns - (t - cf) * 32 : nf - t * 32;
                const int cnt = left > 32 ? 32 : left;
                const int lead = max(s_lead[warp][b][i], 0);   /* scratch of a piece a malformed plan never scored holds anything */
                if (!sp) c_replay_frag<ASSOC, 255>(&s_buf[warp][b][i][c][0], &s_buf[warp][b][i][0][0], c, cnt, lead, s_pm, acc, pend);
                else if (c < 2) c_replay_split<ASSOC>(&s_buf[warp][b][i][c + 1][0], cnt, lead, acc, pend);
                if (t == cf - 1) {                          /* the fragment rows are done */
                    if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
                    sum_frag = acc; acc = 0.0; pend = 0.0;
                }
            }
        }
        __syncwarp();
    }
    if (lane < 3) {
        if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
        const double sum_split = acc;
        double *row = reinterpret_cast<double *>(p.out + site);
        /* ParkedSums: ref_seq, alt_seq, alt_clip, ref_span, alt_span */
        if (lane == 0) { row[0] = sum_frag; row[1] = sum_split; }
        else if (lane == 1) { row[3] = sum_frag; row[2] = sum_split; }
        else row[4] = sum_frag;
    }
}

template <int G>
size_t c_smem_bytes(const SvgtParams &p)
{
    size_t off = (size_t)SVGT_SMEM_LIBS * sizeof(LibK) + (kWLibs + 1) * sizeof(LibF);
    off = (off + 127) & ~(size_t)127;
    off += sizeof(CWarpSmem<G>) * kCWarps;
    off += (size_t)c_hist_words(p.n_hist) * sizeof(unsigned);
    return off;
}

typedef void (*CKernel)(const SvgtCompactParams);

template <int G, int ASSOC>
CKernel c_pick_kernel(bool seg)
{
    if constexpr (G == SVGT_C_G) {
        if (seg) return svgt_compact_kernel<G, ASSOC, true>;
    }
    return svgt_compact_kernel<G, ASSOC, false>;
}

struct CLaunchInfo { int ready[16]; int per_sm[16]; int sms[16]; size_t smem_set[16]; };

template <int G>
int launch_compact(const SvgtCompactParams &cp, int ramp, cudaStream_t stream, bool force_ramp = false)
{
    static CLaunchInfo info[4] = {};
    const SvgtParams &p = cp.base;
    const int a = p.assoc_mode == SVGT_ASSOC_CLASSIC ? 1 : 0;
    const bool seg = cp.entries != nullptr && G == SVGT_C_G;        /* a piece plan: the launch list names sites and pieces */
    auto kern = a ? c_pick_kernel<G, SVGT_ASSOC_CLASSIC>(seg) : c_pick_kernel<G, SVGT_ASSOC_SSO>(seg);
    const size_t smem = c_smem_bytes<G>(p);
    cudaError_t e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    CLaunchInfo &li = info[a + (seg ? 2 : 0)];
    const int di = dev & 15;
    if (!li.ready[di] || li.smem_set[di] < smem) {
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return (int)e;
        if ((e = cudaDeviceGetAttribute(&li.sms[di], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&li.per_sm[di], kern, SVGT_C_THREADS, smem)) != cudaSuccess)
            return (int)e;
        if (li.per_sm[di] < 1) li.per_sm[di] = 1;
        li.smem_set[di] = smem; li.ready[di] = 1;
    }
    const long long cap = (long long)li.sms[di] * li.per_sm[di];      /* persistent: one resident wave */
    const long long n_entries = seg ? cp.n_entries : p.n_sites;
    const long long units = ramp ? n_entries : (n_entries + G - 1) / G;
    const long long want = (units + kCWarps - 1) / kCWarps;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    /* large batches amortise their longest unit: no ramp above `ramp_sites` sites per resident warp
     * (SVGT_C_RAMP_PER_WARP overrides the threshold for measurements) */
    static const long long ramp_per_warp = [] {
        const char *v = getenv("SVGT_C_RAMP_PER_WARP");
        return v && *v ? atoll(v) : (long long)SVGT_C_RAMP_PER_WARP_DEFAULT;
    }();
    if (ramp == 1 && !force_ramp && n_entries >= ramp_per_warp * cap * kCWarps) ramp = 0;
    SvgtCompactParams q = cp;
    q.ramp = ramp;
    kern<<<grid, SVGT_C_THREADS, smem, stream>>>(q);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    if (seg && cp.n_heavy > 0) {
        const int rgrid = (int)((cp.n_heavy + kRWarps - 1) / kRWarps);
        if (a) svgt_replay_pieces_kernel<SVGT_ASSOC_CLASSIC><<<rgrid, kRWarps * 32, 0, stream>>>(q);
        else svgt_replay_pieces_kernel<SVGT_ASSOC_SSO><<<rgrid, kRWarps * 32, 0, stream>>>(q);
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    }
    const int cgrid = (int)((p.n_sites + 255) / 256);
    if (p.n_sites <= kCallSmallMax) svgt_call_compact_kernel<true><<<cgrid, 256, 0, stream>>>(q);
    else svgt_call_compact_kernel<false><<<cgrid, 256, 0, stream>>>(q);
    return (int)cudaGetLastError();
}

}  // namespace

int svgt_launch_compact(const SvgtCompactParams &cp, int unit_mode, cudaStream_t stream)
{
    /* unit_mode: 0 ramped units when the batch is small (by site count), 1 always G-site units, 2 always two-site
     * units, 3 always ramped units (the caller knows the batch is heavy-tailed: Engine picks 1 or 3 from the sites'
     * row counts) */
    if (unit_mode == 2) return launch_compact<2>(cp, 0, stream);
    return launch_compact<SVGT_C_G>(cp, unit_mode == 1 ? 0 : 1, stream, unit_mode == 3);
}
