/*
 * svgt_ring.cu -- warp-cooperative tally kernel with TMA-staged rows (variant 4).
 *
 * Same two-phase mapping as svgt_coop.cu (phase A: one evidence row per lane; phase B: one ordered
 * fp64 chain per lane, singlesample.py:355-404), but the rows do not travel through registers:
 *
 *   - every 32-row chunk of a site (fragment rows first, then split rows: both are 32-byte rows,
 *     so one stream) is ONE cp.async.bulk (TMA 1-D, UBLKCP) of <= 1 KB from HBM into a slot of a
 *     per-warp shared-memory ring, completing on that slot's mbarrier.  The producer side is just
 *     the same warp running up to RS - 4 chunks ahead; completion is tracked by mbarrier phases,
 *     not by register scoreboards, so the copies really are in flight while earlier chunks are
 *     scored (register prefetch deeper than one chunk did not overlap: svgt_coop.cu, profiles/).
 *   - phase A scores a slot IN PLACE: lane l reads its 32-byte row and overwrites it with the
 *     32 bytes it parks for phase B {a + b, LUT indices of a and b, p_ref, p_alt}.
 *   - phase B runs after every group of <= 4 chunks of one super-step: lane 4g+c replays chain c
 *     of site g straight out of the ring slot; the slot is then free for the producer again.
 *
 * The genotype call is the same second launch (svgt_call_kernel in svgt_coop.cu).
 */
#include "svgt_coop.cuh"

namespace {

constexpr int RG = 8;               /* sites per work unit                                        */
constexpr int RS = 4;               /* raw ring slots per warp = chunks in flight                  */
constexpr int kGroup = 4;           /* chunks per phase-B group = parked slots                     */
constexpr int kParkBytes = 33 * 32; /* 32 rows + 32 B pad: chain lanes of different sites hit distinct banks */
constexpr int kFifo = 8;            /* chunk descriptors handed from the issue side to the scoring side */

/* issue side -> scoring side, one per chunk in flight (warp-uniform) */
struct ChunkDesc { int n, g, tag, pad; };   /* rows, site, phase << 24 | step (tag -1 = end of unit) */
/* scoring side -> chain lanes, one per site (warp-uniform) */
struct Pending { int slot_cnt; unsigned newmask; };   /* (parked slot + 1) << 8 | rows, 0 = nothing pending */

struct alignas(128) RingSmem {
    SiteS site[RG];
    Win win[RG][kWLibs];
    unsigned char raw[RS][1024];            /* cp.async.bulk targets                    */
    unsigned char park[kGroup][kParkBytes]; /* phase A -> phase B                        */
    unsigned long long bar[RS];
    ChunkDesc desc[kFifo];
    Pending pend[RG];
    double zero[2];
};

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

/* warp-uniform cursor over a unit's chunk stream: fragment chunks step-major, then split chunks */
struct Cursor {
    int phase, step;
    unsigned mask;
};

template <int ASSOC>
__global__ void __launch_bounds__(SVGT_COOP_THREADS, 2) svgt_ring_kernel(const SvgtParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_pm = reinterpret_cast<double *>(smem_raw);
    LibK *s_lib = reinterpret_cast<LibK *>(s_pm + 512);     /* pm[0..255], then pm[q] / 2 */
    size_t off = 512 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    off = (off + 127) & ~(size_t)127;
    RingSmem *s_warp = reinterpret_cast<RingSmem *>(smem_raw + off);
    off += sizeof(RingSmem) * kCoopWarps;
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
    int big = 0;
    for (long long i = tid; i < p.n_hist; i += SVGT_COOP_THREADS) {
        const unsigned v = p.hist[i];
        if (p.hist_in_smem) s_hist[i] = v;
        big |= v >= (1u << 26);
    }
    RingSmem &ws = s_warp[warp];
    if (lane < 2) ws.zero[lane] = 0.0;
    if (lane < RS) mbar_init(&ws.bar[lane], 1);
    if (lane < RG) { ws.pend[lane].slot_cnt = 0; ws.pend[lane].newmask = 0u; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const bool small_counts = __syncthreads_or(big) == 0;

    Tables t;
    t.pm = s_pm; t.libs = s_lib; t.hist = p.hist_in_smem ? s_hist : p.hist;
    t.conc = p.consts[C_CONC]; t.disc = p.consts[C_DISC];
    const unsigned *hist = t.hist;
    const int m = p.min_aligned, slop = p.split_slop;

    const int gb = lane >> 2, c = lane & 3;                /* phase-B role: chain c of site gb */
    const long long n_units = (p.n_sites + RG - 1) / RG;
    /* running counters of this warp: FIFO entries (chunks + one end-of-unit sentinel per unit) and
     * chunks proper (ring slot and mbarrier phase follow the latter) */
    unsigned f_issue = 0, f_cons = 0, q_issue = 0, q_cons = 0;

    for (;;) {
        long long unit = 0;
        if (lane == 0) unit = (long long)atomicAdd(reinterpret_cast<unsigned *>(p.status + 1), 1u);
        unit = __shfl_sync(full, unit, 0);
        if (unit >= n_units) break;

        /* ---- lanes 0..RG-1 read their site row and publish the scalars ---- */
        int my_nf = 0, my_ns = 0;
        {
            const long long idx = unit * RG + lane;
            const bool valid = lane < RG && idx < p.n_sites;
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
            my_nf = nf; my_ns = ns;
            if (lane < RG) {
                SiteS &S = ws.site[lane];
                S.tA = b.z; S.tB = b.w;
                S.wA0 = a.x - m; S.wA1 = a.x + m; S.wB0 = a.y - m; S.wB1 = a.y + m;
                S.meta = (meta & 15) | ((a.x - m >= 0) << 8) | ((a.y - m >= 0) << 9);
                S.var_length = cc.x;
                S.posA = a.x; S.posB = a.y; S.ciA0 = a.z; S.ciA1 = a.w; S.ciB0 = b.x; S.ciB1 = b.y;
                S.dAB = a.y - a.x; S.nf = nf; S.foff = foff; S.soff = soff; S.ns = ns;
                S.slot = valid ? 1 : 0;
            }
            __syncwarp();
            for (int i = lane; i < RG * kWLibs; i += 32) {
                const int g = i / kWLibs, l = i % kWLibs;
                if (l < nl) { if (ws.site[g].nf) ws.win[g][l] = make_win(ws.site[g], s_lib[l], m, small_counts); }
                else ws.win[g][l].flags = 0u;
            }
            __syncwarp();
        }

        /* ---- issue side: a cursor over the unit's chunk stream (fragment chunks step-major, then
         *      split chunks), one cp.async.bulk per chunk, descriptor into the FIFO ---- */
        Cursor prod = {0, -1, 0u};
        bool prod_more = true;
        auto issue_one = [&]() {
            if (!prod_more) return;
            int ph = 0, st = 0, g = 0;
            for (;;) {
                if (prod.mask) {
                    g = __ffs(prod.mask) - 1;
                    prod.mask &= prod.mask - 1u;
                    ph = prod.phase; st = prod.step;
                    break;
                }
                ++prod.step;
                const unsigned mk = __ballot_sync(full, (prod.phase == 0 ? my_nf : my_ns) > prod.step * 32);
                if (mk) { prod.mask = mk; continue; }
                if (prod.phase == 0) { prod.phase = 1; prod.step = -1; continue; }
                prod_more = false;
                break;
            }
            ChunkDesc d;
            if (prod_more) {
                const SiteS &S = ws.site[g];
                const int cnt = ph == 0 ? S.nf : S.ns;
                d.n = min(32, cnt - st * 32); d.g = g; d.tag = (ph << 24) | st; d.pad = 0;
                const int4 *src = (ph == 0 ? p.frags + 2 * (S.foff + (long long)st * 32)
                                           : p.splits + 2 * (S.soff + (long long)st * 32));
                const int sl = q_issue % RS;
                if (lane == 0) {
                    mbar_expect_tx(&ws.bar[sl], (unsigned)d.n * 32u);
                    bulk_g2s(&ws.raw[sl][0], src, (unsigned)d.n * 32u, &ws.bar[sl]);
                }
            } else {
                d.n = 0; d.g = 0; d.tag = -1; d.pad = 0;
            }
            if (lane == 0) *reinterpret_cast<int4 *>(&ws.desc[f_issue % kFifo]) = make_int4(d.n, d.g, d.tag, 0);
            ++f_issue;
            if (prod_more) ++q_issue;
        };
        for (int i = 0; i < RS; ++i) issue_one();
        __syncwarp();

        double sum_frag = 0.0, sum_split = 0.0;
        double acc = 0.0, pend = 0.0;
        unsigned carryA = 0u, carryB = 0u;
        int pending = 0;

        for (;;) {
            const int4 dq = *reinterpret_cast<const int4 *>(&ws.desc[f_cons % kFifo]);     /* n g tag */
            if (dq.z < 0) { ++f_cons; break; }                                             /* end of unit */
            const int n = dq.x, c_g = dq.y, c_phase = dq.z >> 24;
            const int sl = q_cons % RS;
            mbar_wait(&ws.bar[sl], (q_cons / RS) & 1u);
            const int4 *rowp = reinterpret_cast<const int4 *>(&ws.raw[sl][0]) + 2 * lane;
            const bool rv = lane < n;
            int4 lo = make_int4(0, 0, 0, 0), hi = lo;
            if (rv) { lo = rowp[0]; hi = rowp[1]; }
            const SiteS &S = ws.site[c_g];
            unsigned char *parkp = &ws.park[pending][0] + 32 * lane;
            int lead;
            if (c_phase == 0) {
                /* ---------------- phase A, fragment rows ---------------- */
                const FragOut fo = score_frag_chunk<ASSOC>(p, t, S, &ws.win[c_g][0], s_pm, s_lib, hist, lane, n, c_g, m, lo, hi,
                                                           carryA, carryB, err);
                park_frag(parkp, fo);
                lead = fo.lead;
            } else {
                /* ---------------- phase A, split rows ---------------- */
                const SplitOut so = score_split_chunk<ASSOC>(S, s_pm, lane, n, slop, lo, hi);
                *reinterpret_cast<double2 *>(parkp) = make_double2(so.vseq, so.vclip);
                lead = so.lead;
            }
            if (lane == 0) {
                Pending pd;
                pd.slot_cnt = ((pending + 1) << 8) | n;
                pd.newmask = (unsigned)lead;
                *reinterpret_cast<int2 *>(&ws.pend[c_g]) = make_int2(pd.slot_cnt, (int)pd.newmask);
            }
            ++pending;
            ++q_cons;
            ++f_cons;
            /* the rows of this chunk are in registers (their values have been used above): refill its slot */
            issue_one();
            __syncwarp();

            /* ---------------- phase B: close the group ---------------- */
            const int n_tag = ws.desc[f_cons % kFifo].tag;          /* the chunk after this one, or -1 */
            const bool phase_ends = n_tag < 0 || (n_tag >> 24) != c_phase;
            if (pending == kGroup || n_tag != dq.z) {
                const int2 pd = *reinterpret_cast<const int2 *>(&ws.pend[gb]);
                const int sl_b = (pd.x >> 8) - 1, cnt_b = pd.x & 0xFF;
                const int lead_b = pd.y;
                if (sl_b >= 0 && c < (c_phase == 0 ? 3 : 2)) {
                    if (c_phase == 0) replay_frag<ASSOC>(&ws.park[sl_b][0], c, cnt_b, lead_b, s_pm, acc, pend);
                    else replay_split<ASSOC>(&ws.park[sl_b][0], c, cnt_b, lead_b, acc, pend);
                }
                __syncwarp();
                if (lane < RG) ws.pend[lane].slot_cnt = 0;
                pending = 0;
                if (phase_ends) {
                    if (ASSOC == SVGT_ASSOC_SSO) acc = __dadd_rn(acc, pend);
                    if (c_phase == 0) sum_frag = acc; else sum_split = acc;
                    acc = 0.0; pend = 0.0;
                }
                __syncwarp();
            }
        }

        /* ---- park the five sums in the site's output row (lane 4g+c holds chain c of site g) ---- */
        if (gb < RG && c < 3) {
            const long long idx = unit * RG + gb;
            if (idx < p.n_sites && (ws.site[gb].nf | ws.site[gb].ns)) {
                const long long site = p.order ? (long long)p.order[idx] : idx;
                double *row = reinterpret_cast<double *>(p.out + site);
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

size_t ring_smem_bytes(const SvgtParams &p)
{
    size_t off = 512 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    off = (off + 127) & ~(size_t)127;
    off += sizeof(RingSmem) * kCoopWarps;
    if (p.hist_in_smem) off += (size_t)p.n_hist * sizeof(unsigned);
    return off;
}

}  // namespace

int svgt_launch_ring(const SvgtParams &p, cudaStream_t stream)
{
    auto kern = (p.assoc_mode == SVGT_ASSOC_CLASSIC) ? svgt_ring_kernel<SVGT_ASSOC_CLASSIC>
                                                     : svgt_ring_kernel<SVGT_ASSOC_SSO>;
    const size_t smem = ring_smem_bytes(p);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SVGT_COOP_THREADS, smem)) != cudaSuccess)
        return (int)e;
    if (per_sm < 1) per_sm = 1;
    const long long units = (p.n_sites + RG - 1) / RG;
    const long long want = (units + kCoopWarps - 1) / kCoopWarps;
    const long long cap = (long long)sms * per_sm;      /* persistent: one resident wave */
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, SVGT_COOP_THREADS, smem, stream>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    return svgt_launch_call(p, stream);
}
