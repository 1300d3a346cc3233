/*
 * svgt_kernels.cu -- sm_100a scoring kernels of the SV genotype-likelihood engine.
 *
 * One launch scores a batch of breakpoints: the per-fragment evidence scoring and the
 * Bayesian genotype call of hall-lab/svtyper v0.7.1 (citations relative to the
 * reference tree):
 *
 *   tally_variant_read_fragments   svtyper/singlesample.py:355-404 (classic.py:286-435)
 *     gather_split_read_evidence   svtyper/singlesample.py:246-276
 *     gather_paired_end_evidence   svtyper/singlesample.py:278-353
 *     SamFragment.is_ref_seq / is_pair_straddle / p_concordant   svtyper/parsers.py:801-882
 *     SplitRead.check_split_support / is_split_straddle          svtyper/parsers.py:1122-1215
 *   bayesian_genotype              svtyper/singlesample.py:406-473 (classic.py:437-495)
 *     bayes_gt / log_choose        svtyper/statistics.py:9-37
 *
 * Mapping onto the GPU (DESIGN.md has the full argument):
 *   - The integer FORMAT fields are int() of fp64 sums accumulated in sorted(qname)
 *     order, so each site's rows are folded SEQUENTIALLY by one thread:
 *     one site per lane, 32 work-bucketed sites per warp, persistent CTAs pulling
 *     32-site tiles from an atomic cursor.
 *   - Rows reach a lane either by 2x LDG.128 with a register prefetch (variant 0) or
 *     through a per-warp ring of cp.async.bulk (TMA 1-D) copies completing on
 *     mbarriers (variant 1): lane l streams ITS site's rows into ITS padded smem slot.
 *   - No device transcendental touches an integer field: prob_mapq, log10(n) and the
 *     prior constants are host LUTs; the device does IEEE add/mul/div
 *     with contraction off (explicit __d*_rn, and -fmad=false).
 *   - Window tests are integerised per (site, library): for integer i and real flank,
 *       i <  X - flank  <=>  i <  X - floor(flank)        i >  X + flank  <=>  i >  X + floor(flank)
 *     which is exact when the fp64 rounding of X -/+ flank cannot cross an integer
 *     (frac(flank) == 0 or in [2^-20, 1-2^-20], |X| < 2^31).  Libraries that fail the
 *     test take the literal fp64 path.  p_concordant() is the integer test 19*h1 > h2
 *     except on the exact tie, where the literal fp64 expression is evaluated.
 */
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "svgt_kernels.cuh"

#include "svgt_device.cuh"

namespace {

/* ---------------- mbarrier / bulk-copy PTX ---------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/*
 * Per-warp row stream.  DIRECT: each lane reads its own rows with 2x LDG.128 and a
 * one-row register prefetch.  BULK: each lane issues one cp.async.bulk per step for the
 * next SVGT_STAGE_ROWS rows of ITS site into ITS padded slot; the warp's stage barrier
 * completes when all 32 copies have landed.
 */
struct Ring {
    char *slots;        /* [kStages][32][kSlotBytes] for this warp */
    uint64_t *bars;     /* [kStages] for this warp                 */
    uint32_t step;      /* running step counter (stage = step % kStages, parity from step / kStages) */
};

template <int VARIANT, typename F>
__device__ __forceinline__ void stream_rows(Ring &ring, const int4 *rows, int n, int lane, F &&f)
{
    if (VARIANT == SVGT_VAR_DIRECT) {
        if (n <= 0) return;
        int4 nlo = ldg4(rows), nhi = ldg4(rows + 1);
        for (int j = 0; j < n; ++j) {
            const int4 lo = nlo, hi = nhi;
            if (j + 1 < n) { nlo = ldg4(rows + 2 * (j + 1)); nhi = ldg4(rows + 2 * (j + 1) + 1); }
            f(lo, hi);
        }
    } else {
        const unsigned full = 0xffffffffu;
        const int nmax = __reduce_max_sync(full, n > 0 ? n : 0);
        const int steps = (nmax + SVGT_STAGE_ROWS - 1) / SVGT_STAGE_ROWS;
        auto issue = [&](int s) {
            const uint32_t g = ring.step + s;
            const int st = g % kStages;
            int cnt = n - s * SVGT_STAGE_ROWS;
            cnt = cnt < 0 ? 0 : (cnt > SVGT_STAGE_ROWS ? SVGT_STAGE_ROWS : cnt);
            const uint32_t bytes = cnt * 32;
            const uint32_t total = __reduce_add_sync(full, bytes);
            if (lane == 0) mbar_expect_tx(ring.bars + st, total);
            __syncwarp();
            if (bytes)
                bulk_g2s(ring.slots + (st * 32 + lane) * kSlotBytes, rows + 2 * s * SVGT_STAGE_ROWS, bytes,
                         ring.bars + st);
        };
        for (int s = 0; s < kStages - 1 && s < steps; ++s) issue(s);
        for (int s = 0; s < steps; ++s) {
            if (s + kStages - 1 < steps) issue(s + kStages - 1);
            const uint32_t g = ring.step + s;
            const int st = g % kStages;
            mbar_wait(ring.bars + st, (g / kStages) & 1);
            const int4 *slot = reinterpret_cast<const int4 *>(ring.slots + (st * 32 + lane) * kSlotBytes);
            int cnt = n - s * SVGT_STAGE_ROWS;
            cnt = cnt > SVGT_STAGE_ROWS ? SVGT_STAGE_ROWS : cnt;
            for (int r = 0; r < cnt; ++r) f(slot[2 * r], slot[2 * r + 1]);
            __syncwarp();       /* every lane is done with the stage before it is refilled */
        }
        ring.step += steps;
    }
}


template <int VARIANT, int ASSOC>
__global__ void __launch_bounds__(SVGT_THREADS, 4) svgt_score_kernel(const SvgtParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    /* layout: pm[256] f64 | LibK[nl] | bars | (ring slots) | hist (optional) */
    double *s_pm = reinterpret_cast<double *>(smem_raw);
    LibK *s_lib = reinterpret_cast<LibK *>(s_pm + 256);
    const int nl = p.n_lib < SVGT_SMEM_LIBS ? p.n_lib : SVGT_SMEM_LIBS;
    size_t off = 256 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    uint64_t *s_bars = reinterpret_cast<uint64_t *>(smem_raw + off);
    off += SVGT_WARPS * kStages * sizeof(uint64_t);
    off = (off + 127) & ~(size_t)127;
    char *s_slots = reinterpret_cast<char *>(smem_raw + off);
    if (VARIANT == SVGT_VAR_BULK) off += (size_t)SVGT_WARPS * kStages * 32 * kSlotBytes;
    unsigned *s_hist = reinterpret_cast<unsigned *>(smem_raw + off);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int err = 0;
    for (int i = tid; i < 256; i += SVGT_THREADS) s_pm[i] = p.pm[i];
    for (int i = tid; i < nl; i += SVGT_THREADS) s_lib[i] = derive_lib(p, i, &err);
    if (p.hist_in_smem)
        for (int i = tid; i < (int)p.n_hist; i += SVGT_THREADS) s_hist[i] = p.hist[i];
    if (VARIANT == SVGT_VAR_BULK && tid < SVGT_WARPS * kStages) mbar_init(s_bars + tid, 1);
    if (VARIANT == SVGT_VAR_BULK) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    Tables t;
    t.pm = s_pm; t.libs = s_lib; t.hist = p.hist_in_smem ? s_hist : p.hist;
    t.conc = p.consts[C_CONC]; t.disc = p.consts[C_DISC];
    Ring ring;
    ring.slots = s_slots + (size_t)warp * kStages * 32 * kSlotBytes;
    ring.bars = s_bars + warp * kStages;
    ring.step = 0;

    const int m = p.min_aligned, slop = p.split_slop;
    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(p.status + 1, 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= p.n_tiles) break;
        const long long idx = (long long)tile * 32 + lane;
        const bool valid = idx < p.n_sites;
        long long site = 0;
        if (valid) site = p.order ? (long long)p.order[idx] : idx;
        int4 a = make_int4(0, 0, 0, 0), b = a, c = a, d = a;
        if (valid) {
            const int4 *sp = p.sites + site * 4;
            a = ldg4(sp); b = ldg4(sp + 1); c = ldg4(sp + 2); d = ldg4(sp + 3);
        }
        /* a = posA posB ciA0 ciA1 | b = ciB0 ciB1 tidA tidB | c = var_length meta foff.lo foff.hi | d = nf soff.lo soff.hi ns */
        const int meta = c.y;
        const bool skip = !valid || (meta & SITE_SKIP);
        const bool ranged = site_fields_in_range(a, b, m, slop);
        const bool run = !skip && ranged;
        const long long foff = ((long long)(unsigned)c.z) | ((long long)c.w << 32);
        const long long soff = ((long long)(unsigned)d.y) | ((long long)d.z << 32);
        int nf = run ? d.x : 0, ns = run ? d.w : 0;
        if (nf < 0 || foff < 0 || foff + nf > p.n_frag) { nf = 0; err = SVGT_ERR_ARG; }
        if (ns < 0 || soff < 0 || soff + ns > p.n_split) { ns = 0; err = SVGT_ERR_ARG; }

        SiteK s;
        s.posA = a.x; s.posB = a.y; s.ciA0 = a.z; s.ciA1 = a.w; s.ciB0 = b.x; s.ciB1 = b.y;
        s.tA = b.z; s.tB = b.w;
        s.var_length = c.x;
        s.LA = a.x + a.z - m; s.HA = a.x + a.w - m;
        s.LB = a.y + b.x + m + 1; s.HB = a.y + b.y + m + 1;
        s.wA0 = a.x - m; s.wA1 = a.x + m; s.wB0 = a.y - m; s.wB1 = a.y + m;
        s.dAB = a.y - a.x;
        s.meta = (meta & 15) | ((s.wA0 >= 0) << 8) | ((s.wB0 >= 0) << 9);

        Acc acc;
        acc.ref_seq = 0.0; acc.sub_ref = 0.0; acc.ref_span = 0.0; acc.alt_span = 0.0; acc.pend = 0; acc.err = 0;
        stream_rows<VARIANT>(ring, p.frags + foff * 2, nf, lane,
                             [&](const int4 lo, const int4 hi) { frag_row<ASSOC>(p, t, s, lo, hi, acc); });
        acc.ref_seq = __dadd_rn(acc.ref_seq, acc.sub_ref);

        /* arrange breakends left to right, parsers.py:1143-1161 */
        SplitK k;
        const int o1 = (meta >> 2) & 1, o2 = (meta >> 3) & 1;
        const bool swap = (s.tA != s.tB) || (s.posA > s.posB);
        k.tL = swap ? s.tB : s.tA; k.tR = swap ? s.tA : s.tB;
        const int pL = swap ? s.posB : s.posA, pR = swap ? s.posA : s.posB;
        k.rL = swap ? o2 : o1; k.rR = swap ? o1 : o2;
        k.loL = pL - slop; k.hiL = pL + slop; k.loR = pR - slop; k.hiR = pR + slop;
        k.svtype = meta & 3;
        SAcc sa;
        sa.alt_seq = 0.0; sa.alt_clip = 0.0; sa.sub_seq = 0.0; sa.sub_clip = 0.0;
        stream_rows<VARIANT>(ring, p.splits + soff * 2, ns, lane,
                             [&](const int4 q0, const int4 q1) { split_row<ASSOC>(t, k, q0, q1, sa); });
        sa.alt_seq = __dadd_rn(sa.alt_seq, sa.sub_seq);
        sa.alt_clip = __dadd_rn(sa.alt_clip, sa.sub_clip);

        if (valid) {
            svgt_out_row_t o;
            o.gl[0] = o.gl[1] = o.gl[2] = 0.0; o.sq = 0.0;
            o.gt = 0; o.gq = 0; o.dp = 0; o.ro = 0; o.ao = 0; o.qr = 0; o.qa = 0;
            o.rs = 0; o.as_ = 0; o.asc = 0; o.rp = 0; o.ap = 0;
            if (meta & SITE_SKIP) { o.gt = SVGT_GT_SKIPPED; o.gq = -1; }
            else if (!ranged) { o.gt = SVGT_GT_BLANK; o.gq = -1; err = SVGT_ERR_RANGE; }
            else {
                if (acc.err) err = acc.err;
                call_site(p, t, meta & 3, acc.ref_seq, sa.alt_seq, sa.alt_clip, acc.ref_span, acc.alt_span, o, err);
            }
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

template <int VARIANT>
cudaError_t launch_variant(const SvgtParams &p, cudaStream_t stream)
{
    auto kern = (p.assoc_mode == SVGT_ASSOC_CLASSIC) ? svgt_score_kernel<VARIANT, SVGT_ASSOC_CLASSIC>
                                                     : svgt_score_kernel<VARIANT, SVGT_ASSOC_SSO>;
    const size_t smem = svgt_score_smem_bytes(p, VARIANT);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SVGT_THREADS, smem)) != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = ((long long)p.n_tiles + SVGT_WARPS - 1) / SVGT_WARPS;
    long long cap = (long long)sms * per_sm;        /* persistent: a whole number of resident waves */
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, SVGT_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

size_t svgt_score_smem_bytes(const SvgtParams &p, int variant)
{
    size_t off = 256 * sizeof(double) + (size_t)SVGT_SMEM_LIBS * sizeof(LibK);
    off += SVGT_WARPS * kStages * sizeof(uint64_t);
    off = (off + 127) & ~(size_t)127;
    if (variant == SVGT_VAR_BULK) off += (size_t)SVGT_WARPS * kStages * 32 * kSlotBytes;
    if (p.hist_in_smem) off += (size_t)p.n_hist * sizeof(unsigned);
    return off;
}

int svgt_launch_score(const SvgtParams &p, int variant, cudaStream_t stream)
{
    if (variant == SVGT_VAR_BULK) return (int)launch_variant<SVGT_VAR_BULK>(p, stream);
    return (int)launch_variant<SVGT_VAR_DIRECT>(p, stream);
}
