/*
 * svgt_api.cu -- the C ABI of libsvgt.so (declared in include/svgt.h).
 *
 * svgt_score_compact() (the default path: compact 16-byte rows, svgt_compact.cu) and svgt_score_batch()
 * (wide 32-byte rows, the thread-per-site kernel of svgt_kernels.cu kept as an independent cross-check)
 * replace, for a whole batch of breakpoints, the per-breakpoint calls tally_variant_read_fragments +
 * bayesian_genotype of the reference (svtyper/singlesample.py:523-536, :486-498; inlined in
 * svtyper/classic.py:286-495).
 * There is no CPU fallback: without a CUDA device every compute entry point returns
 * SVGT_ERR_NO_DEVICE.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "svgt_kernels.cuh"

namespace {

thread_local char g_err[512] = "";
int g_variant = -1;   /* -1: default / environment */

int fail(int code, const char *fmt, const char *detail)
{
    snprintf(g_err, sizeof(g_err), fmt, detail ? detail : "");
    return code;
}

int cuda_fail(cudaError_t e, const char *where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return SVGT_ERR_CUDA;
}

int current_variant()
{
    if (g_variant >= 0) return g_variant;
    const char *env = getenv("SVGT_VARIANT");
    if (env && *env) {
        int v = atoi(env);
        if (v >= 0 && v < SVGT_VAR_COUNT) return v;
    }
    return SVGT_VAR_DIRECT;
}

int check_batch(const svgt_batch_t *b)
{
    if (!b) return fail(SVGT_ERR_ARG, "null batch%s", nullptr);
    if (b->n_sites < 0 || b->n_frag < 0 || b->n_split < 0 || b->n_lib < 0 || b->n_hist < 0 || b->n_log < 2)
        return fail(SVGT_ERR_ARG, "negative size in batch%s", nullptr);
    if (b->n_sites > 0 && !b->sites) return fail(SVGT_ERR_ARG, "null %s", "sites");
    if (b->n_frag > 0 && !b->frags) return fail(SVGT_ERR_ARG, "null %s", "frags");
    if (b->n_split > 0 && !b->splits) return fail(SVGT_ERR_ARG, "null %s", "splits");
    if (b->n_lib > 0 && (!b->lib_f64 || !b->lib_i32 || !b->hist)) return fail(SVGT_ERR_ARG, "null %s", "library tables");
    if (!b->pm || !b->logt || !b->consts) return fail(SVGT_ERR_ARG, "null %s", "look-up tables");
    if (b->assoc_mode != SVGT_ASSOC_SSO && b->assoc_mode != SVGT_ASSOC_CLASSIC)
        return fail(SVGT_ERR_ARG, "bad %s", "assoc_mode");
    if (b->n_sites / 32 >= 0x7fffffffLL) return fail(SVGT_ERR_ARG, "too many %s", "sites");
    const uintptr_t align = (uintptr_t)b->sites | (uintptr_t)b->frags | (uintptr_t)b->splits | (uintptr_t)b->lib_i32;
    if (align & 15) return fail(SVGT_ERR_ARG, "%s must be 16-byte aligned", "row arrays");
    return SVGT_OK;
}

}  // namespace

__global__ void svgt_wait_flags_kernel(const volatile int *flags, int n, int value)
{
    const int i = threadIdx.x;
    if (i < n) {
        while (flags[i] < value) __nanosleep(200);
    }
    __threadfence_system();
}

__global__ void svgt_set_flag_kernel(volatile int *flag, int value)
{
    __threadfence_system();
    *flag = value;
    __threadfence_system();
}

extern "C" {

int svgt_peer_copy(void *dst, const void *src, int64_t bytes, void *stream)
{
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(SVGT_ERR_ARG, "bad %s", "svgt_peer_copy arguments");
    if (bytes == 0) return SVGT_OK;
    cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream);
    return e == cudaSuccess ? SVGT_OK : cuda_fail(e, "cudaMemcpyAsync(peer)");
}

int svgt_set_flag(int32_t *flag, int32_t value, void *stream)
{
    if (!flag) return fail(SVGT_ERR_ARG, "null %s", "flag");
    svgt_set_flag_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SVGT_OK : cuda_fail(e, "svgt_set_flag launch");
}

int svgt_wait_flags(const int32_t *flags, int32_t n, int32_t value, void *stream)
{
    if (!flags || n < 0 || n > 1024) return fail(SVGT_ERR_ARG, "bad %s", "svgt_wait_flags arguments");
    if (n == 0) return SVGT_OK;
    svgt_wait_flags_kernel<<<1, ((n + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(flags, n, value);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SVGT_OK : cuda_fail(e, "svgt_wait_flags launch");
}

int svgt_abi_version(void) { return SVGT_ABI_VERSION; }

const char *svgt_last_error(void) { return g_err; }

int svgt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int svgt_set_variant(int variant)
{
    if (variant >= SVGT_VAR_COUNT) return fail(SVGT_ERR_ARG, "bad %s", "variant");
    g_variant = variant < 0 ? -1 : variant;
    return current_variant();
}

int svgt_launches_per_batch(const svgt_batch_t *batch)
{
    if (!batch) return fail(SVGT_ERR_ARG, "null batch%s", nullptr);
    if (batch->n_sites <= 0) return 0;
    return 1;                                               /* the thread-per-site kernel tallies and calls in one launch */
}

int svgt_score_batch(const svgt_batch_t *b, void *out_rows, int32_t *status, void *stream)
{
    int rc = check_batch(b);
    if (rc != SVGT_OK) return rc;
    if (!status || (b->n_sites > 0 && !out_rows)) return fail(SVGT_ERR_ARG, "null %s", "out_rows/status");
    if ((uintptr_t)out_rows & 15) return fail(SVGT_ERR_ARG, "%s must be 16-byte aligned", "out_rows");
    if (svgt_device_count() <= 0) return fail(SVGT_ERR_NO_DEVICE, "no CUDA device%s", nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(status)");
    if (b->n_sites == 0) return SVGT_OK;

    SvgtParams p;
    memset(&p, 0, sizeof(p));
    p.sites = (const int4 *)b->sites; p.n_sites = b->n_sites;
    p.frags = (const int4 *)b->frags; p.n_frag = b->n_frag;
    p.splits = (const int4 *)b->splits; p.n_split = b->n_split;
    p.order = b->order;
    p.lib_f64 = b->lib_f64; p.lib_i32 = (const int4 *)b->lib_i32; p.n_lib = b->n_lib;
    p.hist = b->hist; p.n_hist = b->n_hist;
    p.pm = b->pm; p.logt = b->logt; p.n_log = b->n_log; p.consts = b->consts;
    p.min_aligned = b->min_aligned; p.split_slop = b->split_slop; p.assoc_mode = b->assoc_mode;
    p.split_weight = b->split_weight; p.disc_weight = b->disc_weight;
    p.out = (svgt_out_row_t *)out_rows;
    p.status = status;
    p.n_tiles = (int)((b->n_sites + 31) / 32);
    p.hist_in_smem = (b->n_hist <= SVGT_SMEM_HIST_WORDS) ? 1 : 0;
    const int variant = current_variant();
    e = (cudaError_t)svgt_launch_score(p, variant, st);
    if (e != cudaSuccess) return cuda_fail(e, "svgt_score_kernel launch");
    return SVGT_OK;
}

/* ------------------------------------------------------------------------------------ */
/* host-buffer context                                                                   */
/* ------------------------------------------------------------------------------------ */
enum { kSlices = 8 };            /* pipelined host path: site slices in flight */

struct svgt_ctx {
    int device;
    cudaStream_t stream;         /* compute (and everything, in the one-shot path) */
    cudaStream_t s_h2d, s_d2h;   /* pipelined path: copy streams either side of the kernels */
    cudaEvent_t ev0, ev1;
    cudaEvent_t up[kSlices], k0[kSlices], k1[kSlices];
    void *buf[16];
    size_t cap[16];
    int32_t *plan_host;          /* entries | pieces | heavy of the last planned batch (host staging) */
    size_t plan_host_cap;
    int64_t h2d, d2h;
    int64_t planned_pieces;      /* pieces of the last svgt_ctx_score_host_compact call (0: scored without a plan) */
    float kernel_ms;
};

enum { B_SITES, B_FRAGS, B_SPLITS, B_ORDER, B_LIBF, B_LIBI, B_HIST, B_PM, B_LOG, B_CONSTS, B_OUT, B_STATUS,
       B_ENTRIES, B_PIECES, B_HEAVY, B_SCRATCH, B_COUNT };

static int ctx_reserve(svgt_ctx *c, int slot, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (c->cap[slot] >= bytes) return SVGT_OK;
    if (c->buf[slot]) cudaFree(c->buf[slot]);
    c->buf[slot] = nullptr; c->cap[slot] = 0;
    size_t want = bytes + bytes / 8;           /* a little slack so steady-state batches never regrow */
    cudaError_t e = cudaMalloc(&c->buf[slot], want);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    c->cap[slot] = want;
    return SVGT_OK;
}

static int ctx_upload(svgt_ctx *c, int slot, const void *src, size_t bytes)
{
    int rc = ctx_reserve(c, slot, bytes);
    if (rc != SVGT_OK) return rc;
    if (bytes == 0) return SVGT_OK;
    cudaError_t e = cudaMemcpyAsync(c->buf[slot], src, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(H2D)");
    c->h2d += (int64_t)bytes;
    return SVGT_OK;
}

int svgt_ctx_create(int device, svgt_ctx_t **out)
{
    if (!out) return fail(SVGT_ERR_ARG, "null %s", "ctx pointer");
    *out = nullptr;
    int n = svgt_device_count();
    if (n <= 0) return fail(SVGT_ERR_NO_DEVICE, "no CUDA device%s", nullptr);
    if (device < 0 || device >= n) return fail(SVGT_ERR_ARG, "bad %s", "device index");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    svgt_ctx *c = (svgt_ctx *)calloc(1, sizeof(svgt_ctx));
    if (!c) return fail(SVGT_ERR_ARG, "out of %s", "host memory");
    c->device = device;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { free(c); return cuda_fail(e, "cudaStreamCreate"); }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking);
    for (int i = 0; i < kSlices; ++i) {
        cudaEventCreateWithFlags(&c->up[i], cudaEventDisableTiming);
        cudaEventCreate(&c->k0[i]);
        cudaEventCreate(&c->k1[i]);
    }
    *out = c;
    return SVGT_OK;
}

int svgt_ctx_destroy(svgt_ctx_t *c)
{
    if (!c) return SVGT_OK;
    cudaSetDevice(c->device);
    for (int i = 0; i < B_COUNT; ++i)
        if (c->buf[i]) cudaFree(c->buf[i]);
    if (c->plan_host) cudaFreeHost(c->plan_host);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    for (int i = 0; i < kSlices; ++i) { cudaEventDestroy(c->up[i]); cudaEventDestroy(c->k0[i]); cudaEventDestroy(c->k1[i]); }
    cudaStreamDestroy(c->s_h2d);
    cudaStreamDestroy(c->s_d2h);
    cudaStreamDestroy(c->stream);
    free(c);
    return SVGT_OK;
}

int svgt_ctx_score_host(svgt_ctx_t *c, const svgt_batch_t *hb, void *out_rows_host)
{
    if (!c) return fail(SVGT_ERR_ARG, "null %s", "ctx");
    int rc = check_batch(hb);
    if (rc != SVGT_OK) return rc;
    if (hb->n_sites > 0 && !out_rows_host) return fail(SVGT_ERR_ARG, "null %s", "out_rows_host");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    c->h2d = c->d2h = 0;
    c->kernel_ms = 0.f;

    svgt_batch_t db = *hb;
#define UP(slot, field, type, count)                                                          \
    do {                                                                                      \
        rc = ctx_upload(c, slot, hb->field, (size_t)(count) * sizeof(type));                  \
        if (rc != SVGT_OK) return rc;                                                         \
        db.field = (const type *)c->buf[slot];                                                \
    } while (0)
    UP(B_SITES, sites, int32_t, hb->n_sites * SVGT_SITE_WORDS);
    UP(B_FRAGS, frags, int32_t, hb->n_frag * SVGT_FRAG_WORDS);
    UP(B_SPLITS, splits, int32_t, hb->n_split * SVGT_SPLIT_WORDS);
    if (hb->order) UP(B_ORDER, order, int32_t, hb->n_sites);
    UP(B_LIBF, lib_f64, double, hb->n_lib * 4);
    UP(B_LIBI, lib_i32, int32_t, hb->n_lib * 4);
    UP(B_HIST, hist, uint32_t, hb->n_hist);
    UP(B_PM, pm, double, 256);
    UP(B_LOG, logt, double, hb->n_log);
    UP(B_CONSTS, consts, double, 32);
#undef UP
    if ((rc = ctx_reserve(c, B_OUT, (size_t)hb->n_sites * SVGT_OUT_BYTES)) != SVGT_OK) return rc;
    if ((rc = ctx_reserve(c, B_STATUS, 16)) != SVGT_OK) return rc;

    cudaEventRecord(c->ev0, c->stream);
    rc = svgt_score_batch(&db, c->buf[B_OUT], (int32_t *)c->buf[B_STATUS], c->stream);
    if (rc != SVGT_OK) return rc;
    cudaEventRecord(c->ev1, c->stream);

    int32_t status[4] = {0, 0, 0, 0};
    if (hb->n_sites > 0) {
        e = cudaMemcpyAsync(out_rows_host, c->buf[B_OUT], (size_t)hb->n_sites * SVGT_OUT_BYTES,
                            cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(D2H)");
        c->d2h += hb->n_sites * (int64_t)SVGT_OUT_BYTES;
    }
    e = cudaMemcpyAsync(status, c->buf[B_STATUS], sizeof(status), cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(status)");
    c->d2h += (int64_t)sizeof(status);
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    cudaEventElapsedTime(&c->kernel_ms, c->ev0, c->ev1);
    if (status[0] != 0) {
        char msg[64];
        snprintf(msg, sizeof(msg), "%d site(s), first code %d", status[2], status[0]);
        return fail(status[0], "scoring kernel flagged %s", msg);
    }
    return SVGT_OK;
}


/* ------------------------------------------------------------------------------------ */
/* compact schema (the default path)                                                     */
/* ------------------------------------------------------------------------------------ */
static int check_cbatch(const svgt_cbatch_t *b)
{
    if (!b) return fail(SVGT_ERR_ARG, "null batch%s", nullptr);
    if (b->n_sites < 0 || b->n_rows < 0 || b->n_lib < 0 || b->n_hist < 0 || b->n_log < 2)
        return fail(SVGT_ERR_ARG, "negative size in batch%s", nullptr);
    if (b->n_sites > 0 && !b->sites) return fail(SVGT_ERR_ARG, "null %s", "sites");
    if (b->n_rows > 0 && !b->rows) return fail(SVGT_ERR_ARG, "null %s", "rows");
    if (b->n_lib > 0 && (!b->lib_f64 || !b->lib_i32 || !b->hist)) return fail(SVGT_ERR_ARG, "null %s", "library tables");
    if (!b->pm || !b->logt || !b->consts) return fail(SVGT_ERR_ARG, "null %s", "look-up tables");
    if (b->assoc_mode != SVGT_ASSOC_SSO && b->assoc_mode != SVGT_ASSOC_CLASSIC)
        return fail(SVGT_ERR_ARG, "bad %s", "assoc_mode");
    if (b->unit_mode < 0 || b->unit_mode > 3) return fail(SVGT_ERR_ARG, "bad %s", "unit_mode");
    if (b->rows_min_aligned != b->min_aligned)
        return fail(SVGT_ERR_ARG, "%s: the rows were packed for another min_aligned", "rows_min_aligned");
    if (b->n_sites / 32 >= 0x7fffffffLL) return fail(SVGT_ERR_ARG, "too many %s", "sites");
    const uintptr_t align = (uintptr_t)b->sites | (uintptr_t)b->rows | (uintptr_t)b->lib_i32 | (uintptr_t)b->out_final;
    if (align & 15) return fail(SVGT_ERR_ARG, "%s must be 16-byte aligned", "row arrays");
    return SVGT_OK;
}

int svgt_score_compact(const svgt_cbatch_t *b, void *out_rows, int32_t *status, void *stream)
{
    int rc = check_cbatch(b);
    if (rc != SVGT_OK) return rc;
    if (!status || (b->n_sites > 0 && !out_rows)) return fail(SVGT_ERR_ARG, "null %s", "out_rows/status");
    if ((uintptr_t)out_rows & 15) return fail(SVGT_ERR_ARG, "%s must be 16-byte aligned", "out_rows");
    if (svgt_device_count() <= 0) return fail(SVGT_ERR_NO_DEVICE, "no CUDA device%s", nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(status)");
    if (b->n_sites == 0 && !b->done_flag) return SVGT_OK;

    SvgtCompactParams cp;
    memset(&cp, 0, sizeof(cp));
    SvgtParams &p = cp.base;
    p.n_sites = b->n_sites;
    p.order = b->order;
    p.lib_f64 = b->lib_f64; p.lib_i32 = (const int4 *)b->lib_i32; p.n_lib = b->n_lib;
    p.hist = b->hist; p.n_hist = b->n_hist;
    p.pm = b->pm; p.logt = b->logt; p.n_log = b->n_log; p.consts = b->consts;
    p.min_aligned = b->min_aligned; p.split_slop = b->split_slop; p.assoc_mode = b->assoc_mode;
    p.split_weight = b->split_weight; p.disc_weight = b->disc_weight;
    p.out = (svgt_out_row_t *)out_rows;
    p.status = status;
    cp.sites = (const int4 *)b->sites;
    cp.rows = (const int4 *)b->rows; cp.n_rows = b->n_rows;
    cp.out_final = (svgt_out_row_t *)b->out_final;
    cp.done_flag = b->done_flag; cp.done_value = b->done_value;
    cp.hist_max = b->hist_max;
    if (b->plan && b->plan->n_heavy > 0) {
        const svgt_segplan_t *pl = b->plan;
        if (pl->n_entries < 0 || pl->n_pieces < 0 || pl->scratch_chunks < 0 || !pl->entries || !pl->pieces || !pl->heavy ||
            !pl->scratch || ((uintptr_t)pl->scratch & 15) || ((uintptr_t)pl->pieces & 15) || ((uintptr_t)pl->heavy & 15))
            return fail(SVGT_ERR_ARG, "bad %s", "piece plan");
        cp.entries = pl->entries; cp.n_entries = pl->n_entries;
        cp.pieces = (const int4 *)pl->pieces; cp.n_pieces = pl->n_pieces;
        cp.heavy = (const int4 *)pl->heavy; cp.n_heavy = pl->n_heavy;
        cp.scratch = (double *)pl->scratch;
        cp.scratch_lead = (int *)((char *)pl->scratch + (size_t)pl->scratch_chunks * 768);
        cp.scratch_chunks = pl->scratch_chunks;
    }
    e = (cudaError_t)svgt_launch_compact(cp, b->unit_mode, st);
    if (e != cudaSuccess) return cuda_fail(e, "svgt_compact_kernel launch");
    return SVGT_OK;
}

/* ------------------------------------------------------------------------------------ */
/* piece plan (svgt_segplan_t)                                                           */
/* ------------------------------------------------------------------------------------ */
namespace {

constexpr int kPlanMinChunks = 4;       /* shortest piece: below this the per-piece set-up outweighs the chunks */
constexpr int kPlanShare = 14;          /* a piece may take 1/kPlanShare of a warp's fair share of the batch's chunks
                                           (a lone piece advances ~3.5x slower than a warp among 20 busy ones: the
                                           longest piece then costs about a quarter of the batch's run time) */
constexpr int kPlanWarpsPerSm = 20;     /* SVGT_C_THREADS / 32 */

/* rows of a site as the tally kernel will see them: SKIP sites and sites it refuses (svgt_device.cuh:
 * site_fields_in_range, malformed counts) have none */
inline void plan_site_chunks(const int32_t *row, int m, int slop, int &cf, int &cs)
{
    auto ok = [](int v, int lim) { return v > -lim && v < lim; };
    const bool ranged = ok(row[0], 1 << 30) && ok(row[1], 1 << 30) && ok(row[2], 1 << 28) && ok(row[3], 1 << 28) &&
                        ok(row[4], 1 << 28) && ok(row[5], 1 << 28) && m >= 0 && m < (1 << 20) && slop >= 0 && slop < (1 << 20);
    const int nf = row[10], ns = row[11];
    if ((row[7] & (1 << 4)) || !ranged || nf < 0 || ns < 0) { cf = 0; cs = 0; return; }
    cf = (int)(((int64_t)nf + 31) >> 5); cs = (int)(((int64_t)ns + 31) >> 5);
}

int plan_resident_warps()
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        sms = 148;
    }
    return sms * kPlanWarpsPerSm;
}

}  // namespace

extern "C" int svgt_plan_count(const int32_t *sites, int64_t n_sites, int32_t min_aligned, int32_t split_slop,
                               int32_t resident_warps, int32_t force_chunks, int32_t *max_chunks, int64_t *n_entries,
                               int64_t *n_pieces, int64_t *n_heavy, int64_t *scratch_chunks)
{
    if (n_sites < 0 || (n_sites > 0 && !sites) || !max_chunks || !n_entries || !n_pieces || !n_heavy || !scratch_chunks)
        return fail(SVGT_ERR_ARG, "bad %s", "svgt_plan_count arguments");
    if (n_sites / 32 >= 0x7fffffffLL) return fail(SVGT_ERR_ARG, "too many %s", "sites");
    *n_entries = *n_pieces = *n_heavy = *scratch_chunks = 0;
    int64_t total = 0;
    int longest = 0;
    for (int64_t i = 0; i < n_sites; ++i) {
        int cf, cs;
        plan_site_chunks(sites + i * SVGT_CSITE_WORDS, min_aligned, split_slop, cf, cs);
        total += (int64_t)cf + cs;
        if (cf + cs > longest) longest = cf + cs;
    }
    int64_t L = force_chunks;
    if (L <= 0) {
        const int64_t warps = resident_warps > 0 ? resident_warps : plan_resident_warps();
        static const long long min_chunks = [] {
            const char *v = getenv("SVGT_PLAN_MIN_CHUNKS");
            return v && *v && atoll(v) > 0 ? atoll(v) : (long long)kPlanMinChunks;
        }();
        L = (total + kPlanShare * warps - 1) / (kPlanShare * warps);
        if (L < min_chunks) L = min_chunks;
    }
    if (L > 0x3ffffff) L = 0x3ffffff;
    *max_chunks = (int32_t)L;
    if (longest <= L) return SVGT_OK;                       /* nothing to cut */
    int64_t ne = 0, np = 0, nh = 0, sc = 0;
    for (int64_t i = 0; i < n_sites; ++i) {
        int cf, cs;
        plan_site_chunks(sites + i * SVGT_CSITE_WORDS, min_aligned, split_slop, cf, cs);
        if (cf + cs == 0) continue;
        if (cf + cs <= L) { ++ne; continue; }
        const int64_t k = (cf + L - 1) / L + (cs + L - 1) / L;
        ++nh; np += k; ne += k; sc += (int64_t)cf + cs;
    }
    if (np >= 0x7fffffffLL || sc >= 0x7fffffffLL || ne >= 0x7fffffffLL) return fail(SVGT_ERR_ARG, "too many %s", "pieces");
    *n_entries = ne; *n_pieces = np; *n_heavy = nh; *scratch_chunks = sc;
    return SVGT_OK;
}

extern "C" int svgt_plan_fill(const int32_t *sites, int64_t n_sites, int32_t min_aligned, int32_t split_slop,
                              int32_t max_chunks, int32_t *entries, int32_t *pieces, int32_t *heavy)
{
    if (n_sites < 0 || (n_sites > 0 && !sites) || max_chunks <= 0 || !entries || !pieces || !heavy)
        return fail(SVGT_ERR_ARG, "bad %s", "svgt_plan_fill arguments");
    const int64_t L = max_chunks;
    /* counting sort of the entries by chunk count, heaviest first (an entry has 1..L chunks) */
    int64_t *start = (int64_t *)calloc((size_t)L + 2, sizeof(int64_t));
    if (!start) return fail(SVGT_ERR_ARG, "out of %s", "host memory");
    for (int64_t i = 0; i < n_sites; ++i) {
        int cf, cs;
        plan_site_chunks(sites + i * SVGT_CSITE_WORDS, min_aligned, split_slop, cf, cs);
        if (cf + cs == 0) continue;
        if (cf + cs <= L) { ++start[cf + cs]; continue; }
        for (int part = 0; part < 2; ++part)
            for (int64_t c0 = 0, n = part ? cs : cf; c0 < n; c0 += L) ++start[n - c0 < L ? n - c0 : L];
    }
    {
        int64_t at = 0;
        for (int64_t w = L; w >= 1; --w) { const int64_t n = start[w]; start[w] = at; at += n; }
    }
    int64_t np = 0, nh = 0, sc = 0;
    for (int64_t i = 0; i < n_sites; ++i) {
        const int32_t *row = sites + i * SVGT_CSITE_WORDS;
        int cf, cs;
        plan_site_chunks(row, min_aligned, split_slop, cf, cs);
        if (cf + cs == 0) continue;
        if (cf + cs <= L) { entries[start[cf + cs]++] = (int32_t)i; continue; }
        int32_t *h = heavy + nh * 4;
        h[0] = (int32_t)i; h[1] = (int32_t)sc; h[2] = cf; h[3] = cs;
        ++nh;
        for (int part = 0; part < 2; ++part) {
            const int64_t n = part ? cs : cf, rows = part ? row[11] : row[10];
            for (int64_t c0 = 0; c0 < n; c0 += L) {
                const int64_t w = n - c0 < L ? n - c0 : L;
                const int64_t r0 = c0 * 32, cnt = rows - r0 < w * 32 ? rows - r0 : w * 32;
                int32_t *q = pieces + np * 4;
                q[0] = (int32_t)i; q[1] = (int32_t)r0;
                q[2] = (int32_t)((uint32_t)cnt | (part ? 0x80000000u : 0u));
                q[3] = (int32_t)(sc + (part ? cf : 0) + c0);
                entries[start[w]++] = (int32_t)~np;
                ++np;
            }
        }
        sc += (int64_t)cf + cs;
    }
    free(start);
    return SVGT_OK;
}

static void ctx_sync_all(svgt_ctx *c)
{
    cudaStreamSynchronize(c->s_h2d);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->s_d2h);
}

static int64_t csite_off(const int32_t *row) { return (int64_t)(uint32_t)row[8] | ((int64_t)row[9] << 32); }

/*
 * Pipelined host path for large compact batches whose rows are laid out in site order
 * (SVGT_LAYOUT_SITE_ORDER): the sites are cut into kSlices contiguous slices; slice k's site rows and
 * evidence rows go up on the H2D stream while slice k-1 is scored and slice k-2's result rows come back on
 * the D2H stream.  Slice boundaries are validated on the host before anything is queued; every error exit
 * drains the three streams first, so no copy is still reading or writing the caller's buffers when the
 * call returns.
 */
static int ctx_score_pipelined_compact(svgt_ctx *c, const svgt_cbatch_t *hb, void *out_rows_host)
{
    const int64_t n = hb->n_sites;
    int64_t cut_s[kSlices + 1], cut_r[kSlices + 1];
    for (int k = 0; k <= kSlices; ++k) {
        cut_s[k] = n * k / kSlices;
        cut_r[k] = cut_s[k] < n ? csite_off(hb->sites + cut_s[k] * SVGT_CSITE_WORDS) : hb->n_rows;
        if (cut_r[k] < 0 || cut_r[k] > hb->n_rows || (k > 0 && cut_r[k] < cut_r[k - 1]))
            return fail(SVGT_ERR_ARG, "%s: row offsets are not in site order", "SVGT_LAYOUT_SITE_ORDER");
    }
    if (cut_r[0] != 0) return fail(SVGT_ERR_ARG, "%s: the first site's rows do not start at 0", "SVGT_LAYOUT_SITE_ORDER");
    int rc = SVGT_OK;
    cudaError_t e = cudaSuccess;
    svgt_cbatch_t db = *hb;
#define UPS(slot, field, type, count)                                                         \
    do {                                                                                      \
        const size_t bytes_ = (size_t)(count) * sizeof(type);                                 \
        if ((rc = ctx_reserve(c, slot, bytes_)) != SVGT_OK) goto fail_out;                    \
        if (bytes_) {                                                                         \
            e = cudaMemcpyAsync(c->buf[slot], hb->field, bytes_, cudaMemcpyHostToDevice, c->s_h2d); \
            if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(H2D)"); goto fail_out; } \
            c->h2d += (int64_t)bytes_;                                                        \
        }                                                                                     \
        db.field = (const type *)c->buf[slot];                                                \
    } while (0)
    UPS(B_LIBF, lib_f64, double, hb->n_lib * 4);
    UPS(B_LIBI, lib_i32, int32_t, hb->n_lib * 4);
    UPS(B_HIST, hist, uint32_t, hb->n_hist);
    UPS(B_PM, pm, double, 256);
    UPS(B_LOG, logt, double, hb->n_log);
    UPS(B_CONSTS, consts, double, 32);
#undef UPS
    if ((rc = ctx_reserve(c, B_SITES, (size_t)n * SVGT_CSITE_WORDS * 4)) != SVGT_OK) goto fail_out;
    if ((rc = ctx_reserve(c, B_FRAGS, (size_t)hb->n_rows * SVGT_CROW_WORDS * 4)) != SVGT_OK) goto fail_out;
    if ((rc = ctx_reserve(c, B_OUT, (size_t)n * SVGT_OUT_BYTES)) != SVGT_OK) goto fail_out;
    if ((rc = ctx_reserve(c, B_STATUS, 16 * kSlices)) != SVGT_OK) goto fail_out;
    db.order = nullptr;                               /* the launch permutation spans the whole batch */
    db.plan = nullptr;                                /* slices of >= 16k sites each: throughput-bound, no pieces */
    c->planned_pieces = 0;
    db.rows = (const int32_t *)c->buf[B_FRAGS];
    for (int k = 0; k < kSlices; ++k) {
        const int64_t s0 = cut_s[k], s1 = cut_s[k + 1];
        const size_t bs = (size_t)(s1 - s0) * SVGT_CSITE_WORDS * 4, br = (size_t)(cut_r[k + 1] - cut_r[k]) * SVGT_CROW_WORDS * 4;
        if (bs && (e = cudaMemcpyAsync((char *)c->buf[B_SITES] + (size_t)s0 * SVGT_CSITE_WORDS * 4,
                                       hb->sites + s0 * SVGT_CSITE_WORDS, bs, cudaMemcpyHostToDevice, c->s_h2d)) != cudaSuccess) {
            rc = cuda_fail(e, "cudaMemcpyAsync(H2D)"); goto fail_out;
        }
        if (br && (e = cudaMemcpyAsync((char *)c->buf[B_FRAGS] + (size_t)cut_r[k] * SVGT_CROW_WORDS * 4,
                                       hb->rows + cut_r[k] * SVGT_CROW_WORDS, br, cudaMemcpyHostToDevice, c->s_h2d)) != cudaSuccess) {
            rc = cuda_fail(e, "cudaMemcpyAsync(H2D)"); goto fail_out;
        }
        c->h2d += (int64_t)(bs + br);
        cudaEventRecord(c->up[k], c->s_h2d);
        if (s1 == s0) continue;
        cudaStreamWaitEvent(c->stream, c->up[k], 0);
        svgt_cbatch_t sb = db;
        sb.sites = (const int32_t *)c->buf[B_SITES] + s0 * SVGT_CSITE_WORDS;
        sb.n_sites = s1 - s0; sb.n_rows = cut_r[k + 1];  /* a site whose rows lie beyond the uploaded prefix is a real error */
        cudaEventRecord(c->k0[k], c->stream);
        rc = svgt_score_compact(&sb, (char *)c->buf[B_OUT] + (size_t)s0 * SVGT_OUT_BYTES, (int32_t *)c->buf[B_STATUS] + 4 * k,
                                c->stream);
        if (rc != SVGT_OK) goto fail_out;
        cudaEventRecord(c->k1[k], c->stream);
        cudaStreamWaitEvent(c->s_d2h, c->k1[k], 0);
        e = cudaMemcpyAsync((char *)out_rows_host + (size_t)s0 * SVGT_OUT_BYTES, (char *)c->buf[B_OUT] + (size_t)s0 * SVGT_OUT_BYTES,
                            (size_t)(s1 - s0) * SVGT_OUT_BYTES, cudaMemcpyDeviceToHost, c->s_d2h);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(D2H)"); goto fail_out; }
        c->d2h += (s1 - s0) * (int64_t)SVGT_OUT_BYTES;
    }
    {
        int32_t status[4 * kSlices];
        memset(status, 0, sizeof(status));
        cudaStreamWaitEvent(c->s_d2h, c->k1[kSlices - 1], 0);
        e = cudaMemcpyAsync(status, c->buf[B_STATUS], sizeof(status), cudaMemcpyDeviceToHost, c->s_d2h);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(status)"); goto fail_out; }
        c->d2h += (int64_t)sizeof(status);
        if ((e = cudaStreamSynchronize(c->s_d2h)) != cudaSuccess) { rc = cuda_fail(e, "cudaStreamSynchronize"); goto fail_out; }
        if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) { rc = cuda_fail(e, "cudaStreamSynchronize"); goto fail_out; }
        c->kernel_ms = 0.f;
        int first = 0, count = 0;
        for (int k = 0; k < kSlices; ++k) {
            if (cut_s[k + 1] == cut_s[k]) continue;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->k0[k], c->k1[k]);
            c->kernel_ms += ms;
            if (status[4 * k] != 0 && first == 0) first = status[4 * k];
            count += status[4 * k + 2];
        }
        if (first != 0) {
            char msg[64];
            snprintf(msg, sizeof(msg), "%d site(s), first code %d", count, first);
            return fail(first, "scoring kernel flagged %s", msg);
        }
    }
    return SVGT_OK;
fail_out:
    ctx_sync_all(c);
    return rc;
}

int svgt_ctx_score_host_compact(svgt_ctx_t *c, const svgt_cbatch_t *hb, void *out_rows_host)
{
    if (!c) return fail(SVGT_ERR_ARG, "null %s", "ctx");
    int rc = check_cbatch(hb);
    if (rc != SVGT_OK) return rc;
    if (hb->out_final || hb->done_flag) return fail(SVGT_ERR_ARG, "%s must be NULL on the host path", "out_final/done_flag");
    if (hb->n_sites > 0 && !out_rows_host) return fail(SVGT_ERR_ARG, "null %s", "out_rows_host");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    c->h2d = c->d2h = 0;
    c->kernel_ms = 0.f;

    /* large batches in site order: overlap the copies with the kernels (SVGT_PIPELINE_MIN_SITES overrides
     * the threshold, 0 disables) */
    static const long long min_sites = [] {
        const char *v = getenv("SVGT_PIPELINE_MIN_SITES");
        return v && *v ? atoll(v) : 131072LL;
    }();
    if (min_sites > 0 && hb->n_sites >= min_sites && hb->n_sites >= kSlices && (hb->flags & SVGT_LAYOUT_SITE_ORDER))
        return ctx_score_pipelined_compact(c, hb, out_rows_host);

    svgt_cbatch_t db = *hb;
#define UP(slot, field, type, count)                                                          \
    do {                                                                                      \
        rc = ctx_upload(c, slot, hb->field, (size_t)(count) * sizeof(type));                  \
        if (rc != SVGT_OK) { cudaStreamSynchronize(c->stream); return rc; }                   \
        db.field = (const type *)c->buf[slot];                                                \
    } while (0)
    UP(B_SITES, sites, int32_t, hb->n_sites * SVGT_CSITE_WORDS);
    UP(B_FRAGS, rows, int32_t, hb->n_rows * SVGT_CROW_WORDS);
    if (hb->order) UP(B_ORDER, order, int32_t, hb->n_sites);
    UP(B_LIBF, lib_f64, double, hb->n_lib * 4);
    UP(B_LIBI, lib_i32, int32_t, hb->n_lib * 4);
    UP(B_HIST, hist, uint32_t, hb->n_hist);
    UP(B_PM, pm, double, 256);
    UP(B_LOG, logt, double, hb->n_log);
    UP(B_CONSTS, consts, double, 32);
#undef UP
    if ((rc = ctx_reserve(c, B_OUT, (size_t)hb->n_sites * SVGT_OUT_BYTES)) != SVGT_OK) { cudaStreamSynchronize(c->stream); return rc; }
    if ((rc = ctx_reserve(c, B_STATUS, 16)) != SVGT_OK) { cudaStreamSynchronize(c->stream); return rc; }

    /* the host has the site rows, so it can cut sites that are too long for one warp into pieces (svgt_segplan_t).
     * Off by default -- measured, a plan does not pay on the shapes of BASELINE.json (profiles/README.md): the serial
     * part left, one dependent fp64 add per row, is most of what a long site costs anyway.  SVGT_PLAN=1 plans with
     * svgt_plan_count's policy, SVGT_PLAN_FORCE_CHUNKS=k with pieces of k chunks (tests) */
    svgt_segplan_t plan;
    memset(&plan, 0, sizeof(plan));
    db.plan = nullptr;
    static const long long plan_mode = [] {
        const char *v = getenv("SVGT_PLAN");
        return v && *v ? atoll(v) : 0LL;
    }();
    const char *fv = getenv("SVGT_PLAN_FORCE_CHUNKS");
    const int32_t force = fv && *fv ? (int32_t)atoi(fv) : 0;
    if ((plan_mode != 0 || force > 0) && hb->n_sites > 0 && hb->unit_mode != 2) {
        int32_t L = 0;
        int64_t ne = 0, np = 0, nh = 0, sc = 0;
        rc = svgt_plan_count(hb->sites, hb->n_sites, hb->min_aligned, hb->split_slop, 0, force, &L, &ne, &np, &nh, &sc);
        if (rc != SVGT_OK) { cudaStreamSynchronize(c->stream); return rc; }
        if (nh > 0) {
            const size_t words = (size_t)ne + (size_t)np * 4 + (size_t)nh * 4 + 8;
            if (c->plan_host_cap < words) {
                /* earlier plans may still be in flight from this staging buffer only within a call that has returned */
                if (c->plan_host) cudaFreeHost(c->plan_host);
                c->plan_host = nullptr; c->plan_host_cap = 0;
                if ((e = cudaMallocHost((void **)&c->plan_host, (words + words / 4) * sizeof(int32_t))) != cudaSuccess) {
                    cudaStreamSynchronize(c->stream);
                    return cuda_fail(e, "cudaMallocHost(plan)");
                }
                c->plan_host_cap = words + words / 4;
            }
            int32_t *h_entries = c->plan_host;
            int32_t *h_pieces = h_entries + (((size_t)ne + 3) & ~(size_t)3);
            int32_t *h_heavy = h_pieces + (size_t)np * 4;
            rc = svgt_plan_fill(hb->sites, hb->n_sites, hb->min_aligned, hb->split_slop, L, h_entries, h_pieces, h_heavy);
            if (rc == SVGT_OK) rc = ctx_upload(c, B_ENTRIES, h_entries, (size_t)ne * 4);
            if (rc == SVGT_OK) rc = ctx_upload(c, B_PIECES, h_pieces, (size_t)np * 16);
            if (rc == SVGT_OK) rc = ctx_upload(c, B_HEAVY, h_heavy, (size_t)nh * 16);
            if (rc == SVGT_OK) rc = ctx_reserve(c, B_SCRATCH, (size_t)sc * SVGT_PLAN_CHUNK_BYTES);
            if (rc != SVGT_OK) { cudaStreamSynchronize(c->stream); return rc; }
            plan.entries = (const int32_t *)c->buf[B_ENTRIES]; plan.n_entries = ne;
            plan.pieces = (const int32_t *)c->buf[B_PIECES]; plan.n_pieces = np;
            plan.heavy = (const int32_t *)c->buf[B_HEAVY]; plan.n_heavy = nh;
            plan.scratch = c->buf[B_SCRATCH]; plan.scratch_chunks = sc;
            db.plan = &plan;
            db.unit_mode = 3;                               /* heaviest entries first, one per warp */
        }
    }
    c->planned_pieces = db.plan ? plan.n_pieces : 0;

    cudaEventRecord(c->ev0, c->stream);
    rc = svgt_score_compact(&db, c->buf[B_OUT], (int32_t *)c->buf[B_STATUS], c->stream);
    if (rc != SVGT_OK) { cudaStreamSynchronize(c->stream); return rc; }
    cudaEventRecord(c->ev1, c->stream);

    int32_t status[4] = {0, 0, 0, 0};
    if (hb->n_sites > 0) {
        e = cudaMemcpyAsync(out_rows_host, c->buf[B_OUT], (size_t)hb->n_sites * SVGT_OUT_BYTES,
                            cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) { cudaStreamSynchronize(c->stream); return cuda_fail(e, "cudaMemcpyAsync(D2H)"); }
        c->d2h += hb->n_sites * (int64_t)SVGT_OUT_BYTES;
    }
    e = cudaMemcpyAsync(status, c->buf[B_STATUS], sizeof(status), cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) { cudaStreamSynchronize(c->stream); return cuda_fail(e, "cudaMemcpyAsync(status)"); }
    c->d2h += (int64_t)sizeof(status);
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    cudaEventElapsedTime(&c->kernel_ms, c->ev0, c->ev1);
    if (status[0] != 0) {
        char msg[64];
        snprintf(msg, sizeof(msg), "%d site(s), first code %d", status[2], status[0]);
        return fail(status[0], "scoring kernel flagged %s", msg);
    }
    return SVGT_OK;
}

/* ------------------------------------------------------------------------------------ */
/* peer-visible buffers (multi-GPU output without a collective)                          */
/* ------------------------------------------------------------------------------------ */
int svgt_shared_alloc(int64_t bytes, void **dev_ptr, unsigned char handle[64])
{
    if (!dev_ptr || !handle || bytes <= 0) return fail(SVGT_ERR_ARG, "bad %s", "svgt_shared_alloc arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void *ptr = nullptr;
    cudaError_t e = cudaMalloc(&ptr, (size_t)bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(shared)");
    if ((e = cudaMemset(ptr, 0, (size_t)bytes)) != cudaSuccess) { cudaFree(ptr); return cuda_fail(e, "cudaMemset(shared)"); }
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, ptr)) != cudaSuccess) { cudaFree(ptr); return cuda_fail(e, "cudaIpcGetMemHandle"); }
    memcpy(handle, &h, 64);
    *dev_ptr = ptr;
    return SVGT_OK;
}

int svgt_shared_open(const unsigned char handle[64], void **dev_ptr)
{
    if (!dev_ptr || !handle) return fail(SVGT_ERR_ARG, "bad %s", "svgt_shared_open arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
    *dev_ptr = ptr;
    return SVGT_OK;
}

int svgt_shared_close(void *dev_ptr)
{
    if (!dev_ptr) return SVGT_OK;
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    return e == cudaSuccess ? SVGT_OK : cuda_fail(e, "cudaIpcCloseMemHandle");
}

int svgt_memcpy_d2h(void *dst_host, const void *src_dev, int64_t bytes)
{
    if (bytes < 0 || (bytes > 0 && (!dst_host || !src_dev))) return fail(SVGT_ERR_ARG, "bad %s", "svgt_memcpy_d2h arguments");
    if (bytes == 0) return SVGT_OK;
    cudaError_t e = cudaMemcpy(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? SVGT_OK : cuda_fail(e, "cudaMemcpy(D2H)");
}

int svgt_shared_free(void *dev_ptr)
{
    if (!dev_ptr) return SVGT_OK;
    cudaError_t e = cudaFree(dev_ptr);
    return e == cudaSuccess ? SVGT_OK : cuda_fail(e, "cudaFree(shared)");
}

int svgt_ctx_last_traffic(const svgt_ctx_t *c, int64_t *h2d_bytes, int64_t *d2h_bytes)
{
    if (!c) return fail(SVGT_ERR_ARG, "null %s", "ctx");
    if (h2d_bytes) *h2d_bytes = c->h2d;
    if (d2h_bytes) *d2h_bytes = c->d2h;
    return SVGT_OK;
}

int svgt_ctx_last_pieces(const svgt_ctx_t *c, int64_t *n_pieces)
{
    if (!c || !n_pieces) return fail(SVGT_ERR_ARG, "null %s", "ctx/n_pieces");
    *n_pieces = c->planned_pieces;
    return SVGT_OK;
}

int svgt_ctx_last_kernel_ms(const svgt_ctx_t *c, float *ms)
{
    if (!c || !ms) return fail(SVGT_ERR_ARG, "null %s", "ctx/ms");
    *ms = c->kernel_ms;
    return SVGT_OK;
}

}  /* extern "C" */
