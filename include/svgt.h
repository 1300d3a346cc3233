/*
 * svgt.h -- C ABI of the B200-native SV genotype-likelihood engine (libsvgt.so).
 *
 * This is the drop-in boundary for ONE path of hall-lab/svtyper v0.7.1: the scoring
 * segment of a breakpoint batch,
 *
 *     counts = tally_variant_read_fragments(split_slop, min_aligned, breakpoint,
 *                                           sam_fragments, debug)   singlesample.py:355
 *     result = bayesian_genotype(breakpoint, counts, split_weight,
 *                                disc_weight, debug)                singlesample.py:406
 *
 * as called per breakpoint inside parallel_calculate_genotype (singlesample.py:523-536),
 * serial_calculate_genotype (:486-498) and, inlined, classic.sv_genotype
 * (classic.py:286-495).  One call scores a whole batch of breakpoints.  Read gathering
 * (pysam) and VCF I/O stay on the host side of this boundary.
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  All entry points return 0 on
 * success or a negative svgt_err code, never throw, and are re-entrant per stream.
 * Row layouts are documented in svtyper_b200/compact.py (the default, compact schema), svtyper_b200/evidence.py
 * (the wide interchange schema) and DESIGN.md.
 */
#ifndef SVGT_H
#define SVGT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVGT_ABI_VERSION 3

#define SVGT_SITE_WORDS 16  /* int32 words per site row      (64 B) */
#define SVGT_FRAG_WORDS 8   /* int32 words per fragment row  (32 B) */
#define SVGT_SPLIT_WORDS 8  /* int32 words per split row     (32 B) */
#define SVGT_CSITE_WORDS 12 /* compact schema: int32 words per site row (48 B)  */
#define SVGT_CROW_WORDS 4   /* compact schema: int32 words per evidence row (16 B) */
#define SVGT_OUT_BYTES 80   /* f64 GL[3], f64 SQ, int32 GT,GQ,DP,RO,AO,QR,QA,RS,AS,ASC,RP,AP */

enum svgt_err {
    SVGT_OK = 0,
    SVGT_ERR_ARG = -1,          /* null pointer / negative size / bad mode            */
    SVGT_ERR_CUDA = -2,         /* CUDA runtime error, see svgt_last_error()          */
    SVGT_ERR_LOG_TABLE = -3,    /* a site's QR+QA exceeded the log10 table (n_log)    */
    SVGT_ERR_LIB_INDEX = -4,    /* a fragment row names a library >= n_lib            */
    SVGT_ERR_NO_DEVICE = -5,    /* no CUDA device (there is no CPU fallback)          */
    SVGT_ERR_RANGE = -6         /* a site coordinate / window is outside +-2^30       */
};

/* assoc_mode: which reference entry point's floating-point association order to follow */
#define SVGT_ASSOC_SSO 0      /* singlesample.py:367-378: per-fragment sub-totals      */
#define SVGT_ASSOC_CLASSIC 1  /* classic.py:311,326-328: every read added straight in  */

/* Output GT codes (reference singlesample.py:462-471, :207-243) */
#define SVGT_GT_UNDERFLOW (-1) /* "./." with GL/counts kept: all 10**GL underflowed   */
#define SVGT_GT_BLANK (-2)     /* no evidence: the blank_genotype_result() row        */
#define SVGT_GT_SKIPPED (-3)   /* site row had the SKIP bit (too many reads)          */

typedef struct svgt_out_row {
    double gl[3];
    double sq;
    int32_t gt, gq, dp, ro, ao, qr, qa, rs, as_, asc, rp, ap;
} svgt_out_row_t;

/*
 * One batch.  In svgt_score_batch() every pointer is a DEVICE pointer; in
 * svgt_ctx_score_host() every pointer is a HOST pointer.
 */
typedef struct svgt_batch {
    const int32_t *sites;   int64_t n_sites;  /* [n_sites][16]                           */
    const int32_t *frags;   int64_t n_frag;   /* [n_frag][8]  sorted(query_name) per site */
    const int32_t *splits;  int64_t n_split;  /* [n_split][8]                            */
    const int32_t *order;                     /* optional site permutation (work-bucketed
                                                 launch order), NULL = identity          */
    const double *lib_f64;                    /* [n_lib][4] flank, 2*sd, N, mean         */
    const int32_t *lib_i32;                   /* [n_lib][4] hist_off, hist_len, nondel_L */
    int32_t n_lib;
    const uint32_t *hist;   int64_t n_hist;   /* insert-size histogram counts            */
    const double *pm;                         /* [256]   prob_mapq LUT (utils.py:74)     */
    const double *logt;     int64_t n_log;    /* [n_log] math.log(n, 10) LUT             */
    const double *consts;                     /* [32]    priors / log10 prior constants  */
    int32_t min_aligned;                      /* reference -m, default 20                */
    int32_t split_slop;                       /* reference constant 3                    */
    int32_t assoc_mode;                       /* SVGT_ASSOC_*                            */
    int32_t reserved;
    double split_weight, disc_weight;         /* reference --split_weight/--disc_weight  */
} svgt_batch_t;

/*
 * Piece plan of a compact batch (optional; svgt_plan_count / svgt_plan_fill build it from the HOST site rows).
 *
 * The fp64 sums of a breakpoint run over its fragments in sorted(query_name) order (singlesample.py:364-378), so one
 * warp walks a site's rows from first to last.  In a small or heavy-tailed batch (config `stress1m`: up to 10,000
 * reads per site) the longest site then IS the run time.  With a plan, a site of more than `max_chunks` 32-row
 * chunks is scored in pieces -- runs of whole chunks of its fragment rows or of its split rows, each handed to
 * whichever warp is free -- whose per-row addends are parked in `scratch`; a second, small kernel then adds them
 * up per site in row order.  Results are bit-identical with and without a plan.
 *   entries  the launch list, heaviest first: a site index (>= 0; sites without rows are left out), or ~k for piece k
 *   pieces   [n_pieces][4]  site, first row within the site's fragment (split) rows (a multiple of 32),
 *            rows | (split rows ? 1 << 31 : 0), first scratch chunk
 *   heavy    [n_heavy][4]   site, first scratch chunk, fragment chunks, split chunks -- one per site scored in pieces;
 *            the site's chunks occupy scratch chunks [first, first + fragment chunks + split chunks) in row order
 *   scratch  scratch_chunks * SVGT_PLAN_CHUNK_BYTES bytes of device memory, 16-byte aligned, contents don't matter
 * Every pointer is a DEVICE pointer in svgt_score_compact(); svgt_ctx_score_host_compact() ignores `plan` and
 * plans by itself.
 */
#define SVGT_PLAN_CHUNK_BYTES 784  /* 3 x 32 parked doubles + the chunk's lead-row count, padded to 16 bytes */
typedef struct svgt_segplan {
    const int32_t *entries; int64_t n_entries;
    const int32_t *pieces;  int64_t n_pieces;
    const int32_t *heavy;   int64_t n_heavy;
    void *scratch;          int64_t scratch_chunks;
} svgt_segplan_t;

/*
 * The same batch in the COMPACT schema (svtyper_b200/compact.py; the default product path): 48-byte site
 * rows, and one array of 16-byte rows holding each site's fragment rows followed by its split rows.
 * In svgt_score_compact() every pointer is a DEVICE pointer; in svgt_ctx_score_host_compact() every
 * pointer is a HOST pointer (out_final / done_flag must be NULL there).
 */
#define SVGT_LAYOUT_SITE_ORDER 1   /* flags: rows are laid out in site order (row_off non-decreasing, each
                                      site's rows directly behind the previous site's) -- what every packer
                                      of this repo emits; lets the host path pipeline site slices         */
typedef struct svgt_cbatch {
    const int32_t *sites;   int64_t n_sites;  /* [n_sites][12]                           */
    const int32_t *rows;    int64_t n_rows;   /* [n_rows][4]                             */
    const int32_t *order;                     /* optional launch permutation, NULL = identity */
    const double *lib_f64;                    /* [n_lib][4] flank, 2*sd, N, mean         */
    const int32_t *lib_i32;                   /* [n_lib][4] hist_off, hist_len, nondel_L */
    int32_t n_lib;
    uint32_t hist_max;                        /* largest histogram count; 0 = unknown    */
    const uint32_t *hist;   int64_t n_hist;
    const double *pm;                         /* [256]                                   */
    const double *logt;     int64_t n_log;
    const double *consts;                     /* [32]                                    */
    int32_t min_aligned, split_slop, assoc_mode;
    int32_t unit_mode;                        /* work units of the tally kernel: 0 ramped 1-2-4-6 sites when the batch is
                                                 small (decided by site count); 1 always 6 sites; 3 always ramped --
                                                 what a caller who has the sites' row counts should pick (ramp iff the
                                                 six heaviest sites outweigh a warp's fair share); 2 two sites (tests) */
    double split_weight, disc_weight;
    void *out_final;                          /* optional: the final 80-byte rows go HERE instead of out_rows
                                                 (e.g. a peer-mapped buffer of the gathering GPU); out_rows
                                                 is then scratch                                          */
    int32_t *done_flag;                       /* optional: set to done_value with system scope after the
                                                 last final row is written (may be peer-mapped)           */
    int32_t done_value;
    int32_t flags;                            /* SVGT_LAYOUT_*                           */
    int32_t rows_min_aligned;                 /* the -m the packer evaluated the MULTI rows' is_ref_seq bits
                                                 with; must equal min_aligned (SVGT_ERR_ARG otherwise)     */
    int32_t reserved;
    const svgt_segplan_t *plan;               /* optional piece plan (host struct, device pointers inside); with a
                                                 plan `order` is not used: plan->entries is the launch list   */
} svgt_cbatch_t;

int svgt_abi_version(void);
const char *svgt_last_error(void);            /* thread-local, valid until the next call */
int svgt_device_count(void);

/*
 * Score a device-resident batch.  Asynchronous on `stream` (a cudaStream_t passed as
 * void*).  `out_rows` = n_sites * 80 device bytes in ORIGINAL site order.  `status` = 4
 * device int32 words, zeroed by the call; after the stream is synchronised status[0] is
 * 0 or the first svgt_err a site raised.  The callee allocates nothing.
 */
int svgt_score_batch(const svgt_batch_t *batch, void *out_rows, int32_t *status, void *stream);

/* The same for a compact batch: the default path (svgt_compact_kernel + svgt_call_compact_kernel). */
int svgt_score_compact(const svgt_cbatch_t *batch, void *out_rows, int32_t *status, void *stream);

/*
 * Build a piece plan from HOST site rows ([n_sites][12]).  svgt_plan_count() decides the piece length
 * (*max_chunks; `force_chunks` > 0 dictates it, otherwise it grows with the batch so that no piece outlasts a small
 * share of the batch's run time on `resident_warps` warps -- <= 0: those of the current device -- and is never
 * below 4 chunks) and returns the array sizes; all four are 0 when no site is longer than a piece (score without a
 * plan then).  svgt_plan_fill() writes the arrays (host memory of at least those sizes) for the same arguments.
 */
int svgt_plan_count(const int32_t *sites_host, int64_t n_sites, int32_t min_aligned, int32_t split_slop,
                    int32_t resident_warps, int32_t force_chunks, int32_t *max_chunks, int64_t *n_entries,
                    int64_t *n_pieces, int64_t *n_heavy, int64_t *scratch_chunks);
int svgt_plan_fill(const int32_t *sites_host, int64_t n_sites, int32_t min_aligned, int32_t split_slop,
                   int32_t max_chunks, int32_t *entries, int32_t *pieces, int32_t *heavy);

/* Number of kernel launches svgt_score_batch issues for this batch (bench bookkeeping). */
int svgt_launches_per_batch(const svgt_batch_t *batch);

/*
 * Variant of the WIDE-row kernel behind svgt_score_batch() (kept as an independent cross-check of the default
 * compact path; same results): 0 = thread-per-site, per-lane 128-bit global loads with register prefetch;
 * 1 = thread-per-site, per-lane cp.async.bulk (TMA 1-D) ring in shared memory.  -1 restores the default (0, or
 * the SVGT_VARIANT environment variable).  Returns the variant now in force.
 */
int svgt_set_variant(int variant);

/*
 * Host-buffer convenience used by the reference-facing plug-in: owns device staging
 * buffers and a stream on `device`, copies the batch host->device, scores it, copies
 * the 80-byte rows back into `out_rows_host`, and synchronises.  Returns 0 or svgt_err.
 */
typedef struct svgt_ctx svgt_ctx_t;
int svgt_ctx_create(int device, svgt_ctx_t **ctx);
int svgt_ctx_destroy(svgt_ctx_t *ctx);
int svgt_ctx_score_host(svgt_ctx_t *ctx, const svgt_batch_t *host_batch, void *out_rows_host);
int svgt_ctx_score_host_compact(svgt_ctx_t *ctx, const svgt_cbatch_t *host_batch, void *out_rows_host);
/* bytes moved by the last svgt_ctx_score_host call */
int svgt_ctx_last_traffic(const svgt_ctx_t *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes);
/* pieces the last svgt_ctx_score_host_compact call cut its long sites into (0: it scored without a piece plan) */
int svgt_ctx_last_pieces(const svgt_ctx_t *ctx, int64_t *n_pieces);
/* device time (ms, CUDA events) of the kernel(s) in the last svgt_ctx_score_host call */
int svgt_ctx_last_kernel_ms(const svgt_ctx_t *ctx, float *ms);

/*
 * Multi-GPU output without a collective: a buffer of the gathering rank that the other ranks' call
 * kernels write straight into over NVLink.  svgt_shared_alloc() cudaMalloc's `bytes` on the current
 * device and exports a 64-byte CUDA IPC handle; svgt_shared_open() maps such a handle in another
 * process (peer access is enabled by the mapping); svgt_wait_flags() enqueues, on `stream`, a wait until
 * flags[i] >= value for all i < n (flags: device memory of the current device).  Two ways to fill such a buffer:
 * the call kernel stores its rows there itself (svgt_cbatch_t::out_final / done_flag: the transfer sits at the end
 * of the step), or svgt_peer_copy() + svgt_set_flag() forward a finished step's rows on a side stream -- the copy
 * engines move them under the NEXT step's kernels (cudaMemcpyAsync over NVLink; the flag is set, with system
 * scope, in stream order behind the copy).
 */
int svgt_shared_alloc(int64_t bytes, void **dev_ptr, unsigned char handle[64]);
int svgt_shared_open(const unsigned char handle[64], void **dev_ptr);
int svgt_shared_close(void *dev_ptr);
int svgt_shared_free(void *dev_ptr);
int svgt_wait_flags(const int32_t *flags, int32_t n, int32_t value, void *stream);
int svgt_peer_copy(void *dst, const void *src, int64_t bytes, void *stream);
int svgt_set_flag(int32_t *flag, int32_t value, void *stream);
int svgt_memcpy_d2h(void *dst_host, const void *src_dev, int64_t bytes);   /* synchronous read-back of such a buffer */

#ifdef __cplusplus
}
#endif
#endif /* SVGT_H */
