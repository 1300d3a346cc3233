/*
 * svgt_kernels.cuh -- launch parameters shared by the kernels and the C ABI.
 *
 * Internal header (the public boundary is include/svgt.h).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/svgt.h"

#define SVGT_THREADS 128            /* 4 warps per CTA, one site per thread        */
#define SVGT_WARPS (SVGT_THREADS / 32)
#define SVGT_SMEM_LIBS 64           /* library rows kept in shared memory          */
#define SVGT_SMEM_HIST_WORDS 6144   /* 24 KB of insert-size histogram in smem      */
#define SVGT_STAGE_ROWS 4           /* rows per lane per staged step (bulk path)   */

/* variants of the WIDE-row compatibility kernel (svgt_set_variant / SVGT_VARIANT env): how fragment rows reach a lane */
enum {
    SVGT_VAR_DIRECT = 0,   /* per-lane LDG.128 x2 with register prefetch            */
    SVGT_VAR_BULK = 1,     /* per-lane cp.async.bulk (TMA 1-D) ring in shared memory */
    SVGT_VAR_COUNT = 2
};

struct SvgtParams {
    const int4 *sites;  long long n_sites;
    const int4 *frags;  long long n_frag;
    const int4 *splits; long long n_split;
    const int *order;
    const double *lib_f64;
    const int4 *lib_i32;
    int n_lib;
    const unsigned *hist; long long n_hist;
    const double *pm;
    const double *logt; long long n_log;
    const double *consts;
    int min_aligned, split_slop, assoc_mode;
    double split_weight, disc_weight;
    svgt_out_row_t *out;
    int *status;          /* [0] first error, [1] work cursor, [2] error count, [3] finished-CTA count */
    int n_tiles;
    int hist_in_smem;
};

/* compact-schema launch (svgt_compact.cu): `base` carries the tables, sizes, `out` (local rows: the sums are
 * parked there between the two launches) and `status`; its sites / frags / splits are unused */
struct SvgtCompactParams {
    SvgtParams base;
    const int4 *sites;              /* [n_sites][3]  48-byte site rows          */
    const int4 *rows; long long n_rows;   /* [n_rows]  16-byte evidence rows    */
    svgt_out_row_t *out_final;      /* where the call kernel writes the final rows (NULL: base.out) */
    int *done_flag; int done_value; /* optional: set to done_value (system scope) once every final row is written */
    unsigned hist_max;              /* largest histogram count, 0 = unknown     */
    int ramp;
    /* piece plan (svgt_segplan_t, include/svgt.h): sites too long for one warp are scored in pieces of whole 32-row
     * chunks whose per-row addends go to `scratch`; svgt_replay_pieces_kernel then sums each such site in row order */
    const int *entries; long long n_entries;    /* launch list: site index, or ~k for piece k; NULL: no plan */
    const int4 *pieces; long long n_pieces;     /* site, first row within the site's part, rows | split << 31, scratch chunk */
    const int4 *heavy;  long long n_heavy;      /* site, first scratch chunk, fragment chunks, split chunks */
    double *scratch;                            /* [scratch_chunks][3][32] parked addends */
    int *scratch_lead;                          /* [scratch_chunks] lead rows of each chunk */
    long long scratch_chunks;
};

/* Returns a cudaError_t as int.  `grid` <= 0 lets the launcher size a persistent grid. */
int svgt_launch_score(const SvgtParams &p, int variant, cudaStream_t stream);
size_t svgt_score_smem_bytes(const SvgtParams &p, int variant);
int svgt_launch_compact(const SvgtCompactParams &cp, int unit_mode, cudaStream_t stream);
