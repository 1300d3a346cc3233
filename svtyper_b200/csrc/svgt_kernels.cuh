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

/* kernel variants (svgt_set_variant / SVGT_VARIANT env): how fragment rows reach a lane */
enum {
    SVGT_VAR_DIRECT = 0,   /* per-lane LDG.128 x2 with register prefetch            */
    SVGT_VAR_BULK = 1,     /* per-lane cp.async.bulk (TMA 1-D) ring in shared memory */
    SVGT_VAR_COOP = 2,     /* warp-cooperative, 8 sites interleaved per warp (svgt_coop.cu) */
    SVGT_VAR_COOP4 = 3,    /* warp-cooperative, 4 sites interleaved per warp                 */
    SVGT_VAR_RING = 4,     /* warp-cooperative, rows through a cp.async.bulk smem ring (svgt_ring.cu) */
    SVGT_VAR_LEAN = 5,     /* warp-cooperative with the lean row scorer (svgt_lean.cu), the default:
                              8 sites per work unit after a 1-2-4 ramp over the heaviest sites            */
    SVGT_VAR_LEAN8 = 6,    /* the same, always 8 sites per unit (tests)                                     */
    SVGT_VAR_LEAN2 = 7,    /* the same, always 2 sites per unit (tests)                                     */
    SVGT_VAR_COUNT = 8
};
#define SVGT_COOP_THREADS 256       /* 8 warps per CTA in the cooperative kernel     */

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
    int *status;          /* [0] first error, [1] tile cursor, [2] error count, [3] spare */
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
};

/* Returns a cudaError_t as int.  `grid` <= 0 lets the launcher size a persistent grid. */
int svgt_launch_score(const SvgtParams &p, int variant, cudaStream_t stream);
size_t svgt_score_smem_bytes(const SvgtParams &p, int variant);
int svgt_launch_coop(const SvgtParams &p, int variant, cudaStream_t stream);
int svgt_launch_call(const SvgtParams &p, cudaStream_t stream);
int svgt_launch_ring(const SvgtParams &p, cudaStream_t stream);
int svgt_launch_lean(const SvgtParams &p, int variant, cudaStream_t stream);
int svgt_lean_launches(void);
int svgt_launch_compact(const SvgtCompactParams &cp, int unit_mode, cudaStream_t stream);
