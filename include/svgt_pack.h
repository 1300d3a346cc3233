/*
 * svgt_pack.h -- C ABI of the native evidence packer (libsvgt_pack.so), the host-side producer of the
 * rows libsvgt.so scores.
 *
 * It replaces, for a batch of breakpoints of one BAM, the reference's per-breakpoint read gathering
 *
 *     gather_reads(sample, chromA, posA, ciA, ..., z, max_reads)     svtyper/classic.py:54-100
 *     gather_reads / count + fetch per region                         svtyper/singlesample.py:139-205
 *     SamFragment.add_read (dedup, primary / split bookkeeping)       svtyper/parsers.py:748-768
 *     SplitRead.is_valid (split / soft-clip candidate QC)             svtyper/parsers.py:959-1058
 *
 * and this repo's evidence.BatchPacker (the 32-byte fragment and split rows of DESIGN.md 3).  BAM access
 * is its own BGZF + BAI reader (zlib inflate) with the pysam semantics the reference relies on
 * (SURVEY.md 8c): fetch(c, s, e) = file-order records with pos < e and end > s; count(..., 'all') skips
 * flags 0x4|0x100|0x200|0x400.
 *
 * Plain C: pointers and sizes only.  Every entry point returns 0 or a negative svgt_pack_err and never
 * throws.  CPU only (read gathering stays on the host by design: BASELINE.json north_star); the Python
 * gather path (svtyper_b200/gather.py + evidence.BatchPacker) is its parity checker.
 */
#ifndef SVGT_PACK_H
#define SVGT_PACK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVGT_PACK_ABI_VERSION 2

enum svgt_pack_err {
    SVGT_PACK_OK = 0,
    SVGT_PACK_ERR_ARG = -1,      /* null pointer / bad size / bad mode                                   */
    SVGT_PACK_ERR_IO = -2,       /* open / read / inflate failure, not a BAM / BAI                       */
    SVGT_PACK_ERR_RG = -3,       /* a usable read has no RG tag or an RG the table does not list
                                    (the reference raises KeyError: parsers.py:622, classic.py:85)       */
    SVGT_PACK_ERR_RECORD = -4    /* a mapped primary read without CIGAR, a malformed SA tag               */
};

#define SVGT_PACK_MODE_SSO 0      /* count both windows first, then fetch both (singlesample.py:168-205) */
#define SVGT_PACK_MODE_CLASSIC 1  /* fetch A then B; too many = a record index past max_reads (classic.py:79-91) */

typedef struct svgt_bam svgt_bam_t;

/* Fetch windows of one breakpoint: 0-based half-open, already clamped by the caller
 * (max(pos + ci0 - flank, 0) .. min(pos + ci1 + flank, contig length); the flank arithmetic is fp64 in
 * the reference, so it stays with the Python caller). */
typedef struct svgt_pack_site {
    int32_t tidA, begA, endA;
    int32_t tidB, begB, endB;
} svgt_pack_site_t;

/* Per-site result: rows appended for the site; skip != 0 when the site has too many reads. */
typedef struct svgt_pack_count {
    int32_t n_frag_rows, n_split_rows, skip, n_fragments;
} svgt_pack_count_t;

int svgt_pack_abi_version(void);
const char *svgt_pack_last_error(void);          /* thread-local, valid until the next call */

/* `bai_path` may be NULL: <bam>.bai, then <bam minus extension>.bai are tried. */
int svgt_bam_open(const char *bam_path, const char *bai_path, svgt_bam_t **out);
int svgt_bam_close(svgt_bam_t *bam);
int svgt_bam_n_references(const svgt_bam_t *bam);
const char *svgt_bam_reference_name(const svgt_bam_t *bam, int tid);
int64_t svgt_bam_reference_length(const svgt_bam_t *bam, int tid);

/* pysam-style count(contig, start, stop, read_callback='all' | 'nofilter'); negative = svgt_pack_err */
int64_t svgt_bam_count(svgt_bam_t *bam, int tid, int64_t beg, int64_t end, int filter_all);

/*
 * Gather and pack `n_sites` breakpoints.  Read groups: `rg_names[i]` belongs to library `rg_lib[i]`;
 * reads of libraries with lib_active[lib] == 0 are ignored.  `max_reads` < 0 means no limit.
 * Sites are independent, so blocks of 16 sites are spread over `n_threads` worker threads, each with
 * its own file handle and block cache (<= 0: one per hardware thread); the rows come back in site order.
 * The rows live in buffers owned by `bam` until the next svgt_pack_sites() / svgt_bam_close():
 * fetch them with svgt_pack_rows().  counts[n_sites] is caller-owned.
 */
int svgt_pack_sites(svgt_bam_t *bam, const svgt_pack_site_t *sites, int64_t n_sites,
                    const char *const *rg_names, const int32_t *rg_lib, int32_t n_rg,
                    const uint8_t *lib_active, int32_t n_lib, int32_t mode, int64_t max_reads,
                    int32_t n_threads, svgt_pack_count_t *counts);

/*
 * Library statistics inputs in ONE pass over the head of the BAM, replacing the
 * three per-library passes of Library.calc_read_length / calc_insert_hist / calc_lib_prevalence
 * (svtyper/parsers.py:501-576): per library the longest query length among its first
 * `read_length_reads` + 1 reads (reference: 10000), the template-length counts of its first `num_samp`
 * forward reads with a mapped reverse mate (tlen > 0, primary), and its share of the first
 * `prevalence_records` records (reference: 100000).  The median / MAD trimming and mean / sd stay in
 * Python (sample.py), on the table fetched with svgt_bam_scan_hist() in first-seen key order.
 */
typedef struct svgt_lib_scan {
    int64_t read_length, lib_records, records_seen, n_hist;
} svgt_lib_scan_t;

int svgt_bam_scan_libraries(svgt_bam_t *bam, const char *const *rg_names, const int32_t *rg_lib, int32_t n_rg,
                            int32_t n_lib, int64_t num_samp, int64_t read_length_reads, int64_t prevalence_records,
                            svgt_lib_scan_t *out);
int svgt_bam_scan_hist(const svgt_bam_t *bam, int32_t lib, const int32_t **keys, const int64_t **counts, int64_t *n);

/* Row buffers of the last svgt_pack_sites(): [n_frag][8] and [n_split][8] int32 words. */
int svgt_pack_rows(const svgt_bam_t *bam, const int32_t **frags, int64_t *n_frag, const int32_t **splits,
                   int64_t *n_split);

/*
 * Wide rows (svtyper_b200/evidence.py: what svgt_pack_sites() and the Python gather emit) -> the compact
 * 16-byte rows libsvgt.so's default kernel streams (svtyper_b200/compact.py is the specification; its numpy
 * converter is the parity checker).  Two calls, so the caller can allocate the destination (e.g. pinned host
 * memory) in between: svgt_compact_count() fills counts[n_sites][2] = compact fragment / split rows per site
 * and row_off[n_sites + 1] = their prefix sum; svgt_compact_fill() writes out_sites[n_sites][12] and
 * out_rows[row_off[n_sites]][4].  `min_aligned` is the -m the is_ref_seq bits of gapped reads are evaluated
 * with (reference parsers.py:801-816).  Sites are independent: blocks of them go to `n_threads` threads.
 */
int svgt_compact_count(const int32_t *sites, int64_t n_sites, const int32_t *frags, int64_t n_frag, const int32_t *splits,
                       int64_t n_split, int32_t min_aligned, int32_t n_threads, int64_t *row_off, int32_t *counts);
int svgt_compact_fill(const int32_t *sites, int64_t n_sites, const int32_t *frags, int64_t n_frag, const int32_t *splits,
                      int64_t n_split, int32_t min_aligned, int32_t n_threads, const int64_t *row_off, const int32_t *counts,
                      int32_t *out_sites, int32_t *out_rows);

/*
 * FORMAT text of `n` scored 80-byte rows (svgt_out_row_t), the vectorised twin of Genotype.set_format +
 * get_gt_string (reference parsers.py:375-398) for the values bayesian_genotype() produces
 * (singlesample.py:406-473): `order[n_fields]` lists the fields in header order (0 GT 1 GQ 2 SQ 3 GL 4 DP 5 RO
 * 6 AO 7 QR 8 QA 9 RS 10 AS 11 ASC 12 RP 13 AP 14 AB); style[i] = 0 every field of a scored row (GT / GQ / SQ
 * read "./." / "." / "." when the row is not called), 1 the blank row (blank_genotype_result()), 2 "./." and "."
 * for every other field.  Floats print as %0.2f, GL as %.0f, AB as %.2g of QA / (QR + QA).  Row i's text goes to
 * out + i * stride (no terminator), its length to lengths[i].  svgt_format_quals() does %0.2f of QUAL values.
 */
int svgt_format_calls(const void *rows, int64_t n, const int32_t *order, int32_t n_fields, const uint8_t *style,
                      int32_t n_threads, char *out, int64_t stride, int32_t *lengths);
int svgt_format_quals(const double *qual, int64_t n, char *out, int64_t stride, int32_t *lengths);

/* GT / GQ / SQ of `n` scored rows recomputed in place from their GL values with the host libm, the way CPython
 * evaluates singlesample.py:447-471 (the device's pow / log are not correctly rounded; GL and every count are
 * bit-exact already). */
int svgt_host_sq(void *rows, int64_t n, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* SVGT_PACK_H */
