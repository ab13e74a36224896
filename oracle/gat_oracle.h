/*
 * gat_oracle.h -- CPU restatement of the GAT simulation hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the *checker* for gat_b200's CUDA path, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  Every function cites the
 * reference (AndreasHeger/gat 1.3.6, paths relative to /root/reference) it restates.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against
 * (a) known answers from the reference's own tests (test/test_SegmentList.py, test/test_gat.py) and
 * (b) golden vectors produced by the compiled reference itself (tests/golden/make_golden.py, run in
 *     the build container against oracle/_ref); the sampler is pinned bit-for-bit by driving it with
 *     numpy's legacy global RNG through the randint callback, exactly as the reference does.
 *
 * Domain: Position = uint32, PositionDifference = int32 (gat/SegmentList.pxd:31-38); like the
 * reference, arithmetic is only meaningful for coordinates < 2^31.
 */
#ifndef GAT_ORACLE_H
#define GAT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t start; uint32_t end; } go_seg;

/* --- utils/gat_utils.c ------------------------------------------------------------------------*/
long go_searchsorted_u32(const uint32_t *base, size_t n, uint32_t target);          /* :8-32 + cmpPosition */
long go_searchsorted_seg(const go_seg *base, size_t n, go_seg target);              /* :8-32 + cmpSegments */
long go_searchargsorted_f64(const double *base, const int *sorted, size_t n, double target); /* :37-61 + cmpDouble */

/* --- gat/SegmentList.pyx ----------------------------------------------------------------------*/
void     go_sort(go_seg *s, size_t n);                                              /* :478-486 */
size_t   go_normalize(go_seg *s, size_t n);                                         /* :697-754 */
size_t   go_merge(go_seg *s, size_t n, int32_t distance);                           /* :756-816 */
size_t   go_filter(const go_seg *self, size_t n, const go_seg *other, size_t m, go_seg *out);      /* :1401-1467 */
size_t   go_intersect(const go_seg *self, size_t n, const go_seg *other, size_t m, go_seg *out);   /* :1469-1549; out cap n+m */
uint32_t go_sum(const go_seg *s, size_t n);                                         /* :1607-1616 */
uint32_t go_overlap_with_segments(const go_seg *self, size_t n, const go_seg *other, size_t m);    /* :1026-1076 */
uint32_t go_intersection_with_segments(const go_seg *self, size_t n, const go_seg *other, size_t m,
                                       int midpoint);                                /* :1078-1146 */
int      go_get_insertion_point(const go_seg *s, size_t n, go_seg other);            /* :853-887 */
int      go_trim_ends(go_seg *s, size_t n, uint32_t pos, uint32_t size, int forward);/* :545-597 */
/* histogram of ceil(len/bucket) (:1148-1184); returns the bucket size used, 0 on "segment too large" */
uint32_t go_length_distribution(const go_seg *s, size_t n, uint32_t bucket_size, uint32_t nbuckets,
                                int64_t *histogram);

/* --- random numbers -----------------------------------------------------------------------------
 * The sampler draws through a callback so that ONE restatement can be driven by
 *  - numpy.random.randint of the process-global legacy RNG (pins it against the compiled reference), or
 *  - the counter-based Philox stream of the CUDA kernel (pins the kernel against it).
 * randint(ctx, slot, lo, hi) returns an integer uniform on [lo, hi) (numpy.random.randint semantics);
 * `slot` names the draw inside the current loop turn; next_turn(ctx) is called at the top of every
 * turn of the placement loop (gat/Engine.pyx:572).  A sequential generator ignores both. */
enum { GO_SLOT_LEN = 0, GO_SLOT_JITTER = 1, GO_SLOT_WS_R = 2, GO_SLOT_WS_P = 3,
       GO_SLOT_TRIM_R = 4, GO_SLOT_TRIM_P = 5, GO_SLOT_TRIM_DIR = 6,
       /* SamplerShift: one Philox block per segment (block 0: words 0,1 = position, words 2,3 = direction) */
       GO_SLOT_SHIFT_POS = 0, GO_SLOT_SHIFT_DIR = 2 };
typedef int64_t (*go_randint_fn)(void *ctx, int slot, int64_t lo, int64_t hi);
typedef void    (*go_turn_fn)(void *ctx);

/* Philox4x32-10 stream shared bit-for-bit with gat_b200/csrc (see DESIGN.md "RNG contract"). */
typedef struct {
    uint64_t seed; uint32_t track; uint32_t unit; uint32_t sample; uint32_t turn;
} go_philox_ctx;
void    go_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void    go_philox_begin(go_philox_ctx *c, uint64_t seed, uint32_t track, uint32_t unit, uint32_t sample);
int64_t go_philox_randint(void *ctx, int slot, int64_t lo, int64_t hi);
void    go_philox_next_turn(void *ctx);

/* --- gat/Engine.pyx samplers --------------------------------------------------------------------*/
/* SegmentListSampler.sample (:279-348) on list ws; returns 0 ok */
int go_segmentlist_sample(const go_seg *ws, size_t m, const uint32_t *cdf, uint32_t total,
                          uint32_t sample_length, go_randint_fn rnd, void *ctx, int slot_r, int slot_p,
                          uint32_t *start, uint32_t *end, int32_t *overlap);

typedef struct {
    int32_t  ltotal;            /* bases to reproduce                               */
    int32_t  true_remaining;    /* at exit                                          */
    int32_t  nunsuccessful;     /* non-improving checkpoints (max 20)               */
    uint32_t nturns;            /* loop turns executed                              */
    uint32_t nplaced;           /* placements appended                              */
    uint32_t ncheckpoints;
    uint32_t ntrims;
    uint32_t bucket_size;       /* bucket size actually used                        */
} go_sample_info;

/* SamplerAnnotator.sample (:515-646).  out must hold `cap` segments; returns the number of segments
 * written, or -1 capacity exceeded, -2 segment too large for nbuckets*bucket_size (ValueError in
 * the reference), -3 out of memory. */
long go_sampler_annotator(const go_seg *segments, size_t n, const go_seg *workspace, size_t m,
                          uint32_t bucket_size, uint32_t nbuckets,
                          go_randint_fn rnd, go_turn_fn next_turn, void *ctx,
                          go_seg *out, size_t cap, go_sample_info *info);

/* SamplerSegments.sample (gat/Engine.pyx:653-737): len(segments) placements in draw order (unsorted,
 * unmerged).  Returns the count, or -1/-2/-3 as above. */
long go_sampler_segments(const go_seg *segments, size_t n, const go_seg *workspace, size_t m,
                         uint32_t bucket_size, uint32_t nbuckets,
                         go_randint_fn rnd, go_turn_fn next_turn, void *ctx, go_seg *out, size_t cap);
/* sampler used by go_compute_sample_philox: 0 = SamplerAnnotator (default), 1 = SamplerSegments */
void go_set_sampler_kind(int kind);
/* SamplerShift.sample (gat/Engine.pyx:998-1111); -5 where the reference raises ValueError (empty local workspace) */
long go_sampler_shift(const go_seg *segments, size_t n, const go_seg *workspace, size_t m,
                      double radius, int32_t extension,
                      go_randint_fn rnd, go_turn_fn next_turn, void *ctx, go_seg *out, size_t cap);
void go_set_shift_params(double radius, int32_t extension);

/* --- counters (gat/Engine.pyx:1412-1472) ----------------------------------------------------------
 * counter ids are shared with include/gat_b200.h */
enum { GO_NUCLEOTIDE_OVERLAP = 0, GO_NUCLEOTIDE_DENSITY = 1, GO_SEGMENT_OVERLAP = 2,
       GO_SEGMENT_MIDOVERLAP = 3, GO_ANNOTATION_OVERLAP = 4, GO_ANNOTATION_MIDOVERLAP = 5,
       GO_NCOUNTERS = 6 };
double go_counter(int counter, const go_seg *segments, size_t n, const go_seg *annotations, size_t m,
                  size_t workspace_nsegments);

/* one whole sample (gat/__init__.py:494-591): U placement units -> per-contig merge(0)
 * (Engine.pyx:2857-2876) -> counts[counter][annotation] summed over the sample's contigs in
 * first-appearance order.  All lists are CSR (offsets + go_seg arrays).
 *   unit_contig[u]  contig index of unit u;  has_isochores: apply merge(0) per contig
 *   anno: A*C lists, annotation-major;  cws_nseg[c]: number of contig-workspace segments
 *   counts: ncounters*A doubles, counter-major.  Philox keyed by (seed, track, unit, sample).
 * Optionally returns the contig-level sample (placed_off C+1, placed cap).  Returns 0 or <0. */
void go_set_unit_cap(size_t cap);   /* room per unit for a sampler's result (0 = default) */
int go_compute_sample_philox(int U, int C, int A, const int32_t *unit_contig, int has_isochores,
                             const uint64_t *seg_off, const go_seg *seg,
                             const uint64_t *ws_off, const go_seg *ws,
                             const uint64_t *anno_off, const go_seg *anno,
                             const uint32_t *cws_nseg,
                             uint32_t bucket_size, uint32_t nbuckets,
                             uint64_t seed, uint32_t track, uint32_t sample,
                             int ncounters, const int32_t *counters, double *counts,
                             uint64_t *placed_off, go_seg *placed, size_t placed_cap);

/* counts for an already placed, contig-level sample (same-placement parity) */
void go_count_placed(int C, int A, const uint64_t *placed_off, const go_seg *placed,
                     const uint64_t *anno_off, const go_seg *anno, const uint32_t *cws_nseg,
                     int ncounters, const int32_t *counters, double *counts);

/* --- statistics (gat/Engine.pyx:1543-1576, 1635-1718) -------------------------------------------*/
typedef struct {
    double observed, expected, stddev, lower95, upper95, fold, pvalue, qvalue;
    uint32_t nsamples;
} go_stats;
/* has_reference: expected *= ref_fold, p-value of observed/ref_fold, CI *= ref_fold (:1673-1714) */
int    go_enrichment_statistics(double observed, const double *samples, size_t l, int has_reference,
                                double ref_fold, double pseudo_count, go_stats *out);
double go_two_sided_pvalue(const double *sorted_samples, size_t l, double expected, double val);

#ifdef __cplusplus
}
#endif
#endif
