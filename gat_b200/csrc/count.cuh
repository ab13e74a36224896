// count.cuh -- counting kernel (K3/K4), the annotation tile format and column statistics (K5).
//
// Replaces overlapWithSegments / intersectionWithSegments (gat/SegmentList.pyx:1026-1146) as used by
// the Counter* classes (gat/Engine.pyx:1412-1472) inside computeSample (gat/__init__.py:580-587) and
// computeCounts (gat/Engine.pyx:2189-2202).
#pragma once

#include "common.cuh"

namespace gatb {

constexpr int KMAX = 8;                 // annotation tracks per tile (register accumulators)

// Tile = up to KMAX annotation tracks on ONE key (contig), contiguous in global memory:
//
//   TileHeader
//   uint32 bm[bm_words]      occupancy bitmap of the UNION of the tracks' intervals      } "filter":
//   uint16 idx[nbins+1]      bin index: where in civ[] a query starting in the bin begins } staged in
//   uint8  cslot[n_cons+2]   track slot of civ[c]                                        } shared
//   uint2  civ[n_cons+2]     every interval of every track, sorted by start, + 2 sentinels } memory
//   uint2  uiv[n_union+2]    union intervals (sorted, disjoint) + 2 sentinels   } global memory only: build
//   uint32 uoff[n_union+1]   union interval u = civ[uoff[u] .. uoff[u+1])       } time, and the nbins == 0 path
//
// About 93 % of the simulated segments overlap no interval of ANY of the 8 tracks.  The count kernel
// therefore tests a segment first against the bitmap -- bit b is set when a union interval touches
// positions [b << bm_shift, (b+2) << bm_shift); one 4-byte shared-memory load answers "can [s,e) overlap
// anything?" for every segment no longer than 1 << bm_shift, longer ones are candidates outright -- and
// only the candidates go on, through a per-warp queue, to the exact pass: idx[bin(s)] = uoff[first union
// interval whose end is > the lowest position of the bin], bin(x) = umulhi(x, inv) -- a one-probe
// replacement for the binary search (utils/gat_utils.c:8-32) over sorted interval ends; every interval
// overlapping [s,e) lies at or after that index, so a forward walk over civ[] until start >= e, skipping
// ends <= s, visits them all, entirely in shared memory.  nbins == 0 (more than 65534 intervals) falls
// back to the binary search over uiv[] in global memory.
struct TileHeader {
    uint32_t n_union;       // written by the build kernel
    uint32_t n_cons;
    uint32_t nbins;
    uint32_t inv;           // bin(x) = min(umulhi(x, inv), nbins)
    uint32_t bm_off;        // byte offsets from the tile start; bitmap: one zero word of padding at the end
    uint32_t bm_shift;      // log2 of the positions per bit
    uint32_t bm_bits;       // bits that can be set: positions >= bm_bits << bm_shift hold no interval
    uint32_t idx_off;       // [0, idx_off) = header + bitmap: staged for every tile
    uint32_t cslot_off;
    uint32_t civ_off;
    uint32_t stage_bytes;   // [0, stage_bytes) = header .. civ: staged when it fits
    uint32_t uiv_off;
    uint32_t uoff_off;
    uint32_t pad[3];
};

struct CountParams {
    // annotations
    const uint8_t *tiles;           // tile blob
    const uint64_t *tile_off;       // [n_groups][n_keys] byte offset
    const uint32_t *tile_stage;     // [n_groups][n_keys] bytes of the filter part (stage_bytes)
    const uint32_t *key_ws_nseg;    // [n_keys] or NULL
    uint32_t n_annot, n_keys, n_groups, ka;   // ka = tracks per group
    uint32_t smem_tile_budget;      // filters up to this many bytes are staged in shared memory
    // segment sets
    const uint64_t *placed;         // [n_samples][sample_stride] packed
    uint64_t sample_stride;
    const uint64_t *key_base;       // [n_keys]
    const uint32_t *placed_n;       // [n_samples][n_keys]
    const uint8_t *key_present;     // [n_samples][n_keys] or NULL
    uint32_t n_samples;
    uint32_t schunk;                // samples per CTA
    // output
    uint32_t *out_u32;              // [n_samples][n_annot] (integer counters)
    double *out_f64;                // [n_samples][n_annot] (nucleotide-density)
};

// shared memory a count launch needs besides the staged filter (accumulators + per-warp queues)
size_t count_smem_overhead(int threads, uint32_t schunk, bool density);

// Builds every tile from the raw annotation CSR arrays on the device (one CTA per tile): merges the
// tracks' lists by start, derives the union and its bin index, and validates the lists (error bit 0:
// coordinate >= 2^31, bit 1: empty / unsorted / overlapping = not normalized).
struct BuildTilesParams {
    uint8_t *tiles;
    const uint64_t *tile_off;       // [n_groups][n_keys]
    const TileHeader *headers;      // [n_groups][n_keys], geometry computed on the host
    const uint64_t *offs;           // [n_annot*n_keys+1]
    const uint32_t *start, *end;
    uint32_t n_annot, n_keys, n_groups, ka;
    uint32_t *error;
};
void launch_build_tiles(cudaStream_t st, const BuildTilesParams &p);

// counter: GATB_* id.  Returns cudaError from the launch configuration.
cudaError_t launch_count(cudaStream_t st, int counter, const CountParams &p, int threads);

// column statistics (K5)
struct StatsParams {
    const void *counts;             // [n_samples][n_cols] uint32 or float64
    int is_float;
    uint64_t n_samples;
    uint32_t n_cols;
    const double *observed;         // device [n_cols]; already divided by ref fold where applicable
    // outputs, device [n_cols]
    double *sum;                    // exact integer sum (as double) or float sum
    double *sumsq_dev;              // sum (x-mean)^2
    unsigned long long *n_lt;       // #{ x < obs } = insertion point of obs (gat/Engine.pyx:1549-1557)
    unsigned long long *n_eq;       // #{ x == obs }
    double *q_lo, *q_hi;            // order statistics at ranks rank_lo / rank_hi
    uint64_t rank_lo, rank_hi;
    const double *mean;             // device [n_cols] (second pass)
};
// gat-compare (scripts/gat-compare.py:218-241, :300-323): the sampled log ratio of two fold-change columns,
//   out[s][p] = log((obs1[p] / (m1[s][col1[p]] + pc) + 1e-4) / (obs2[p] / (m2[s][col2[p]] + pc) + 1e-4)) + delta[p]
// for pairs p0 <= p < p0 + n_pairs; out is [n_samples][n_pairs] float64, ready for the column statistics
struct CompareParams {
    const double *m1, *m2;          // [n_samples][n_cols1], [n_samples][n_cols2]
    uint32_t n_cols1, n_cols2;
    uint64_t n_samples;
    const int32_t *col1, *col2;     // [n_pairs] (already offset to the chunk)
    const double *obs1, *obs2, *delta;
    uint32_t n_pairs;
    double pseudo_count;
    double *out;
};
void launch_compare_derive(cudaStream_t st, const CompareParams &p);

void launch_stats_pass1(cudaStream_t st, const StatsParams &p);
void launch_stats_pass2(cudaStream_t st, const StatsParams &p);
void launch_stats_select(cudaStream_t st, const StatsParams &p);

}  // namespace gatb
