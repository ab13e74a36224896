// count.cuh -- counting kernel (K3/K4) and the annotation tile format.
//
// Replaces overlapWithSegments / intersectionWithSegments (gat/SegmentList.pyx:1026-1146) as used by
// the Counter* classes (gat/Engine.pyx:1412-1472) inside computeSample (gat/__init__.py:580-587) and
// computeCounts (gat/Engine.pyx:2189-2202).
#pragma once

#include "common.cuh"

namespace gatb {

constexpr int KMAX = 8;                 // annotation tracks per shared-memory tile (register accumulators)

// Tile = the intervals of up to KMAX annotation tracks on ONE key (contig), contiguous in global
// memory so that one cooperative (or bulk) copy stages it in shared memory:
//   TileHeader | uint4 idx8[nbins+1] | for each of the KMAX slots: uint2 iv[n+2]
// iv ends with two sentinels (INT_MAX, INT_MAX); unused slots (tile with fewer than KMAX tracks) hold
// only the sentinels.  idx8[b] packs, for all 8 slots, the uint16 index of the first interval whose
// end is > (b << shift): a one-probe replacement for the binary search (utils/gat_utils.c:8-32) over
// the sorted interval ends, fetched for 8 tracks with a single 16-byte load.  The bin width
// (1 << shift) is common to the tile.  nbins == 0: no index (a track with > 65534 intervals); the
// kernel then binary-searches the intervals.
struct TileHeader {
    uint32_t iv_off[KMAX];      // byte offsets from the tile start
    uint32_t n[KMAX];
    uint32_t idx_off;
    uint32_t nbins;
    uint32_t shift;
    uint32_t pad;
};

struct CountParams {
    // annotations
    const uint8_t *tiles;           // tile blob
    const uint64_t *tile_off;       // [n_groups][n_keys] byte offset
    const uint32_t *tile_bytes;     // [n_groups][n_keys]
    const uint32_t *key_ws_nseg;    // [n_keys] or NULL
    uint32_t n_annot, n_keys, n_groups, ka;   // ka = tracks per group
    uint32_t smem_tile_budget;      // tiles up to this many bytes are staged in shared memory
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

// Builds every tile from the raw annotation CSR arrays on the device (one CTA per tile): copies the
// intervals, writes sentinels and the interleaved bin index, and validates the lists (error bit 0:
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
    unsigned long long *n_trunc_lt; // insertion point of obs = #{ x < obs } (lower_bound, gat/Engine.pyx:1549-1557)
    unsigned long long *n_lt;       // #{ x < obs }
    unsigned long long *n_eq;       // #{ x == obs }
    double *q_lo, *q_hi;            // order statistics at ranks rank_lo / rank_hi
    uint64_t rank_lo, rank_hi;
    const double *mean;             // device [n_cols] (second pass)
};
void launch_stats_pass1(cudaStream_t st, const StatsParams &p);
void launch_stats_pass2(cudaStream_t st, const StatsParams &p);
void launch_stats_select(cudaStream_t st, const StatsParams &p, uint64_t *scratch_keys);

}  // namespace gatb
