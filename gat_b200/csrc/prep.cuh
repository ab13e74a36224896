// prep.cuh -- device-side input preparation (prep.cu): normalize, intersect / filter, isochore split and merge of
// interval lists held on the GPU as CSR.  Replaces the host loops of IntervalCollection.normalize / intersect /
// toIsochores / fromIsochores (gat/Engine.pyx:2837-2876, 2941-2956; gat/IO.py:188-293) for large collections.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gatb {

// rows being sorted and merged: key = list << 32 | start, val = end (double-buffered for the radix sort)
struct MergeRows {
    uint64_t *key, *key_alt;
    uint32_t *val, *val_alt;
    uint32_t *head, *head_excl;     // [n]
    void *temp;
    size_t temp_bytes;
    uint64_t n;
    int end_bit;                    // key bits that matter: 32 + bits of the largest list id
    int join_adjacent;              // 0: SegmentList.normalize, 1: merge(0)
};

struct RestrictParams {
    // input lists
    const uint64_t *in_offs;
    const uint32_t *in_start, *in_end, *in_list;
    uint64_t n;
    // the lists restricted against: list (key, f) = o_offs[key * fanout + f]
    const uint64_t *o_offs;
    const uint32_t *o_start, *o_end;
    uint32_t n_keys, fanout;
    int truncate;                   // 1: intersect (pieces), 0: filter (whole intervals with >= 1 base in common)
    // pass 1 / 2
    uint32_t *cnt;                  // [n * fanout] pieces per (interval, fan), in output order
    const uint32_t *cnt_excl;       // its exclusive prefix sums
    uint64_t *out_offs;
    uint32_t *out_start, *out_end, *out_list;
};

size_t sort_temp_bytes(uint64_t n, int end_bit);
void launch_rows_key(cudaStream_t st, const uint32_t *l, const uint32_t *s, const uint32_t *e, uint64_t n, uint32_t n_lists,
                     uint64_t *key, uint32_t *val, uint32_t *error);
void launch_collapse_key(cudaStream_t st, const uint32_t *in_list, const uint32_t *s, const uint32_t *e, uint64_t n, uint32_t fanout,
                         uint64_t *key, uint32_t *val);
cudaError_t merge_sorted_rows(cudaStream_t st, MergeRows &m);
void launch_rows_emit(cudaStream_t st, const MergeRows &m, uint32_t n_lists, uint64_t *offs, uint32_t *out_start, uint32_t *out_end,
                      uint32_t *out_list);
void launch_csr_list(cudaStream_t st, const uint64_t *offs, uint32_t n_lists, uint64_t n, uint32_t *list_of);
cudaError_t launch_restrict(cudaStream_t st, const RestrictParams &p, bool emit);
void launch_restrict_offs(cudaStream_t st, const RestrictParams &p, uint32_t n_in_lists, uint64_t total);
void launch_select_count(cudaStream_t st, const uint64_t *in_offs, const uint32_t *src, uint32_t n_out, uint32_t n_in, unsigned long long *len);
void launch_select_copy(cudaStream_t st, const uint64_t *in_offs, const uint32_t *in_start, const uint32_t *in_end, const uint32_t *src,
                        uint32_t n_out, uint32_t n_in, const unsigned long long *out_offs, uint32_t *out_start, uint32_t *out_end,
                        uint32_t *out_list);
void launch_sizes(cudaStream_t st, const uint32_t *list_of, const uint32_t *s, const uint32_t *e, uint64_t n, unsigned long long *bases);
void launch_last_end(cudaStream_t st, const uint64_t *offs, const uint32_t *end, uint32_t n_lists, uint32_t *last_end);

size_t exclusive_sum_bytes(uint64_t n, bool u64);
cudaError_t exclusive_sum_u32(cudaStream_t st, void *temp, size_t bytes, const uint32_t *in, uint32_t *out, uint64_t n);
cudaError_t exclusive_sum_u64(cudaStream_t st, void *temp, size_t bytes, const unsigned long long *in, unsigned long long *out, uint64_t n);

}  // namespace gatb
