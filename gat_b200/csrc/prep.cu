// prep.cu -- input preparation on the GPU: interval lists live on the device as CSR
// (gatb_lists) from the moment the parsed BED columns arrive until the annotation index is built from them.
//
//   gatb_lists_from_rows     IntervalCollection.normalize()  (gat/Engine.pyx:2941-2956 -> SegmentList.normalize,
//                            gat/SegmentList.pyx:697-754): rows (list, start, end) in file order -> sorted,
//                            normalized lists.  One radix sort by (list, start), one max-scan, one compaction.
//   gatb_lists_restrict      IntervalCollection.intersect / filter (gat/SegmentList.pyx:1469-1549, :1401-1467) and
//                            toIsochores (gat/Engine.pyx:2837-2855; gat/IO.py:188-293): every list against `fanout`
//                            other lists of its key (the workspace: fanout 1; the isochore tracks: fanout = their
//                            number), truncating (intersect) or not (filter).  Binary search per interval, count,
//                            scan, emit.
//   gatb_lists_collapse      IntervalDictionary.fromIsochores (gat/Engine.pyx:2857-2876): the lists of a key's
//                            isochores extended into one list, then merge(0) (adjacent segments joined).
//   gatb_lists_select        re-order / subset lists (the key order of another dictionary)
// All coordinates are < 2^31; lists hold < 2^32 intervals in total.
#include "../../include/gat_b200.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <string>
#include <vector>

#include "common.cuh"
#include "prep.cuh"

namespace gatb {

// ---- kernels ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rows_key_kernel(const uint32_t *__restrict__ list_id, const uint32_t *__restrict__ start,
                                                       const uint32_t *__restrict__ end, uint64_t n, uint32_t n_lists,
                                                       uint64_t *__restrict__ key, uint32_t *__restrict__ val, uint32_t *error)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t l = list_id[i], s = start[i], e = end[i];
    if (l >= n_lists) atomicOr(error, 8u);
    if (e >= 0x80000000u || s >= 0x80000000u) atomicOr(error, 1u);
    if (s > e) atomicOr(error, 2u);
    key[i] = ((uint64_t)l << 32) | s;
    val[i] = e;
}

// packed (list, end + 1) of a non-empty row, (list, 0) of an empty one: rows are sorted by list, so a plain
// max-scan of the packed values is the running max end SINCE THE START OF THE ROW'S LIST
__global__ void __launch_bounds__(256) rows_pack_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val,
                                                        uint64_t n, uint64_t *__restrict__ packed)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = (uint32_t)key[i], e = val[i];
    packed[i] = (key[i] & 0xffffffff00000000ull) | (e != s ? e + 1u : 0u);
}

// head of a merged segment: a non-empty row that the segments before it (same list) do not reach.
// SegmentList.normalize opens a new segment when start >= max_end (adjacent segments stay apart,
// gat/SegmentList.pyx:735-745); merge(0) when start > max_end (adjacent ones join, :793-806)
__global__ void __launch_bounds__(256) rows_head_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val,
                                                        const uint64_t *__restrict__ runmax, uint64_t n, int join_adjacent,
                                                        uint32_t *__restrict__ head)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = (uint32_t)key[i], e = val[i];
    uint32_t h = 0;
    if (e != s) {
        int64_t prev = -1;                          // running max end of the list's earlier rows, -1: none
        if (i > 0 && (runmax[i - 1] >> 32) == (key[i] >> 32)) prev = (int64_t)(uint32_t)runmax[i - 1] - 1;
        h = join_adjacent ? ((int64_t)s > prev) : ((int64_t)s >= prev);
    }
    head[i] = h;
}

// compaction: a head writes its start, every non-empty row raises the end of its group
__global__ void __launch_bounds__(256) rows_emit_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val,
                                                        const uint32_t *__restrict__ head, const uint32_t *__restrict__ head_excl,
                                                        uint64_t n, uint32_t *__restrict__ out_start, uint32_t *__restrict__ out_end,
                                                        uint32_t *__restrict__ out_list)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = (uint32_t)key[i], e = val[i];
    if (e == s) return;
    const uint32_t g = head_excl[i] + head[i] - 1u;        // group = heads up to and including this row, - 1
    if (head[i]) { out_start[g] = s; out_list[g] = (uint32_t)(key[i] >> 32); }
    atomicMax(out_end + g, e);
}

// offs[l] = merged segments of the lists before l = heads before the first row of list l
__global__ void __launch_bounds__(256) rows_offs_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ head_excl,
                                                        const uint32_t *__restrict__ head, uint64_t n, uint32_t n_lists,
                                                        uint64_t *__restrict__ offs)
{
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > n_lists) return;
    uint64_t lo = 0, hi = n;                        // first row with list >= l
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if ((uint32_t)(key[mid] >> 32) < l) lo = mid + 1; else hi = mid;
    }
    offs[l] = lo < n ? head_excl[lo] : (n ? head_excl[n - 1] + head[n - 1] : 0u);
}

// list of every interval of a CSR (binary search in the offsets)
__global__ void __launch_bounds__(256) csr_list_kernel(const uint64_t *__restrict__ offs, uint32_t n_lists, uint64_t n,
                                                       uint32_t *__restrict__ list_of)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t lo = 0, hi = n_lists;                  // last l with offs[l] <= i
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (offs[mid] <= i) lo = mid; else hi = mid;
    }
    list_of[i] = lo;
}

// pieces of `other` list o that overlap [s, e): [first, last)
__device__ __forceinline__ void overlapping_pieces(const uint32_t *__restrict__ os, const uint32_t *__restrict__ oe,
                                                   uint64_t b, uint64_t e_, uint32_t s, uint32_t e, uint64_t *first, uint64_t *last)
{
    uint64_t lo = b, hi = e_;                       // first piece with end > s
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (oe[mid] > s) hi = mid; else lo = mid + 1;
    }
    *first = lo;
    hi = e_;                                        // first piece with start >= e
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (os[mid] >= e) hi = mid; else lo = mid + 1;
    }
    *last = lo;
}

// slot of (interval i of list l, fan f) in output order: lists (l, f) ascending, intervals in input order
__device__ __forceinline__ uint64_t fan_slot(const uint64_t *__restrict__ in_offs, uint32_t l, uint32_t f, uint32_t fanout, uint64_t i)
{
    const uint64_t b = in_offs[l], len = in_offs[l + 1] - b;
    return b * fanout + (uint64_t)f * len + (i - b);
}

template <bool EMIT>
__global__ void __launch_bounds__(256) restrict_kernel(RestrictParams p)
{
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= p.n * p.fanout) return;
    const uint64_t i = id / p.fanout;
    const uint32_t f = (uint32_t)(id % p.fanout);
    const uint32_t l = p.in_list[i], s = p.in_start[i], e = p.in_end[i];
    const uint32_t o = (l % p.n_keys) * p.fanout + f;
    uint64_t first, last;
    overlapping_pieces(p.o_start, p.o_end, p.o_offs[o], p.o_offs[o + 1], s, e, &first, &last);
    const uint64_t slot = fan_slot(p.in_offs, l, f, p.fanout, i);
    const uint32_t cnt = p.truncate ? (uint32_t)(last - first) : (last > first ? 1u : 0u);
    if (!EMIT) { p.cnt[slot] = cnt; return; }
    uint64_t w = p.cnt_excl[slot];
    if (!p.truncate) {
        if (cnt) { p.out_start[w] = s; p.out_end[w] = e; p.out_list[w] = l * p.fanout + f; }
        return;
    }
    for (uint64_t q = first; q < last; q++, w++) {
        p.out_start[w] = max(s, p.o_start[q]);
        p.out_end[w] = min(e, p.o_end[q]);
        p.out_list[w] = l * p.fanout + f;
    }
}

// offsets of the output lists (l, f): the scanned count at the list's first slot
__global__ void __launch_bounds__(256) restrict_offs_kernel(RestrictParams p, uint32_t n_in_lists, uint64_t total)
{
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_out = (uint64_t)n_in_lists * p.fanout;
    if (id > n_out) return;
    if (id == n_out) { p.out_offs[id] = total; return; }
    const uint32_t l = (uint32_t)(id / p.fanout), f = (uint32_t)(id % p.fanout);
    const uint64_t b = p.in_offs[l], len = p.in_offs[l + 1] - b;
    const uint64_t slot = b * p.fanout + (uint64_t)f * len;
    p.out_offs[id] = slot < p.n * p.fanout ? p.cnt_excl[slot] : total;
}

__global__ void __launch_bounds__(256) collapse_key_kernel(const uint32_t *__restrict__ in_list, const uint32_t *__restrict__ start,
                                                           const uint32_t *__restrict__ end, uint64_t n, uint32_t fanout,
                                                           uint64_t *__restrict__ key, uint32_t *__restrict__ val)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = ((uint64_t)(in_list[i] / fanout) << 32) | start[i];
    val[i] = end[i];
}

__global__ void __launch_bounds__(256) select_count_kernel(const uint64_t *__restrict__ in_offs, const uint32_t *__restrict__ src,
                                                           uint32_t n_out, uint32_t n_in, unsigned long long *__restrict__ len)
{
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > n_out) return;
    len[l] = (l < n_out && src[l] < n_in) ? in_offs[src[l] + 1] - in_offs[src[l]] : 0ull;
}

__global__ void __launch_bounds__(256) select_copy_kernel(const uint64_t *__restrict__ in_offs, const uint32_t *__restrict__ in_start,
                                                          const uint32_t *__restrict__ in_end, const uint32_t *__restrict__ src,
                                                          uint32_t l0, uint32_t n_in, const unsigned long long *__restrict__ out_offs,
                                                          uint32_t *__restrict__ out_start, uint32_t *__restrict__ out_end,
                                                          uint32_t *__restrict__ out_list)
{
    const uint32_t l = l0 + blockIdx.y;             // one grid row per output list
    if (src[l] >= n_in) return;
    const uint64_t b = in_offs[src[l]], len = in_offs[src[l] + 1] - b, o = out_offs[l];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
        out_start[o + i] = in_start[b + i];
        out_end[o + i] = in_end[b + i];
        out_list[o + i] = l;
    }
}

__global__ void __launch_bounds__(256) sizes_kernel(const uint32_t *__restrict__ list_of, const uint32_t *__restrict__ start,
                                                    const uint32_t *__restrict__ end, uint64_t n, unsigned long long *__restrict__ bases)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(bases + list_of[i], (unsigned long long)(end[i] - start[i]));
}

__global__ void __launch_bounds__(256) last_end_kernel(const uint64_t *__restrict__ offs, const uint32_t *__restrict__ end,
                                                       uint32_t n_lists, uint32_t *__restrict__ last_end)
{
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n_lists) last_end[l] = offs[l + 1] > offs[l] ? end[offs[l + 1] - 1] : 0u;
}

// ---- launch wrappers -------------------------------------------------------------------------------------------
static inline unsigned nblocks(uint64_t n) { return (unsigned)((n + 255) / 256); }

void launch_rows_key(cudaStream_t st, const uint32_t *l, const uint32_t *s, const uint32_t *e, uint64_t n, uint32_t n_lists,
                     uint64_t *key, uint32_t *val, uint32_t *error)
{
    if (n) rows_key_kernel<<<nblocks(n), 256, 0, st>>>(l, s, e, n, n_lists, key, val, error);
}
void launch_collapse_key(cudaStream_t st, const uint32_t *in_list, const uint32_t *s, const uint32_t *e, uint64_t n, uint32_t fanout,
                         uint64_t *key, uint32_t *val)
{
    if (n) collapse_key_kernel<<<nblocks(n), 256, 0, st>>>(in_list, s, e, n, fanout, key, val);
}
void launch_csr_list(cudaStream_t st, const uint64_t *offs, uint32_t n_lists, uint64_t n, uint32_t *list_of)
{
    if (n) csr_list_kernel<<<nblocks(n), 256, 0, st>>>(offs, n_lists, n, list_of);
}
void launch_sizes(cudaStream_t st, const uint32_t *list_of, const uint32_t *s, const uint32_t *e, uint64_t n, unsigned long long *bases)
{
    if (n) sizes_kernel<<<nblocks(n), 256, 0, st>>>(list_of, s, e, n, bases);
}
void launch_last_end(cudaStream_t st, const uint64_t *offs, const uint32_t *end, uint32_t n_lists, uint32_t *last_end)
{
    if (n_lists) last_end_kernel<<<nblocks(n_lists), 256, 0, st>>>(offs, end, n_lists, last_end);
}

size_t sort_temp_bytes(uint64_t n, int end_bit)
{
    size_t bytes = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int)n, 0, end_bit);
    size_t scan = 0;
    cub::DeviceScan::InclusiveScan((void *)nullptr, scan, (uint64_t *)nullptr, (uint64_t *)nullptr, cub::Max(), (int)n);
    size_t sum = 0;
    cub::DeviceScan::ExclusiveSum((void *)nullptr, sum, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    return std::max(bytes, std::max(scan, sum));
}

// sorted (key = list << 32 | start, val = end) rows -> heads, exclusive head sums (in `head_excl`), *n_out on the host
cudaError_t merge_sorted_rows(cudaStream_t st, MergeRows &m)
{
    const uint64_t n = m.n;
    cub::DoubleBuffer<uint64_t> k(m.key, m.key_alt);
    cub::DoubleBuffer<uint32_t> v(m.val, m.val_alt);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(m.temp, m.temp_bytes, k, v, (int)n, 0, m.end_bit, st);
    if (e != cudaSuccess) return e;
    m.key = k.Current(); m.key_alt = k.Alternate();
    m.val = v.Current(); m.val_alt = v.Alternate();
    uint64_t *packed = m.key_alt;                    // (the sort's spare buffers are free again)
    rows_pack_kernel<<<nblocks(n), 256, 0, st>>>(m.key, m.val, n, packed);
    e = cub::DeviceScan::InclusiveScan(m.temp, m.temp_bytes, packed, packed, cub::Max(), (int)n, st);
    if (e != cudaSuccess) return e;
    rows_head_kernel<<<nblocks(n), 256, 0, st>>>(m.key, m.val, packed, n, m.join_adjacent, m.head);
    e = cub::DeviceScan::ExclusiveSum(m.temp, m.temp_bytes, m.head, m.head_excl, (int)n, st);
    return e != cudaSuccess ? e : cudaGetLastError();
}

void launch_rows_emit(cudaStream_t st, const MergeRows &m, uint32_t n_lists, uint64_t *offs, uint32_t *out_start, uint32_t *out_end,
                      uint32_t *out_list)
{
    if (m.n) rows_emit_kernel<<<nblocks(m.n), 256, 0, st>>>(m.key, m.val, m.head, m.head_excl, m.n, out_start, out_end, out_list);
    rows_offs_kernel<<<nblocks((uint64_t)n_lists + 1), 256, 0, st>>>(m.key, m.head_excl, m.head, m.n, n_lists, offs);
}

cudaError_t launch_restrict(cudaStream_t st, const RestrictParams &p, bool emit)
{
    const uint64_t n = p.n * p.fanout;
    if (n == 0) return cudaSuccess;
    if (emit) restrict_kernel<true><<<nblocks(n), 256, 0, st>>>(p);
    else restrict_kernel<false><<<nblocks(n), 256, 0, st>>>(p);
    return cudaGetLastError();
}

void launch_restrict_offs(cudaStream_t st, const RestrictParams &p, uint32_t n_in_lists, uint64_t total)
{
    restrict_offs_kernel<<<nblocks((uint64_t)n_in_lists * p.fanout + 1), 256, 0, st>>>(p, n_in_lists, total);
}

void launch_select_count(cudaStream_t st, const uint64_t *in_offs, const uint32_t *src, uint32_t n_out, uint32_t n_in, unsigned long long *len)
{
    select_count_kernel<<<nblocks((uint64_t)n_out + 1), 256, 0, st>>>(in_offs, src, n_out, n_in, len);
}

void launch_select_copy(cudaStream_t st, const uint64_t *in_offs, const uint32_t *in_start, const uint32_t *in_end, const uint32_t *src,
                        uint32_t n_out, uint32_t n_in, const unsigned long long *out_offs, uint32_t *out_start, uint32_t *out_end,
                        uint32_t *out_list)
{
    if (n_out == 0) return;
    // lists are a few hundred to a few thousand intervals: 4 CTAs per list; the y dimension holds <= 65535 lists a launch
    for (uint32_t l0 = 0; l0 < n_out; l0 += 65535u) {
        const dim3 grid(4, std::min(65535u, n_out - l0));
        select_copy_kernel<<<grid, 256, 0, st>>>(in_offs, in_start, in_end, src, l0, n_in, out_offs, out_start, out_end, out_list);
    }
}

size_t exclusive_sum_bytes(uint64_t n, bool u64)
{
    size_t b = 0;
    if (u64) cub::DeviceScan::ExclusiveSum((void *)nullptr, b, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n);
    else cub::DeviceScan::ExclusiveSum((void *)nullptr, b, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    return b;
}
cudaError_t exclusive_sum_u32(cudaStream_t st, void *temp, size_t bytes, const uint32_t *in, uint32_t *out, uint64_t n)
{
    return cub::DeviceScan::ExclusiveSum(temp, bytes, in, out, (int)n, st);
}
cudaError_t exclusive_sum_u64(cudaStream_t st, void *temp, size_t bytes, const unsigned long long *in, unsigned long long *out, uint64_t n)
{
    return cub::DeviceScan::ExclusiveSum(temp, bytes, in, out, (int)n, st);
}

}  // namespace gatb
