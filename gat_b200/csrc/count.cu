// count.cu -- K3/K4 counting kernel, construction of the annotation grid index, and K5 column statistics.  sm_100a.
//
// Counting: a CTA owns (group of tracks -- normally ALL tracks) x (chunk of samples) and keeps the chunk's
// count matrix [sample][track] in shared memory.  It walks the keys (contigs) in order; the warps of the
// CTA share the chunk's segment lists in blocks of 32 segments (lane = segment).  A lane turns its segment
// into the run of grid-index entries that can overlap it (count.cuh: two or three 4-byte loads), then the
// warp works the 32 runs off cooperatively, LPS lanes per run: one 8-byte + one 2-byte load per entry,
// the overlap test, and an integer atomic into the (sample, track) accumulator.  The entry loads of the
// next runs are issued before the current one is consumed (software pipeline in registers), which is what
// hides the L2 latency of this gather-bound kernel.  The index of one key (a few tens of MB at 1000 tracks)
// stays in L2 while the CTAs walk the keys in the same order.  The float64 nucleotide-density sum is formed
// per key from the integers, in the same key order and with the same compensated summation as the
// reference's Python sum() (gat/__init__.py:583-587).
#include <algorithm>
#include <cub/device/device_scan.cuh>
#include "count.cuh"
#include "../../include/gat_b200.h"

namespace gatb {

size_t count_smem_bytes(uint32_t schunk, uint32_t ka, bool density)
{
    return (size_t)schunk * ka * (density ? 20u : 4u) + 16u;
}

// ---------------------------------------------------------------------------------------------------
// One overlapping, de-duplicated pair: interval [x,y) of track slot t (previous interval of the track ends
// at pv) against segment [s,e) (previous segment of the list ends at pe).  Coordinates are < 2^31.
//   nucleotide-overlap   overlapWithSegments            gat/SegmentList.pyx:1026-1076
//   segment-overlap      intersectionWithSegments(base) :1078-1146  (a segment counts once per track:
//                        at the first interval of the track that overlaps it, i.e. pv <= s)
//   segment-midoverlap   midpoint tested against that FIRST overlapping interval only (:1137-1144)
//   annotation-*         roles swapped: an interval is counted by the first segment overlapping it,
//                        i.e. when it does not already overlap the previous segment (x >= pe)
//   overlap-pieces       len(a.intersect(b)): every overlapping pair is one piece (:1469-1549)
template <int COUNTER>
__device__ __forceinline__ void count_pair(uint32_t s, uint32_t e, uint32_t pe, uint32_t x, uint32_t y, uint32_t pv,
                                           uint32_t *__restrict__ cell)
{
    if (COUNTER == GATB_NUCLEOTIDE_OVERLAP) {
        atomicAdd(cell, min(e, y) - max(s, x));
    } else if (COUNTER == GATB_SEGMENT_OVERLAP) {
        if (pv <= s) atomicAdd(cell, 1u);
    } else if (COUNTER == GATB_SEGMENT_MIDOVERLAP) {
        const uint32_t mid = s + ((e - s) >> 1);
        if (pv <= s && x <= mid && mid < y) atomicAdd(cell, 1u);
    } else if (COUNTER == GATB_OVERLAP_PIECES) {
        atomicAdd(cell, 1u);
    } else if (COUNTER == GATB_ANNOTATION_OVERLAP) {
        if (x >= pe) atomicAdd(cell, 1u);
    } else {
        const uint32_t m = x + ((y - x) >> 1);
        if (x >= pe && s <= m && m < e) atomicAdd(cell, 1u);
    }
}

template <int COUNTER> struct NeedPrevInterval { static constexpr bool value = COUNTER == GATB_SEGMENT_OVERLAP || COUNTER == GATB_SEGMENT_MIDOVERLAP; };
template <int COUNTER> struct NeedPrevSegment { static constexpr bool value = COUNTER == GATB_ANNOTATION_OVERLAP || COUNTER == GATB_ANNOTATION_MIDOVERLAP; };

// one entry in flight: loaded ahead of its use
struct Flight {
    uint2 v;            // interval
    uint32_t t, pv;     // track slot | first flag; previous end of the track
    uint32_t j;         // entry index
    int pass;           // which run(s) of the block it belongs to; -1: none left
};

// One block of 32 segments (lane = segment i of sample slot `slot` on this key) against the grid index.
template <int COUNTER, int LPS, int DEPTH>
__device__ __forceinline__ void count_block(const CountParams &p, const uint32_t *__restrict__ boff, uint32_t nbins,
                                            uint32_t shift, const uint64_t *__restrict__ segs, uint32_t n, uint32_t b,
                                            int lane, uint32_t *__restrict__ acc)
{
    constexpr int SPP = 32 / LPS;                  // segments (runs) worked on per pass
    const uint32_t i = b * 32u + (uint32_t)lane;
    uint32_t s = 0, e = 0, r0 = 0, rf = 0, r1 = 0, pe = 0;
    if (i < n) {
        const uint64_t sg = segs[i];
        s = seg_start(sg); e = seg_end(sg);
        const uint32_t b0 = s >> shift;
        if (b0 < nbins) {
            const uint32_t b1 = min((e - 1u) >> shift, nbins - 1u);
            r0 = boff[b0]; rf = boff[b0 + 1u];
            r1 = (b1 == b0) ? rf : boff[b1 + 1u];
        }
    }
    if (NeedPrevSegment<COUNTER>::value) {
        pe = __shfl_up_sync(GATB_FULL, e, 1);
        if (lane == 0) pe = (i > 0 && i < n) ? seg_end(segs[i - 1u]) : 0u;
    }
    const uint32_t have = __ballot_sync(GATB_FULL, r1 > r0);
    if (!have) return;
    const int sub = lane & (LPS - 1), own = lane / LPS;

    int next_pass = 0;
    // issue the loads of the first entries of the next non-empty pass
    auto fetch = [&](Flight &f) {
        f.pass = -1;
        while (next_pass < LPS) {
            const uint32_t om = (SPP == 32) ? GATB_FULL : (((1u << SPP) - 1u) << (next_pass * SPP));
            if (have & om) break;
            next_pass++;
        }
        if (next_pass >= LPS) return;
        const int owner = next_pass * SPP + own;
        const uint32_t q0 = (LPS == 1) ? r0 : __shfl_sync(GATB_FULL, r0, owner);
        const uint32_t q1 = (LPS == 1) ? r1 : __shfl_sync(GATB_FULL, r1, owner);
        f.j = q0 + (uint32_t)sub;
        f.pass = next_pass++;
        if (f.j < q1) {
            f.v = p.civ[f.j];
            f.t = p.ctrk[f.j];
            if (NeedPrevInterval<COUNTER>::value) f.pv = p.cprev[f.j];
        }
    };
    auto consume = [&](Flight &f) {
        const int owner = f.pass * SPP + own;
        const uint32_t os = (LPS == 1) ? s : __shfl_sync(GATB_FULL, s, owner);
        const uint32_t oe = (LPS == 1) ? e : __shfl_sync(GATB_FULL, e, owner);
        const uint32_t of = (LPS == 1) ? rf : __shfl_sync(GATB_FULL, rf, owner);
        const uint32_t o1 = (LPS == 1) ? r1 : __shfl_sync(GATB_FULL, r1, owner);
        uint32_t ope = 0;
        if (NeedPrevSegment<COUNTER>::value) ope = (LPS == 1) ? pe : __shfl_sync(GATB_FULL, pe, owner);
        uint32_t j = f.j;
        uint2 v = f.v;
        uint32_t t = f.t, pv = NeedPrevInterval<COUNTER>::value ? f.pv : 0u;
        while (j < o1) {
            // overlapping, and met in the bin that holds the first base of the intersection
            if (v.x < oe && v.y > os && ((v.x >= os) ? (t & 0x8000u) != 0u : j < of))
                count_pair<COUNTER>(os, oe, ope, v.x, v.y, pv, acc + (t & 0x7fffu));
            j += LPS;
            if (j < o1) {                                   // runs longer than LPS: not prefetched
                v = p.civ[j];
                t = p.ctrk[j];
                if (NeedPrevInterval<COUNTER>::value) pv = p.cprev[j];
            }
        }
    };

    Flight fl[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; d++) fetch(fl[d]);
    while (true) {
#pragma unroll
        for (int d = 0; d < DEPTH; d++) {
            if (fl[d].pass < 0) return;                     // warp-uniform
            consume(fl[d]);
            fetch(fl[d]);
        }
    }
}

template <int COUNTER, bool DENSITY, int LPS, int DEPTH>
__global__ void __launch_bounds__(1024, 1) count_kernel(CountParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t g = blockIdx.x;
    const uint32_t a0 = g * p.ka;
    const uint32_t ka = min(p.ka, p.n_annot - a0);
    const uint32_t s_begin = blockIdx.y * p.schunk;
    const uint32_t s_end = min(s_begin + p.schunk, p.n_samples);
    const uint32_t ns = s_end - s_begin, ncells = ns * ka;
    // layout: [density only: (sum, compensation) doubles per cell][u32 per cell]
    double *acc_d = reinterpret_cast<double *>(smem);
    uint32_t *acc_u = reinterpret_cast<uint32_t *>(smem + (DENSITY ? (size_t)p.schunk * p.ka * 16u : 0u));

    for (uint32_t i = threadIdx.x; i < ncells; i += blockDim.x) {
        acc_u[i] = 0u;
        if (DENSITY) { acc_d[2 * i] = 0.0; acc_d[2 * i + 1] = 0.0; }
    }
    __syncthreads();

    for (uint32_t k = 0; k < p.n_keys; k++) {
        const KeyBins kb = p.keybins[(uint64_t)g * p.n_keys + k];
        if (kb.nbins == 0) continue;                      // no interval of any track on this key
        if (DENSITY && p.key_ws_nseg[k] == 0) continue;   // counter returns 0 (gat/Engine.pyx:1438-1440)
        const uint32_t *boff = p.boff + kb.base;
        const uint64_t *placed_key = p.placed + p.key_base[k];
        // the blocks of 32 segments of every sample of the chunk, dealt out to the warps round-robin
        for (uint32_t q0 = 0; q0 < ns; q0 += 32) {
            uint32_t my_n = 0;
            if (q0 + lane < ns) {
                const uint64_t sl = s_begin + q0 + lane;
                if (!p.key_present || p.key_present[sl * p.n_keys + k]) my_n = p.placed_n[sl * p.n_keys + k];
            }
            const uint32_t cnt = min(32u, ns - q0);
            for (uint32_t q = 0; q < cnt; q++) {
                const uint32_t n = __shfl_sync(GATB_FULL, my_n, q);
                const uint32_t slot = q0 + q;
                const uint64_t *segs = placed_key + (uint64_t)(s_begin + slot) * p.sample_stride;
                for (uint32_t b = ((uint32_t)warp + (uint32_t)nwarps - slot % (uint32_t)nwarps) % (uint32_t)nwarps;
                     b * 32u < n; b += nwarps)
                    count_block<COUNTER, LPS, DEPTH>(p, boff, kb.nbins, kb.shift, segs, n, b, lane, acc_u + slot * ka);
            }
        }
        if (DENSITY) {
            // per key: float(overlap) / len(workspace), accumulated in key order like the reference's
            // Python sum(): CPython >= 3.12 sums floats with Neumaier compensation (Python/bltinmodule.c)
            __syncthreads();
            const double den = (double)p.key_ws_nseg[k];
            for (uint32_t i = threadIdx.x; i < ncells; i += blockDim.x) {
                const uint32_t v = acc_u[i];
                if (v) {
                    const double x = (double)v / den, f = acc_d[2 * i], t = f + x;
                    acc_d[2 * i + 1] += (fabs(f) >= fabs(x)) ? ((f - t) + x) : ((x - t) + f);
                    acc_d[2 * i] = t;
                    acc_u[i] = 0u;
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < ncells; i += blockDim.x) {
        const uint64_t o = (uint64_t)(s_begin + i / ka) * p.n_annot + a0 + i % ka;
        if (DENSITY) {
            const double f = acc_d[2 * i], c = acc_d[2 * i + 1];
            p.out_f64[o] = (c != 0.0 && isfinite(c)) ? f + c : f;
        } else p.out_u32[o] = acc_u[i];
    }
}

template <int COUNTER, bool DENSITY, int LPS, int DEPTH>
static cudaError_t launch_count_t3(cudaStream_t st, const CountParams &p, int threads)
{
    const size_t smem = count_smem_bytes(p.schunk, p.ka, DENSITY);
    cudaError_t e = cudaFuncSetAttribute(count_kernel<COUNTER, DENSITY, LPS, DEPTH>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(p.n_groups, (p.n_samples + p.schunk - 1) / p.schunk);
    count_kernel<COUNTER, DENSITY, LPS, DEPTH><<<grid, threads, smem, st>>>(p);
    return cudaGetLastError();
}

template <int COUNTER, bool DENSITY, int LPS>
static cudaError_t launch_count_t2(cudaStream_t st, const CountParams &p, int threads)
{
    if (LPS == 1 || p.depth <= 1) return launch_count_t3<COUNTER, DENSITY, LPS, 1>(st, p, threads);
    if (p.depth == 2) return launch_count_t3<COUNTER, DENSITY, LPS, 2>(st, p, threads);
    return launch_count_t3<COUNTER, DENSITY, LPS, 4>(st, p, threads);
}

template <int COUNTER, bool DENSITY>
static cudaError_t launch_count_t(cudaStream_t st, const CountParams &p, int threads)
{
    switch (p.lps) {
    case 1:  return launch_count_t2<COUNTER, DENSITY, 1>(st, p, threads);
    case 4:  return launch_count_t2<COUNTER, DENSITY, 4>(st, p, threads);
    case 8:  return launch_count_t2<COUNTER, DENSITY, 8>(st, p, threads);
    case 16: return launch_count_t2<COUNTER, DENSITY, 16>(st, p, threads);
    default: return launch_count_t2<COUNTER, DENSITY, 32>(st, p, threads);
    }
}

cudaError_t launch_count(cudaStream_t st, int counter, const CountParams &p, int threads)
{
    if (p.n_samples == 0 || p.n_annot == 0) return cudaSuccess;
    switch (counter) {
    case GATB_NUCLEOTIDE_OVERLAP:    return launch_count_t<GATB_NUCLEOTIDE_OVERLAP, false>(st, p, threads);
    case GATB_NUCLEOTIDE_DENSITY:    return launch_count_t<GATB_NUCLEOTIDE_OVERLAP, true>(st, p, threads);
    case GATB_SEGMENT_OVERLAP:       return launch_count_t<GATB_SEGMENT_OVERLAP, false>(st, p, threads);
    case GATB_SEGMENT_MIDOVERLAP:    return launch_count_t<GATB_SEGMENT_MIDOVERLAP, false>(st, p, threads);
    case GATB_ANNOTATION_OVERLAP:    return launch_count_t<GATB_ANNOTATION_OVERLAP, false>(st, p, threads);
    case GATB_ANNOTATION_MIDOVERLAP: return launch_count_t<GATB_ANNOTATION_MIDOVERLAP, false>(st, p, threads);
    case GATB_OVERLAP_PIECES:        return launch_count_t<GATB_OVERLAP_PIECES, false>(st, p, threads);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------------
// Grid index construction (replaces a host loop over every interval).  Thread = interval, located in the
// CSR by a binary search over the list offsets, so skewed list sizes cost nothing.
//   1  bins_count_kernel: validate, and count the entries of every bin into boff[base + 1 + b]
//   2  exclusive scan of boff[] (cub): boff[base + 1 + b] = first entry of bin b, boff[base] = of the key
//   3  bins_fill_kernel: entry position = atomicAdd(boff[base + 1 + b], 1); afterwards boff[base + 1 + b]
//      is the END of bin b = the start of bin b + 1, i.e. boff[base + b] .. boff[base + b + 1] is bin b
__device__ __forceinline__ uint32_t find_list(const uint64_t *__restrict__ offs, uint32_t n_lists, uint64_t i)
{
    uint32_t lo = 0, hi = n_lists;                  // last l with offs[l] <= i
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (offs[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

template <bool FILL>
__global__ void __launch_bounds__(256) bins_pass_kernel(BuildBinsParams p)
{
    __shared__ unsigned long long s_total[8];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b0 = 1, b1 = 0, x = 0, y = 0, py = 0, t = 0;       // no bins unless the interval is valid
    uint32_t *cur = nullptr;
    if (i < p.n_intervals) {
        const uint32_t l = find_list(p.offs, p.n_annot * p.n_keys, i);
        const uint32_t a = l / p.n_keys, k = l % p.n_keys;
        const uint32_t g = a / p.ka;
        t = a % p.ka;
        const KeyBins kb = p.keybins[(uint64_t)g * p.n_keys + k];
        x = p.start[i]; y = p.end[i];
        py = (i > p.offs[l]) ? p.end[i - 1] : 0u;
        uint32_t err = 0;
        if (y >= 0x80000000u) err |= 1u;
        if (x >= y || py > x) err |= 2u;
        if (err) { if (!FILL) atomicOr(p.error, err); }
        else if (kb.nbins) {
            cur = p.boff + kb.base + 1u;
            b0 = min(x >> kb.shift, kb.nbins - 1u);
            b1 = min((y - 1u) >> kb.shift, kb.nbins - 1u);
        }
    }
    for (uint32_t b = b0; b <= b1; b++) {
        if (!FILL) atomicAdd(cur + b, 1u);
        else {
            const uint64_t pos = atomicAdd(cur + b, 1u);
            if (pos < p.capacity) {
                p.civ[pos] = make_uint2(x, y);
                p.ctrk[pos] = (uint16_t)(t | (b == b0 ? 0x8000u : 0u));
                p.cprev[pos] = py;
            }
        }
    }
    if (!FILL) {                                    // entries needed, in 64 bits: one atomic per CTA
        unsigned long long nb = (b1 >= b0) ? (unsigned long long)(b1 - b0 + 1u) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nb += __shfl_xor_sync(GATB_FULL, nb, d);
        if ((threadIdx.x & 31) == 0) s_total[threadIdx.x >> 5] = nb;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long sum = 0;
            for (int w = 0; w < 8; w++) sum += s_total[w];
            if (sum) atomicAdd(p.total, sum);
        }
    }
}

__global__ void bins_total_kernel(BuildBinsParams p)
{
    if (*p.total > p.capacity) atomicOr(p.error, 4u);
}

size_t build_bins_scan_bytes(uint64_t n_boff)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum((void *)nullptr, bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_boff);
    return bytes;
}

cudaError_t launch_build_bins(cudaStream_t st, const BuildBinsParams &p, void *scan_tmp, size_t scan_bytes)
{
    if (p.n_intervals == 0 || p.n_boff == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((p.n_intervals + 255) / 256);
    bins_pass_kernel<false><<<blocks, 256, 0, st>>>(p);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, p.boff, p.boff, (int)p.n_boff, st);
    if (e != cudaSuccess) return e;
    bins_total_kernel<<<1, 1, 0, st>>>(p);
    bins_pass_kernel<true><<<blocks, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// K5 column statistics (gat/Engine.pyx:1635-1718, :1543-1576).  One CTA per column.
template <typename T>
__device__ __forceinline__ T block_reduce_sum(T v, T *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(GATB_FULL, v, d);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    T t = (threadIdx.x < (unsigned)nw) ? scratch[threadIdx.x] : (T)0;
    if (warp == 0) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(GATB_FULL, t, d);
    }
    return t;   // valid in thread 0
}

__device__ __forceinline__ double load_val(const StatsParams &p, uint64_t s, uint32_t col)
{
    if (p.is_float) return reinterpret_cast<const double *>(p.counts)[s * p.n_cols + col];
    return (double)reinterpret_cast<const uint32_t *>(p.counts)[s * p.n_cols + col];
}

__global__ void __launch_bounds__(256) stats_pass1_kernel(StatsParams p)
{
    __shared__ unsigned long long su[32];
    __shared__ double sd[32];
    const uint32_t col = blockIdx.x;
    const double obs = p.observed[col];
    unsigned long long isum = 0, nlt = 0, neq = 0;
    double fsum = 0.0;
    for (uint64_t s = threadIdx.x; s < p.n_samples; s += blockDim.x) {
        double x;
        if (p.is_float) { x = reinterpret_cast<const double *>(p.counts)[s * p.n_cols + col]; fsum += x; }
        else { uint32_t v = reinterpret_cast<const uint32_t *>(p.counts)[s * p.n_cols + col]; isum += v; x = (double)v; }
        nlt += (x < obs) ? 1ull : 0ull;      // searchargsorted + cmpDouble (gat/Engine.pyx:122-127, 1549-1557)
        neq += (x == obs) ? 1ull : 0ull;
    }
    unsigned long long r;
    r = block_reduce_sum(nlt, su); if (threadIdx.x == 0) p.n_lt[col] = r;
    r = block_reduce_sum(neq, su); if (threadIdx.x == 0) p.n_eq[col] = r;
    if (p.is_float) { double t = block_reduce_sum(fsum, sd); if (threadIdx.x == 0) p.sum[col] = t; }
    else { r = block_reduce_sum(isum, su); if (threadIdx.x == 0) p.sum[col] = (double)r; }
}

__global__ void __launch_bounds__(256) stats_pass2_kernel(StatsParams p)
{
    __shared__ double sd[32];
    const uint32_t col = blockIdx.x;
    const double mean = p.mean[col];
    double acc = 0.0;
    for (uint64_t s = threadIdx.x; s < p.n_samples; s += blockDim.x) {
        double d = load_val(p, s, col) - mean;
        acc += d * d;
    }
    double t = block_reduce_sum(acc, sd);
    if (threadIdx.x == 0) p.sumsq_dev[col] = t;
}

// order-preserving 64-bit key of a value
__device__ __forceinline__ unsigned long long val_key(const StatsParams &p, uint64_t s, uint32_t col)
{
    if (!p.is_float) return (unsigned long long)reinterpret_cast<const uint32_t *>(p.counts)[s * p.n_cols + col];
    unsigned long long b = (unsigned long long)__double_as_longlong(reinterpret_cast<const double *>(p.counts)[s * p.n_cols + col]);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double key_val(int is_float, unsigned long long k)
{
    if (!is_float) return (double)k;
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// radix select of two order statistics per column: 8 bits per pass from the top byte
__global__ void __launch_bounds__(256) stats_select_kernel(StatsParams p)
{
    __shared__ unsigned int hist[2][256];
    __shared__ unsigned long long prefix_s[2];
    __shared__ unsigned long long rank_s[2];
    const uint32_t col = blockIdx.x;
    const int top = p.is_float ? 56 : 24;
    if (threadIdx.x == 0) { prefix_s[0] = prefix_s[1] = 0; rank_s[0] = p.rank_lo; rank_s[1] = p.rank_hi; }
    __syncthreads();
    for (int shift = top; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 512; i += blockDim.x) (&hist[0][0])[i] = 0;
        __syncthreads();
        const unsigned long long pre0 = prefix_s[0], pre1 = prefix_s[1];
        const unsigned long long hmask = (shift == 56) ? 0ull : (~0ull << (shift + 8));
        for (uint64_t s = threadIdx.x; s < p.n_samples; s += blockDim.x) {
            unsigned long long k = val_key(p, s, col);
            unsigned int b = (unsigned int)(k >> shift) & 255u;
            if ((k & hmask) == pre0) atomicAdd(&hist[0][b], 1u);
            if ((k & hmask) == pre1) atomicAdd(&hist[1][b], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            const int w = threadIdx.x;
            unsigned long long r = rank_s[w], cum = 0;
            int b = 0;
            for (; b < 256; b++) { if (cum + hist[w][b] > r) break; cum += hist[w][b]; }
            if (b > 255) b = 255;
            rank_s[w] = r - cum;
            prefix_s[w] |= ((unsigned long long)b << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.q_lo[col] = key_val(p.is_float, prefix_s[0]);
        p.q_hi[col] = key_val(p.is_float, prefix_s[1]);
    }
}

__global__ void __launch_bounds__(256) compare_derive_kernel(CompareParams p)
{
    const uint64_t total = p.n_samples * p.n_pairs;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = i / p.n_pairs;
        const uint32_t q = (uint32_t)(i % p.n_pairs);
        // same operation order as the numpy expressions: divide, add the fold pseudo-count, divide, log, shift
        const double f1 = p.obs1[q] / (p.m1[s * p.n_cols1 + (uint32_t)p.col1[q]] + p.pseudo_count) + 0.0001;
        const double f2 = p.obs2[q] / (p.m2[s * p.n_cols2 + (uint32_t)p.col2[q]] + p.pseudo_count) + 0.0001;
        p.out[i] = log(f1 / f2) + p.delta[q];
    }
}

void launch_compare_derive(cudaStream_t st, const CompareParams &p)
{
    const uint64_t total = p.n_samples * p.n_pairs;
    if (total == 0) return;
    const unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, 148u * 16u);
    compare_derive_kernel<<<blocks, 256, 0, st>>>(p);
}

void launch_stats_pass1(cudaStream_t st, const StatsParams &p)
{
    if (p.n_cols) stats_pass1_kernel<<<p.n_cols, 256, 0, st>>>(p);
}
void launch_stats_pass2(cudaStream_t st, const StatsParams &p)
{
    if (p.n_cols) stats_pass2_kernel<<<p.n_cols, 256, 0, st>>>(p);
}
void launch_stats_select(cudaStream_t st, const StatsParams &p)
{
    if (p.n_cols) stats_select_kernel<<<p.n_cols, 256, 0, st>>>(p);
}

}  // namespace gatb
