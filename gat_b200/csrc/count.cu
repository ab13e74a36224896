// count.cu -- K3/K4 counting kernel and K5 column statistics.  sm_100a.
//
// Counting: a CTA owns (group of <= KMAX annotation tracks) x (chunk of samples) and walks the keys
// (contigs) in order.  Per key it stages the group's tile (intervals + bin index) in shared memory,
// then every warp streams its samples' segments on that key through the tile: lane = segment,
// KMAX independent lookups per lane (register accumulators, ILP across tracks), warp `redux` per
// track, and a shared-memory accumulator per (sample, track) that is only ever touched by the
// owning warp -- no atomics, deterministic, and the float64 nucleotide-density sum runs in the same
// key order as the reference's Python sum() (gat/__init__.py:583-587).
#include "count.cuh"
#include "../../include/gat_b200.h"

namespace gatb {

// ---------------------------------------------------------------------------------------------------
// Exact count of one segment [s,e) against one annotation track, starting the scan at interval j
// (any j not past the first interval whose end is > s).  pe = end of the previous segment of the list.
// Coordinates are < 2^31, so signed compares are exact and the two sentinels are (INT_MAX, INT_MAX).
template <int COUNTER>
__device__ __forceinline__ uint32_t scan_from(const uint2 *__restrict__ iv, uint32_t j, int s, int e, int pe)
{
    uint2 a = iv[j];
    while ((int)a.y <= s) a = iv[++j];          // first interval with end > s; the sentinel stops the scan
    uint32_t r = 0;
    if (COUNTER == GATB_SEGMENT_OVERLAP) {
        // intersectionWithSegments(base), gat/SegmentList.pyx:1078-1146: the segment counts once
        r = ((int)a.x < e) ? 1u : 0u;
    } else if (COUNTER == GATB_SEGMENT_MIDOVERLAP) {
        // midpoint tested against the FIRST overlapping interval only (:1137-1144)
        int mid = s + ((e - s) >> 1);
        r = ((int)a.x < e && (int)a.x <= mid && mid < (int)a.y) ? 1u : 0u;
    } else if (COUNTER == GATB_NUCLEOTIDE_OVERLAP) {
        // overlapWithSegments, gat/SegmentList.pyx:1026-1076
        while ((int)a.x < e) { r += (uint32_t)(min(e, (int)a.y) - max(s, (int)a.x)); a = iv[++j]; }
    } else if (COUNTER == GATB_ANNOTATION_OVERLAP) {
        // roles swapped: an interval is counted by the first segment overlapping it, i.e. when it does
        // not already overlap the previous segment (start >= pe)
        while ((int)a.x < e) { r += ((int)a.x >= pe) ? 1u : 0u; a = iv[++j]; }
    } else {  // GATB_ANNOTATION_MIDOVERLAP
        while ((int)a.x < e) {
            if ((int)a.x >= pe) { int m = (int)a.x + (((int)a.y - (int)a.x) >> 1); r += (s <= m && m < e) ? 1u : 0u; }
            a = iv[++j];
        }
    }
    return r;
}

// Deferred exact scans.  A (segment, track) pair whose count the branch-free fast path cannot prove
// (about 2 % of the pairs) is pushed on a per-warp queue in shared memory instead of being resolved
// on the spot with 2-3 of 32 lanes active; whenever 32 entries are waiting the warp resolves them
// with all lanes busy.  The queue is drained at the end of every sample.
constexpr uint32_t QCAP = 64;            // entries per warp; the flush threshold is 32
struct __align__(16) QEntry { int s, e, pe; uint32_t jk; };     // jk = j | slot << 16
__host__ __device__ inline uint32_t count_queue_bytes(uint32_t threads)
{
    return (threads >> 5) * QCAP * (uint32_t)sizeof(QEntry) + (((threads >> 5) * 4u + 15u) & ~15u);
}

template <int COUNTER>
__device__ __forceinline__ void queue_resolve(const uint8_t *__restrict__ tile, const QEntry *__restrict__ q,
                                              uint32_t first, uint32_t count, int lane, uint32_t (&acc)[KMAX])
{
    if ((uint32_t)lane < count) {
        const QEntry en = q[first + lane];
        const uint32_t kk = en.jk >> 16, j = en.jk & 0xffffu;
        const uint32_t off = reinterpret_cast<const TileHeader *>(tile)->iv_off[kk];
        const uint32_t r = scan_from<COUNTER>(reinterpret_cast<const uint2 *>(tile + off), j, en.s, en.e, en.pe);
#pragma unroll
        for (int t = 0; t < KMAX; t++) acc[t] += ((uint32_t)t == kk) ? r : 0u;
    }
}

// One sample's segments on one key against the <= KMAX tracks of a tile.  `tile` points into shared
// memory (staged tiles) or global memory; the code is instantiated once per address space.
//
// Indexed fast path per (segment, track): ONE 16-byte load fetches the bin entry of all 8 tracks,
// two independent 8-byte loads fetch the candidate interval and its successor, and the result is
// resolved branch-free; pairs that need a longer scan (more than one interval ends inside the bin
// before s, or the segment runs past the candidate's end into the next interval) are deferred to
// the warp's queue.  The next batch of segments is loaded while the current one is processed.
template <int COUNTER, bool INDEXED>
__device__ __forceinline__ void count_sample(const uint8_t *__restrict__ tile, const uint32_t (&iv_off)[KMAX],
                                             const uint32_t (&nn)[KMAX], uint32_t idx_off, uint32_t nbins,
                                             uint32_t shift, uint32_t ka,
                                             const uint64_t *__restrict__ segs, uint32_t n, int lane,
                                             QEntry *__restrict__ queue, uint32_t *__restrict__ qtail,
                                             uint32_t (&acc)[KMAX])
{
    const bool need_prev = (COUNTER == GATB_ANNOTATION_OVERLAP || COUNTER == GATB_ANNOTATION_MIDOVERLAP);
    uint64_t xnext = ((uint32_t)lane < n) ? segs[lane] : 0;
    for (uint32_t b0 = 0; b0 < n; b0 += 32) {
        const uint32_t i = b0 + lane;
        const uint64_t x = xnext;
        xnext = (i + 32 < n) ? segs[i + 32] : 0;                      // software prefetch of the next batch
        const int s = (int)seg_start(x), e = (int)seg_end(x);
        int pe = 0;
        if (need_prev && i > 0 && i < n) pe = (int)seg_end(segs[i - 1]);
        if (INDEXED) {
            uint32_t slow = 0;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (i < n) {
                const uint32_t b = min((uint32_t)s >> shift, nbins);
                q = *reinterpret_cast<const uint4 *>(tile + idx_off + (size_t)b * 16);
                const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int kk = 0; kk < KMAX; kk++) {
                    const uint32_t j = (kk & 1) ? (qw[kk >> 1] >> 16) : (qw[kk >> 1] & 0xffffu);
                    const uint2 *iv = reinterpret_cast<const uint2 *>(tile + iv_off[kk]);
                    const uint2 c0 = iv[j], c1 = iv[j + 1];
                    const bool skip = (int)c0.y <= s;
                    const int ax = (int)(skip ? c1.x : c0.x), ay = (int)(skip ? c1.y : c0.y);
                    bool more = skip && ((int)c1.y <= s);
                    // the segment runs past the candidate's end: exact only if the next interval starts
                    // at or after e (known when the candidate is c0, whose successor c1 is loaded)
                    const bool tail = (ay < e) && (skip || ((int)c1.x < e));
                    uint32_t r;
                    if (COUNTER == GATB_SEGMENT_OVERLAP) {
                        r = (ax < e) ? 1u : 0u;
                    } else if (COUNTER == GATB_SEGMENT_MIDOVERLAP) {
                        const int mid = s + ((e - s) >> 1);
                        r = (ax < e && ax <= mid && mid < ay) ? 1u : 0u;
                    } else if (COUNTER == GATB_NUCLEOTIDE_OVERLAP) {
                        r = (uint32_t)max(min(e, ay) - max(s, ax), 0);
                        more = more || tail;
                    } else if (COUNTER == GATB_ANNOTATION_OVERLAP) {
                        r = (ax < e && ax >= pe) ? 1u : 0u;
                        more = more || tail;
                    } else {
                        const int m = ax + ((ay - ax) >> 1);
                        r = (ax < e && ax >= pe && s <= m && m < e) ? 1u : 0u;
                        more = more || tail;
                    }
                    acc[kk] += more ? 0u : r;
                    slow |= more ? (1u << kk) : 0u;
                }
            }
            const uint32_t cnt = __popc(slow);
            const uint32_t total = __reduce_add_sync(GATB_FULL, cnt);
            if (total) {
                if (total <= QCAP - 32) {
                    if (slow) {                                        // push this lane's pairs
                        uint32_t pos = atomicAdd(qtail, cnt);
                        while (slow) {
                            const int kk = __ffs(slow) - 1;
                            slow &= slow - 1;
                            const uint32_t w = (kk & 4) ? ((kk & 2) ? q.w : q.z) : ((kk & 2) ? q.y : q.x);
                            QEntry en;
                            en.s = s; en.e = e; en.pe = pe;
                            en.jk = ((kk & 1) ? (w >> 16) : (w & 0xffffu)) | ((uint32_t)kk << 16);
                            queue[pos++] = en;
                        }
                    }
                    __syncwarp();
                    const uint32_t t = *qtail;
                    if (t >= 32) {
                        queue_resolve<COUNTER>(tile, queue, t - 32, 32, lane, acc);
                        __syncwarp();
                        if (lane == 0) *qtail = t - 32;
                        __syncwarp();
                    }
                } else {
                    // a batch with more than 32 unresolved pairs (dense overlaps): resolve in place
                    while (slow) {
                        const int kk = __ffs(slow) - 1;
                        slow &= slow - 1;
                        const uint32_t w = (kk & 4) ? ((kk & 2) ? q.w : q.z) : ((kk & 2) ? q.y : q.x);
                        const uint32_t j = (kk & 1) ? (w >> 16) : (w & 0xffffu);
                        const uint32_t off = reinterpret_cast<const TileHeader *>(tile)->iv_off[kk];
                        const uint32_t r = scan_from<COUNTER>(reinterpret_cast<const uint2 *>(tile + off), j, s, e, pe);
#pragma unroll
                        for (int t = 0; t < KMAX; t++) acc[t] += (t == kk) ? r : 0u;
                    }
                }
            }
        } else if (i < n) {
#pragma unroll
            for (int kk = 0; kk < KMAX; kk++) {
                if ((uint32_t)kk < ka) {
                    const uint2 *iv = reinterpret_cast<const uint2 *>(tile + iv_off[kk]);
                    uint32_t lo = 0, hi = nn[kk];         // first j with end > s (utils/gat_utils.c:8-32)
                    while (lo < hi) {
                        uint32_t mid = (lo + hi) >> 1;
                        if ((int)iv[mid].y <= s) lo = mid + 1; else hi = mid;
                    }
                    acc[kk] += scan_from<COUNTER>(iv, lo, s, e, pe);
                }
            }
        }
    }
    if (INDEXED) {                                                     // drain the queue: acc is per sample
        const uint32_t t = *qtail;
        if (t) {
            queue_resolve<COUNTER>(tile, queue, 0, t, lane, acc);
            __syncwarp();
            if (lane == 0) *qtail = 0;
            __syncwarp();
        }
    }
}

template <int COUNTER, bool DENSITY>
__global__ void __launch_bounds__(512, 1) count_kernel(CountParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    // layout: [acc: schunk*KMAX*(DENSITY?16:4) bytes][per-warp queues + tails][tile]; density keeps
    // (sum, compensation) per slot
    const uint32_t acc_bytes = (p.schunk * KMAX * (DENSITY ? 16u : 4u) + 15u) & ~15u;
    uint32_t *acc_u = reinterpret_cast<uint32_t *>(smem);
    double *acc_d = reinterpret_cast<double *>(smem);
    QEntry *queues = reinterpret_cast<QEntry *>(smem + acc_bytes);
    uint32_t *qtails = reinterpret_cast<uint32_t *>(smem + acc_bytes + (blockDim.x >> 5) * QCAP * sizeof(QEntry));
    uint8_t *tile_s = smem + acc_bytes + count_queue_bytes(blockDim.x);

    const uint32_t g = blockIdx.x;
    const uint32_t a0 = g * p.ka;
    const uint32_t ka = min(p.ka, p.n_annot - a0);
    const uint32_t s_begin = blockIdx.y * p.schunk;
    const uint32_t s_end = min(s_begin + p.schunk, p.n_samples);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    for (uint32_t i = threadIdx.x; i < p.schunk * KMAX; i += blockDim.x) {
        if (DENSITY) { acc_d[2 * i] = 0.0; acc_d[2 * i + 1] = 0.0; } else acc_u[i] = 0u;
    }
    if (lane == 0) qtails[warp] = 0;
    QEntry *queue = queues + (size_t)warp * QCAP;
    uint32_t *qtail = qtails + warp;

    for (uint32_t k = 0; k < p.n_keys; k++) {
        const uint32_t tbytes = p.tile_bytes[(uint64_t)g * p.n_keys + k];
        const uint8_t *tile_g = p.tiles + p.tile_off[(uint64_t)g * p.n_keys + k];
        const bool staged = tbytes <= p.smem_tile_budget;
        __syncthreads();                                  // previous tile fully consumed / acc init
        if (staged) {
            const uint4 *src = reinterpret_cast<const uint4 *>(tile_g);
            uint4 *dst = reinterpret_cast<uint4 *>(tile_s);
            for (uint32_t i = threadIdx.x; i < (tbytes >> 4); i += blockDim.x) dst[i] = src[i];
            __syncthreads();
        }
        if (DENSITY && p.key_ws_nseg[k] == 0) continue;   // counter returns 0 (gat/Engine.pyx:1438-1440)
        const double den = DENSITY ? (double)p.key_ws_nseg[k] : 1.0;
        // tile metadata -> registers, once per key
        const TileHeader *h = reinterpret_cast<const TileHeader *>(tile_g);
        uint32_t iv_off[KMAX], nn[KMAX];
#pragma unroll
        for (int kk = 0; kk < KMAX; kk++) { iv_off[kk] = h->iv_off[kk]; nn[kk] = h->n[kk]; }
        const uint32_t idx_off = h->idx_off, nbins = h->nbins, shift = h->shift;

        for (uint32_t sl = s_begin + warp; sl < s_end; sl += nwarps) {
            if (p.key_present && !p.key_present[(uint64_t)sl * p.n_keys + k]) continue;
            const uint32_t n = p.placed_n[(uint64_t)sl * p.n_keys + k];
            if (n == 0) continue;
            const uint64_t *segs = p.placed + (uint64_t)sl * p.sample_stride + p.key_base[k];
            uint32_t acc[KMAX];
#pragma unroll
            for (int kk = 0; kk < KMAX; kk++) acc[kk] = 0;
            if (staged) {
                if (nbins) count_sample<COUNTER, true>(tile_s, iv_off, nn, idx_off, nbins, shift, ka, segs, n, lane, queue, qtail, acc);
                else count_sample<COUNTER, false>(tile_s, iv_off, nn, idx_off, nbins, shift, ka, segs, n, lane, queue, qtail, acc);
            } else {
                if (nbins) count_sample<COUNTER, true>(tile_g, iv_off, nn, idx_off, nbins, shift, ka, segs, n, lane, queue, qtail, acc);
                else count_sample<COUNTER, false>(tile_g, iv_off, nn, idx_off, nbins, shift, ka, segs, n, lane, queue, qtail, acc);
            }
            uint32_t mine = 0;
#pragma unroll
            for (int kk = 0; kk < KMAX; kk++) {
                uint32_t tot = __reduce_add_sync(GATB_FULL, acc[kk]);
                if (lane == kk) mine = tot;
            }
            if ((uint32_t)lane < ka) {
                const uint32_t slot = (sl - s_begin) * KMAX + lane;
                if (DENSITY) {
                    // float(overlap) / len(workspace), accumulated like the reference's Python sum():
                    // CPython >= 3.12 sums floats with Neumaier compensation (Python/bltinmodule.c)
                    const double x = (double)mine / den, f = acc_d[2 * slot], t = f + x;
                    acc_d[2 * slot + 1] += (fabs(f) >= fabs(x)) ? ((f - t) + x) : ((x - t) + f);
                    acc_d[2 * slot] = t;
                } else acc_u[slot] += mine;
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (s_end - s_begin) * KMAX; i += blockDim.x) {
        const uint32_t sl = s_begin + i / KMAX, kk = i % KMAX;
        if (kk < ka) {
            if (DENSITY) {
                const double f = acc_d[2 * i], c = acc_d[2 * i + 1];
                p.out_f64[(uint64_t)sl * p.n_annot + a0 + kk] = (c != 0.0 && isfinite(c)) ? f + c : f;
            }
            else p.out_u32[(uint64_t)sl * p.n_annot + a0 + kk] = acc_u[i];
        }
    }
}

template <int COUNTER, bool DENSITY>
static cudaError_t launch_count_t(cudaStream_t st, const CountParams &p, int threads)
{
    const uint32_t acc_bytes = (p.schunk * KMAX * (DENSITY ? 16u : 4u) + 15u) & ~15u;
    const size_t smem = (size_t)acc_bytes + count_queue_bytes((uint32_t)threads) + p.smem_tile_budget;
    cudaError_t e = cudaFuncSetAttribute(count_kernel<COUNTER, DENSITY>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(p.n_groups, (p.n_samples + p.schunk - 1) / p.schunk);
    count_kernel<COUNTER, DENSITY><<<grid, threads, smem, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// tile construction on the device (replaces a host loop over every interval and bin)
__global__ void __launch_bounds__(256) build_tiles_kernel(BuildTilesParams p)
{
    const uint32_t tile = blockIdx.x;
    const uint32_t g = tile / p.n_keys, k = tile % p.n_keys;
    const TileHeader h = p.headers[tile];
    uint8_t *tp = p.tiles + p.tile_off[tile];
    if (threadIdx.x < sizeof(TileHeader) / 4)
        reinterpret_cast<uint32_t *>(tp)[threadIdx.x] = reinterpret_cast<const uint32_t *>(&p.headers[tile])[threadIdx.x];
    uint32_t err = 0;
    for (uint32_t kk = 0; kk < (uint32_t)KMAX; kk++) {
        uint64_t base = 0;
        uint32_t n = 0;
        if (kk < p.ka && g * p.ka + kk < p.n_annot) {
            const uint64_t l = (uint64_t)(g * p.ka + kk) * p.n_keys + k;
            base = p.offs[l];
            n = (uint32_t)(p.offs[l + 1] - base);
        }
        const uint32_t *ls = p.start + base, *le = p.end + base;
        uint2 *iv = reinterpret_cast<uint2 *>(tp + h.iv_off[kk]);
        for (uint32_t i = threadIdx.x; i < n + 2; i += blockDim.x) {
            uint2 v = make_uint2(0x7fffffffu, 0x7fffffffu);          // two sentinels after the list
            if (i < n) {
                v = make_uint2(ls[i], le[i]);
                if (v.y >= 0x80000000u) err |= 1u;
                if (v.x >= v.y) err |= 2u;
                if (i > 0 && le[i - 1] > v.x) err |= 2u;
            }
            iv[i] = v;
        }
        if (h.nbins) {
            uint16_t *idx = reinterpret_cast<uint16_t *>(tp + h.idx_off);
            for (uint32_t b = threadIdx.x; b <= h.nbins; b += blockDim.x) {
                uint32_t lo = 0, hi = n;                               // first j with end > (b << shift)
                if (b < h.nbins) {
                    const uint64_t pos = (uint64_t)b << h.shift;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if ((uint64_t)le[mid] <= pos) lo = mid + 1; else hi = mid;
                    }
                } else lo = n;
                idx[(size_t)b * 8 + kk] = (uint16_t)lo;
            }
        }
    }
    if (err) atomicOr(p.error, err);
}

void launch_build_tiles(cudaStream_t st, const BuildTilesParams &p)
{
    const uint32_t tiles = p.n_groups * p.n_keys;
    if (tiles) build_tiles_kernel<<<tiles, 256, 0, st>>>(p);
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
    }
    return cudaErrorInvalidValue;
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
    unsigned long long isum = 0, ntl = 0, nlt = 0, neq = 0;
    double fsum = 0.0;
    for (uint64_t s = threadIdx.x; s < p.n_samples; s += blockDim.x) {
        double x;
        if (p.is_float) { x = reinterpret_cast<const double *>(p.counts)[s * p.n_cols + col]; fsum += x; }
        else { uint32_t v = reinterpret_cast<const uint32_t *>(p.counts)[s * p.n_cols + col]; isum += v; x = (double)v; }
        ntl += (x < obs) ? 1ull : 0ull;                 // searchargsorted + cmpDouble (gat/Engine.pyx:122-127)
        nlt += (x < obs) ? 1ull : 0ull;
        neq += (x == obs) ? 1ull : 0ull;
    }
    unsigned long long r;
    r = block_reduce_sum(ntl, su); if (threadIdx.x == 0) p.n_trunc_lt[col] = r;
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

void launch_stats_pass1(cudaStream_t st, const StatsParams &p)
{
    if (p.n_cols) stats_pass1_kernel<<<p.n_cols, 256, 0, st>>>(p);
}
void launch_stats_pass2(cudaStream_t st, const StatsParams &p)
{
    if (p.n_cols) stats_pass2_kernel<<<p.n_cols, 256, 0, st>>>(p);
}
void launch_stats_select(cudaStream_t st, const StatsParams &p, uint64_t *)
{
    if (p.n_cols) stats_select_kernel<<<p.n_cols, 256, 0, st>>>(p);
}

}  // namespace gatb
