// count.cu -- K3/K4 counting kernel, tile construction, and K5 column statistics.  sm_100a.
//
// Counting: a CTA owns (group of <= KMAX annotation tracks) x (chunk of samples) and walks the keys
// (contigs) in order.  Per key it stages the group's FILTER (occupancy bitmap + bin index + union of
// the tracks' intervals, see count.cuh) in shared memory, then every warp streams its samples' segments
// on that key through it: lane = segment, one 4-byte shared-memory load of the bitmap decides whether
// the segment can overlap ANY of the group's tracks.  The ~11 % that can are pushed on a per-warp queue
// in shared memory and resolved 32 at a time, all lanes busy: bin probe + union intervals give the exact
// answer (7 % do overlap), then the per-track pass over the union interval's constituents (global
// memory / L2) adds every overlap to the (sample, track) accumulator in shared memory with integer
// atomics (order-independent, hence deterministic).  The
// float64 nucleotide-density sum is formed per key from those integers, in the same key order and with
// the same compensated summation as the reference's Python sum() (gat/__init__.py:583-587).
#include <algorithm>
#include "count.cuh"
#include "../../include/gat_b200.h"

namespace gatb {

constexpr uint32_t QCAP = 64;            // queue entries per warp; flushed whenever 32 are waiting
struct __align__(16) QEntry { int s, e; uint32_t is, pad; };  // pad is never read // segment; its index in the list (< 2^24) | sample slot << 24

// ---------------------------------------------------------------------------------------------------
// TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers for the filter staging
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

size_t count_smem_overhead(int threads, uint32_t schunk, bool density)
{
    // integer accumulators [schunk][KMAX] (+ for density: (sum, compensation) doubles per slot)
    const size_t acc = (size_t)schunk * KMAX * (density ? 20u : 4u);
    return ((acc + 15u) & ~(size_t)15u) + (size_t)(threads / 32) * QCAP * sizeof(QEntry);
}


// ---------------------------------------------------------------------------------------------------
// Exact counts of one queued segment [s,e) against every track of the tile.  Coordinates are < 2^31,
// so signed compares are exact and the sentinels are (INT_MAX, INT_MAX).  `filt` holds the bin index
// and the intervals (shared or global memory), `tile_g` is the tile in global memory.
//   nucleotide-overlap   overlapWithSegments            gat/SegmentList.pyx:1026-1076
//   segment-overlap      intersectionWithSegments(base) :1078-1146  (a segment counts once per track)
//   segment-midoverlap   midpoint tested against the FIRST overlapping interval of the track (:1137-1144)
//   annotation-*         roles swapped: an interval is counted by the first segment overlapping it,
//                        i.e. when it does not already overlap the previous segment (start >= pe)
// one overlapping interval [x,y) of track slot t against the queued segment [s,e)
template <int COUNTER>
__device__ __forceinline__ void count_pair(int s, int e, int x, int y, uint32_t t, int pe, uint32_t *__restrict__ acc,
                                           uint32_t &seen, uint32_t &hit)
{
    if (COUNTER == GATB_NUCLEOTIDE_OVERLAP) {
        atomicAdd(acc + t, (uint32_t)(min(e, y) - max(s, x)));
    } else if (COUNTER == GATB_SEGMENT_OVERLAP) {
        hit |= 1u << t;
    } else if (COUNTER == GATB_SEGMENT_MIDOVERLAP) {
        if (!((seen >> t) & 1u)) {
            seen |= 1u << t;
            const int mid = s + ((e - s) >> 1);
            if (x <= mid && mid < y) hit |= 1u << t;
        }
    } else if (COUNTER == GATB_OVERLAP_PIECES) {
        atomicAdd(acc + t, 1u);                     // every overlapping pair is one piece of intersect()
    } else if (COUNTER == GATB_ANNOTATION_OVERLAP) {
        if (x >= pe) atomicAdd(acc + t, 1u);
    } else {
        if (x >= pe) {
            const int m = x + ((y - x) >> 1);
            if (s <= m && m < e) atomicAdd(acc + t, 1u);
        }
    }
}

// `hs` is the tile header in shared memory (always staged), read here rather than carried in registers
// through the streaming loop.
template <int COUNTER>
__device__ __forceinline__ void resolve_entry(const TileHeader *__restrict__ hs, const uint8_t *__restrict__ filt,
                                              const uint8_t *__restrict__ tile_g, const QEntry en,
                                              const uint64_t *__restrict__ placed_key,
                                              uint64_t sample_stride, uint32_t s_begin, uint32_t *__restrict__ acc_s)
{
    const bool need_prev = (COUNTER == GATB_ANNOTATION_OVERLAP || COUNTER == GATB_ANNOTATION_MIDOVERLAP);
    const uint2 *civ = reinterpret_cast<const uint2 *>(filt + hs->civ_off);
    const int s = en.s, e = en.e;
    // where the walk starts: the first interval of the first union interval with end > (bin of) s
    uint32_t c;
    const uint32_t nbins = hs->nbins;
    if (nbins) {
        c = reinterpret_cast<const uint16_t *>(filt + hs->idx_off)[min(__umulhi((uint32_t)s, hs->inv), nbins)];
    } else {
        const uint2 *uiv = reinterpret_cast<const uint2 *>(tile_g + hs->uiv_off);
        uint32_t lo = 0, hi = hs->n_union;                 // lower_bound (utils/gat_utils.c:8-32)
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((int)uiv[mid].y <= s) lo = mid + 1; else hi = mid;
        }
        c = reinterpret_cast<const uint32_t *>(tile_g + hs->uoff_off)[lo];
    }
    uint2 v = civ[c];
    while ((int)v.y <= s) v = civ[++c];             // the sentinel stops the scan
    if ((int)v.x >= e) return;                      // bitmap false positive: the segment falls in a gap
    const uint8_t *cslot = filt + hs->cslot_off;
    const uint32_t slot = en.is >> 24, i = en.is & 0xffffffu;
    uint32_t *acc = acc_s + slot * KMAX;            // shared accumulators of this sample: integer atomics,
    int pe = 0;                                     // so the result does not depend on the order of arrival
    if (need_prev && i > 0)
        pe = (int)seg_end(placed_key[(uint64_t)(s_begin + slot) * sample_stride + i - 1]);
    uint32_t seen = 0, hit = 0;
    do {                                            // intervals are sorted by start: stop at start >= e
        if ((int)v.y > s) count_pair<COUNTER>(s, e, (int)v.x, (int)v.y, cslot[c], pe, acc, seen, hit);
        v = civ[++c];
    } while ((int)v.x < e);
    if (COUNTER == GATB_SEGMENT_OVERLAP || COUNTER == GATB_SEGMENT_MIDOVERLAP) {
        while (hit) {
            const int t = __ffs(hit) - 1;
            hit &= hit - 1;
            atomicAdd(acc + t, 1u);
        }
    }
}

// All samples of one warp on one key against the tile.  `bm` is the staged bitmap (shared memory);
// `filt` points at the staged tile in shared memory or, for filters too large to stage, at the tile in
// global memory (the code is instantiated once per address space).  A lane loads two segments (16 bytes)
// per access, 64 per warp, and the two halves of the unrolled loop keep 128 segments in flight ahead of
// the one being tested (they swap roles, so no register rotation); the queue is carried across the
// warp's samples and drained once per key.
//
// Bitmap test: bit b is set when a union interval touches [b << sh, (b + 2) << sh), so ONE bit decides
// for every segment no longer than 1 << sh (it lies inside that window); longer segments are candidates
// outright.  Bits >= bm_bits are zero (padding word), which also retires lanes past the end of the list:
// they carry the all-ones pattern, whose start lies beyond every coordinate and whose length is 0.
template <int COUNTER>
__device__ __forceinline__ void count_key(const TileHeader *__restrict__ hs, const uint32_t *__restrict__ bm,
                                          const uint8_t *__restrict__ filt, const uint8_t *__restrict__ tile_g,
                                          const CountParams &p, uint32_t k,
                                          uint32_t s_begin, uint32_t s_end, int lane, int warp, int nwarps,
                                          QEntry *__restrict__ queue, uint32_t *__restrict__ acc_s)
{
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t sh = hs->bm_shift, nbits = hs->bm_bits, wlen = 1u << sh;
    const uint64_t *placed_key = p.placed + p.key_base[k];
    const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
    uint32_t qn = 0;                                                   // warp-uniform queue fill
    for (uint32_t sl = s_begin + warp; sl < s_end; sl += nwarps) {
        if (p.key_present && !p.key_present[(uint64_t)sl * p.n_keys + k]) continue;
        const uint32_t n = p.placed_n[(uint64_t)sl * p.n_keys + k];
        if (n == 0) continue;
        // (key bases and the sample stride are even, so pairs are 16-byte aligned; reading the unused
        // slot after an odd n stays inside the key's buffer)
        const uint4 *segs = reinterpret_cast<const uint4 *>(placed_key + (uint64_t)sl * p.sample_stride) + lane;
        const uint32_t slot24 = (sl - s_begin) << 24;
        const uint32_t i0 = 2u * (uint32_t)lane;
        const uint32_t nl = n > i0 ? n - i0 : 0u;                      // this lane's pair of block b0 exists if b0 < nl
        uint4 xa = (0u < nl) ? segs[0] : ones;                         // block b0      (64 segments per warp)
        uint4 xb = (64u < nl) ? segs[32] : ones;                       // block b0 + 64
        // one segment: test, compact the candidates onto the queue, resolve when 32 are waiting
#define GATB_COUNT_ONE(S, E, I)                                                                            \
        {                                                                                                  \
            const uint32_t s_ = (S), e_ = (E);                                                             \
            const uint32_t bs_ = min(s_ >> sh, nbits);                                                     \
            const bool flag_ = (((bm[bs_ >> 5] >> (bs_ & 31u)) & 1u) != 0u) || (e_ - s_ > wlen);           \
            const uint32_t m_ = __ballot_sync(GATB_FULL, flag_);                                           \
            if (flag_) {                                                                                   \
                QEntry en_;                                                                                \
                en_.s = (int)s_; en_.e = (int)e_; en_.is = (I) | slot24;                                   \
                queue[qn + __popc(m_ & lt_mask)] = en_;                                                    \
            }                                                                                              \
            qn += __popc(m_);                                                                              \
            if (qn >= 32) {                                                                                \
                __syncwarp();                                                                              \
                qn -= 32;                                                                                  \
                resolve_entry<COUNTER>(hs, filt, tile_g, queue[qn + lane], placed_key, p.sample_stride,    \
                                       s_begin, acc_s);                                                    \
                __syncwarp();                                                                              \
            }                                                                                              \
        }
#define GATB_COUNT_BLOCK(X, B0)                                                                            \
        {                                                                                                  \
            uint4 v_ = X;                                                                                  \
            if ((B0) + 1u >= nl) { v_.z = ~0u; v_.w = ~0u; }          /* odd n: the pair's second slot */  \
            X = ((B0) + 128u < nl) ? segs[((B0) + 128u) >> 1] : ones; /* prefetch, two blocks ahead */     \
            GATB_COUNT_ONE(v_.y, v_.x, (B0) + i0)                                                          \
            GATB_COUNT_ONE(v_.w, v_.z, (B0) + i0 + 1u)                                                     \
        }
        for (uint32_t b0 = 0; b0 < n; b0 += 128) {
            GATB_COUNT_BLOCK(xa, b0)
            if (b0 + 64 < n) GATB_COUNT_BLOCK(xb, b0 + 64u)
        }
#undef GATB_COUNT_BLOCK
#undef GATB_COUNT_ONE
    }
    if (qn) {                                                          // drain once per key
        __syncwarp();
        if ((uint32_t)lane < qn)
            resolve_entry<COUNTER>(hs, filt, tile_g, queue[lane], placed_key, p.sample_stride, s_begin, acc_s);
        __syncwarp();
    }
}

template <int COUNTER, bool DENSITY>
__global__ void __launch_bounds__(1024, 1) count_kernel(CountParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    // layout: [acc_u: schunk*KMAX u32][density only: acc_d: schunk*KMAX (sum, compensation) doubles]
    //         [per-warp queues][filter]
    const uint32_t nslots = p.schunk * KMAX;
    const uint32_t acc_bytes = (nslots * (DENSITY ? 20u : 4u) + 15u) & ~15u;
    uint32_t *acc_u = reinterpret_cast<uint32_t *>(smem + (DENSITY ? nslots * 16u : 0u));
    double *acc_d = reinterpret_cast<double *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    QEntry *queue = reinterpret_cast<QEntry *>(smem + acc_bytes) + (size_t)warp * QCAP;
    uint8_t *filt_s = smem + acc_bytes + (size_t)nwarps * QCAP * sizeof(QEntry);

    const uint32_t g = blockIdx.x;
    const uint32_t a0 = g * p.ka;
    const uint32_t ka = min(p.ka, p.n_annot - a0);
    const uint32_t s_begin = blockIdx.y * p.schunk;
    const uint32_t s_end = min(s_begin + p.schunk, p.n_samples);

    for (uint32_t i = threadIdx.x; i < nslots; i += blockDim.x) {
        acc_u[i] = 0u;
        if (DENSITY) { acc_d[2 * i] = 0.0; acc_d[2 * i + 1] = 0.0; }
    }
    __shared__ __align__(8) uint64_t stage_bar;           // mbarrier of the filter staging (one arrival + tx bytes)
    uint32_t stage_phase = 0;
    if (threadIdx.x == 0) mbar_init(&stage_bar, 1);

    for (uint32_t k = 0; k < p.n_keys; k++) {
        const uint8_t *tile_g = p.tiles + p.tile_off[(uint64_t)g * p.n_keys + k];
        const TileHeader h = *reinterpret_cast<const TileHeader *>(tile_g);
        const bool staged = p.tile_stage[(uint64_t)g * p.n_keys + k] <= p.smem_tile_budget;
        __syncthreads();                                  // previous filter fully consumed / acc init
        if (h.n_union == 0) continue;                     // no interval of any track on this key
        if (DENSITY && p.key_ws_nseg[k] == 0) continue;   // counter returns 0 (gat/Engine.pyx:1438-1440)
        {
            // ONE bulk copy by the TMA engine (cp.async.bulk, 1-D), completion signalled on an mbarrier:
            // header + bitmap + bin index + the intervals, or header + bitmap alone when the whole filter
            // does not fit; meanwhile the next key's filter is prefetched into L2
            const uint32_t bytes = staged ? h.stage_bytes : h.idx_off;
            if (threadIdx.x == 0) {
                fence_proxy_async();                          // order earlier generic reads of filt_s before the async write
                mbar_expect_tx(&stage_bar, bytes);
                bulk_copy_g2s(filt_s, tile_g, bytes, &stage_bar);
                if (k + 1 < p.n_keys) {
                    const uint32_t nb = p.tile_stage[(uint64_t)g * p.n_keys + k + 1];
                    if (nb <= p.smem_tile_budget)
                        bulk_prefetch_l2(p.tiles + p.tile_off[(uint64_t)g * p.n_keys + k + 1], nb);
                }
            }
            mbar_wait(&stage_bar, stage_phase);
            stage_phase ^= 1u;
        }
        const TileHeader *hs = reinterpret_cast<const TileHeader *>(filt_s);
        const uint32_t *bm = reinterpret_cast<const uint32_t *>(filt_s + h.bm_off);
        if (staged) count_key<COUNTER>(hs, bm, filt_s, tile_g, p, k, s_begin, s_end, lane, warp, nwarps, queue, acc_u);
        else count_key<COUNTER>(hs, bm, tile_g, tile_g, p, k, s_begin, s_end, lane, warp, nwarps, queue, acc_u);
        if (DENSITY) {
            // per key: float(overlap) / len(workspace), accumulated in key order like the reference's
            // Python sum(): CPython >= 3.12 sums floats with Neumaier compensation (Python/bltinmodule.c)
            __syncthreads();
            const double den = (double)p.key_ws_nseg[k];
            for (uint32_t i = threadIdx.x; i < nslots; i += blockDim.x) {
                const uint32_t v = acc_u[i];
                if (v) {
                    const double x = (double)v / den, f = acc_d[2 * i], t = f + x;
                    acc_d[2 * i + 1] += (fabs(f) >= fabs(x)) ? ((f - t) + x) : ((x - t) + f);
                    acc_d[2 * i] = t;
                    acc_u[i] = 0u;
                }
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
            } else p.out_u32[(uint64_t)sl * p.n_annot + a0 + kk] = acc_u[i];
        }
    }
}

template <int COUNTER, bool DENSITY>
static cudaError_t launch_count_t(cudaStream_t st, const CountParams &p, int threads)
{
    const size_t smem = count_smem_overhead(threads, p.schunk, DENSITY) + p.smem_tile_budget;
    cudaError_t e = cudaFuncSetAttribute(count_kernel<COUNTER, DENSITY>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(p.n_groups, (p.n_samples + p.schunk - 1) / p.schunk);
    count_kernel<COUNTER, DENSITY><<<grid, threads, smem, st>>>(p);
    return cudaGetLastError();
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
// Tile construction on the device, one CTA per tile (replaces a host loop over every interval):
//   A  merge the <= KMAX sorted track lists by start into civ[] / cslot[] (rank = own index + lower/upper
//      bounds in the other lists: no sort, deterministic ties by slot) and validate the lists
//   B  union: running max of ends over civ[] -> heads, union intervals, CSR offsets
//   C  sentinels + header
//   D  bin index over the union ends
//   E  occupancy bitmap of the union
constexpr int BT = 256;

__global__ void __launch_bounds__(BT) build_tiles_kernel(BuildTilesParams p)
{
    __shared__ int s_max[BT];
    __shared__ uint32_t s_cnt[BT];
    __shared__ uint64_t s_base[KMAX];
    __shared__ uint32_t s_n[KMAX];
    __shared__ uint32_t s_nu;
    const uint32_t tile = blockIdx.x, tid = threadIdx.x;
    const uint32_t g = tile / p.n_keys, k = tile % p.n_keys;
    const TileHeader h = p.headers[tile];
    uint8_t *tp = p.tiles + p.tile_off[tile];
    uint16_t *idx = reinterpret_cast<uint16_t *>(tp + h.idx_off);
    uint8_t *cslot = tp + h.cslot_off;
    uint2 *civ = reinterpret_cast<uint2 *>(tp + h.civ_off);
    uint2 *uiv = reinterpret_cast<uint2 *>(tp + h.uiv_off);
    uint32_t *uoff = reinterpret_cast<uint32_t *>(tp + h.uoff_off);
    if (tid < (uint32_t)KMAX) {
        uint64_t base = 0;
        uint32_t n = 0;
        if (tid < p.ka && g * p.ka + tid < p.n_annot) {
            const uint64_t l = (uint64_t)(g * p.ka + tid) * p.n_keys + k;
            base = p.offs[l];
            n = (uint32_t)(p.offs[l + 1] - base);
        }
        s_base[tid] = base;
        s_n[tid] = n;
    }
    __syncthreads();

    // ---- A: rank merge + validation.  Warp = list, lane = a contiguous chunk of it: the ranks of
    // consecutive elements in the other lists only move forward, so after one binary search per list
    // for the chunk's first element the lane just gallops
    uint32_t err = 0;
    for (uint32_t kk = tid >> 5; kk < (uint32_t)KMAX; kk += BT / 32) {
        const uint32_t n = s_n[kk];
        const uint32_t *ls = p.start + s_base[kk], *le = p.end + s_base[kk];
        const uint32_t chunk = (n + 31u) / 32u;
        const uint32_t lo = min((tid & 31u) * chunk, n), hi = min(lo + chunk, n);
        if (lo >= hi) continue;
        uint32_t pos[KMAX];
        const uint32_t x0 = ls[lo];
#pragma unroll
        for (uint32_t t = 0; t < (uint32_t)KMAX; t++) {
            pos[t] = 0;
            const uint32_t m = s_n[t];
            if (t == kk || m == 0) continue;
            const uint32_t *os = p.start + s_base[t];
            uint32_t l2 = 0, h2 = m;                 // elements of list t placed before (x0, kk)
            while (l2 < h2) {
                const uint32_t mid = (l2 + h2) >> 1;
                const uint32_t v = os[mid];
                if (v < x0 || (v == x0 && t < kk)) l2 = mid + 1; else h2 = mid;
            }
            pos[t] = l2;
        }
        uint32_t py = lo > 0 ? le[lo - 1] : 0u;
        for (uint32_t i = lo; i < hi; i++) {
            const uint32_t x = ls[i], y = le[i];
            if (y >= 0x80000000u) err |= 1u;
            if (x >= y) err |= 2u;
            if (i > 0 && py > x) err |= 2u;
            py = y;
            uint32_t rank = i;
#pragma unroll
            for (uint32_t t = 0; t < (uint32_t)KMAX; t++) {
                const uint32_t m = s_n[t];
                if (t == kk || m == 0) continue;
                const uint32_t *os = p.start + s_base[t];
                uint32_t q = pos[t];
                while (q < m) {
                    const uint32_t v = os[q];
                    if (v < x || (v == x && t < kk)) q++; else break;
                }
                pos[t] = q;
                rank += q;
            }
            civ[rank] = make_uint2(x, y);
            cslot[rank] = (uint8_t)kk;
        }
    }
    if (err) atomicOr(p.error, err);
    __syncthreads();

    // ---- B: union of the merged list
    const uint32_t nc = h.n_cons;
    const uint32_t chunk = (nc + BT - 1) / BT;
    const uint32_t lo = min(tid * chunk, nc), hi = min(lo + chunk, nc);
    int mx = -1;
    for (uint32_t i = lo; i < hi; i++) mx = max(mx, (int)civ[i].y);
    s_max[tid] = mx;
    __syncthreads();
    if (tid == 0) {                                  // exclusive prefix max over the chunks
        int run = -1;
        for (int t = 0; t < BT; t++) { const int v = s_max[t]; s_max[t] = run; run = max(run, v); }
    }
    __syncthreads();
    const int carry = s_max[tid];
    {
        int run = carry;
        uint32_t heads = 0;
        for (uint32_t i = lo; i < hi; i++) {
            const uint2 v = civ[i];
            if ((int)v.x > run) heads++;             // starts a new union interval (touching ones merge)
            run = max(run, (int)v.y);
        }
        s_cnt[tid] = heads;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int t = 0; t < BT; t++) { const uint32_t v = s_cnt[t]; s_cnt[t] = run; run += v; }
        s_nu = run;
    }
    __syncthreads();
    const uint32_t nu = s_nu;
    {
        int run = carry;
        uint32_t u = s_cnt[tid];
        for (uint32_t i = lo; i < hi; i++) {
            const uint2 v = civ[i];
            if ((int)v.x > run) {
                uiv[u].x = v.x;
                uoff[u] = i;
                if (u > 0) uiv[u - 1].y = (uint32_t)run;     // the previous union interval ends at the running max
                u++;
            }
            run = max(run, (int)v.y);
        }
        if (hi == nc && lo < hi) uiv[nu - 1].y = (uint32_t)run;
    }
    // ---- C: sentinels + header
    if (tid == 0) {
        uiv[nu] = make_uint2(0x7fffffffu, 0x7fffffffu);
        uiv[nu + 1] = make_uint2(0x7fffffffu, 0x7fffffffu);
        uoff[nu] = nc;
        civ[nc] = make_uint2(0x7fffffffu, 0x7fffffffu);
        civ[nc + 1] = make_uint2(0x7fffffffu, 0x7fffffffu);
        cslot[nc] = 0;
        cslot[nc + 1] = 0;
        TileHeader out = h;
        out.n_union = nu;
        *reinterpret_cast<TileHeader *>(tp) = out;
    }
    __syncthreads();

    // ---- E: occupancy bitmap of the union (bit b: some union interval touches [b << shift, (b+2) << shift))
    {
        uint32_t *bm = reinterpret_cast<uint32_t *>(tp + h.bm_off);
        const uint32_t words = (h.idx_off - h.bm_off) >> 2;
        for (uint32_t w = tid; w < words; w += BT) bm[w] = 0u;
        __syncthreads();
        for (uint32_t u = tid; u < nu; u += BT) {
            const uint2 v = uiv[u];
            uint32_t b0 = v.x >> h.bm_shift;
            const uint32_t b1 = min((v.y - 1u) >> h.bm_shift, h.bm_bits - 1u);
            if (b0 >= h.bm_bits) continue;           // only with invalid lists, which are rejected anyway
            if (b0 > 0) b0--;                        // the window of bit b reaches one bin to the right
            for (uint32_t w = b0 >> 5; w <= (b1 >> 5) && b0 <= b1; w++) {
                const uint32_t lo = (w == (b0 >> 5)) ? (b0 & 31u) : 0u, hi = (w == (b1 >> 5)) ? (b1 & 31u) : 31u;
                atomicOr(&bm[w], (0xffffffffu >> (31u - hi)) & (0xffffffffu << lo));
            }
        }
    }

    // ---- D: bin index: idx[b] = first interval of the first union interval with end > lowest position of
    // bin b.  Scattered from the union side: union u owns the bins whose lowest position lies in
    // [end(u-1), end(u)); the first bin whose lowest position is >= y is bin(y - 1) + 1.
    if (h.nbins) {
        for (uint32_t u = tid; u <= nu; u += BT) {
            const uint32_t yp = u > 0 ? uiv[u - 1].y : 0u;
            const uint32_t b_lo = yp ? min(__umulhi(yp - 1u, h.inv) + 1u, h.nbins) : 0u;
            uint32_t b_hi, val;
            if (u < nu) {
                const uint32_t y = uiv[u].y;
                b_hi = y ? min(__umulhi(y - 1u, h.inv) + 1u, h.nbins) : 0u;
                val = uoff[u];
            } else {
                b_hi = h.nbins + 1u;                 // the rest, including entry nbins: the sentinel
                val = nc;
            }
            for (uint32_t b = b_lo; b < b_hi; b++) idx[b] = (uint16_t)val;
        }
    }
}

void launch_build_tiles(cudaStream_t st, const BuildTilesParams &p)
{
    const uint32_t tiles = p.n_groups * p.n_keys;
    if (tiles) build_tiles_kernel<<<tiles, BT, 0, st>>>(p);
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
