// count.cu -- K3/K4 counting kernel, tile construction, and K5 column statistics.  sm_100a.
//
// Counting: a CTA owns (group of <= KMAX annotation tracks) x (chunk of samples) and walks the keys
// (contigs) in order.  Per key it stages the group's FILTER (bin index + union of the tracks'
// intervals, see count.cuh) in shared memory, then every warp streams its samples' segments on that
// key through it: lane = segment, one bin probe + two 8-byte shared-memory loads decide whether the
// segment can overlap ANY of the group's tracks.  The ~7 % that can are pushed on a per-warp queue in
// shared memory and resolved 32 at a time, all lanes busy, by the exact per-track pass over the
// union interval's constituents (global memory / L2), which adds every overlap to the (sample, track)
// accumulator in shared memory with integer atomics (order-independent, hence deterministic).  The
// float64 nucleotide-density sum is formed per key from those integers, in the same key order and with
// the same compensated summation as the reference's Python sum() (gat/__init__.py:583-587).
#include "count.cuh"
#include "../../include/gat_b200.h"

namespace gatb {

constexpr uint32_t QCAP = 64;            // queue entries per warp; flushed whenever 32 are waiting
struct __align__(16) QEntry { int s, e; uint32_t is, j; }; // segment; its index in the list (< 2^24) | sample slot << 24; union start index

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
// so signed compares are exact and the sentinels are (INT_MAX, INT_MAX).  `filt` holds the union
// intervals (shared or global memory), `tile_g` the constituents (global).
//   nucleotide-overlap   overlapWithSegments            gat/SegmentList.pyx:1026-1076
//   segment-overlap      intersectionWithSegments(base) :1078-1146  (a segment counts once per track)
//   segment-midoverlap   midpoint tested against the FIRST overlapping interval of the track (:1137-1144)
//   annotation-*         roles swapped: an interval is counted by the first segment overlapping it,
//                        i.e. when it does not already overlap the previous segment (start >= pe)
template <int COUNTER>
__device__ __forceinline__ void resolve_entry(const uint8_t *__restrict__ filt, const uint8_t *__restrict__ tile_g,
                                              uint32_t uiv_off, uint32_t uoff_off, uint32_t cons_off,
                                              const QEntry en, const uint64_t *__restrict__ placed_key,
                                              uint64_t sample_stride, uint32_t s_begin, uint32_t *__restrict__ acc_s)
{
    const bool need_prev = (COUNTER == GATB_ANNOTATION_OVERLAP || COUNTER == GATB_ANNOTATION_MIDOVERLAP);
    const uint2 *uiv = reinterpret_cast<const uint2 *>(filt + uiv_off);
    const uint32_t *uoff = reinterpret_cast<const uint32_t *>(tile_g + uoff_off);
    const uint4 *cons = reinterpret_cast<const uint4 *>(tile_g + cons_off);
    const int s = en.s, e = en.e;
    const uint32_t slot = en.is >> 24, i = en.is & 0xffffffu;
    uint32_t *acc = acc_s + slot * KMAX;            // shared accumulators of this sample: integer atomics,
    int pe = 0;                                     // so the result does not depend on the order of arrival
    if (need_prev && i > 0)
        pe = (int)seg_end(placed_key[(uint64_t)(s_begin + slot) * sample_stride + i - 1]);
    uint32_t u = en.j;
    uint2 a = uiv[u];
    while ((int)a.y <= s) a = uiv[++u];         // first union interval with end > s; the sentinel stops the scan
    uint32_t seen = 0, hit = 0;
    while ((int)a.x < e) {
        uint32_t c = uoff[u];
        const uint32_t cend = uoff[u + 1];
        for (; c < cend; c++) {
            const uint4 v = cons[c];
            if ((int)v.x >= e) break;           // constituents are sorted by start
            if ((int)v.y <= s) continue;
            const uint32_t t = v.z;
            if (COUNTER == GATB_NUCLEOTIDE_OVERLAP) {
                atomicAdd(acc + t, (uint32_t)(min(e, (int)v.y) - max(s, (int)v.x)));
            } else if (COUNTER == GATB_SEGMENT_OVERLAP) {
                hit |= 1u << t;
            } else if (COUNTER == GATB_SEGMENT_MIDOVERLAP) {
                if (!((seen >> t) & 1u)) {
                    seen |= 1u << t;
                    const int mid = s + ((e - s) >> 1);
                    if ((int)v.x <= mid && mid < (int)v.y) hit |= 1u << t;
                }
            } else if (COUNTER == GATB_ANNOTATION_OVERLAP) {
                if ((int)v.x >= pe) atomicAdd(acc + t, 1u);
            } else {
                if ((int)v.x >= pe) {
                    const int m = (int)v.x + (((int)v.y - (int)v.x) >> 1);
                    if (s <= m && m < e) atomicAdd(acc + t, 1u);
                }
            }
        }
        a = uiv[++u];
    }
    if (COUNTER == GATB_SEGMENT_OVERLAP || COUNTER == GATB_SEGMENT_MIDOVERLAP) {
        while (hit) {
            const int t = __ffs(hit) - 1;
            hit &= hit - 1;
            atomicAdd(acc + t, 1u);
        }
    }
}

// All samples of one warp on one key against the tile.  `filt` points into shared memory (staged
// filters) or global memory; the code is instantiated once per address space.  Segments are loaded two
// batches ahead; the queue is carried across the warp's samples and drained once per key.
template <int COUNTER, bool INDEXED>
__device__ __forceinline__ void count_key(const uint8_t *__restrict__ filt, const uint8_t *__restrict__ tile_g,
                                          const TileHeader &h, const CountParams &p, uint32_t k,
                                          uint32_t s_begin, uint32_t s_end, int lane, int warp, int nwarps,
                                          QEntry *__restrict__ queue, uint32_t *__restrict__ acc_s)
{
    const uint2 *uiv = reinterpret_cast<const uint2 *>(filt + h.uiv_off);
    const uint16_t *idx = reinterpret_cast<const uint16_t *>(filt + h.idx_off);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint64_t *placed_key = p.placed + p.key_base[k];
    uint32_t qn = 0;                                                   // warp-uniform queue fill
    for (uint32_t sl = s_begin + warp; sl < s_end; sl += nwarps) {
        if (p.key_present && !p.key_present[(uint64_t)sl * p.n_keys + k]) continue;
        const uint32_t n = p.placed_n[(uint64_t)sl * p.n_keys + k];
        if (n == 0) continue;
        const uint64_t *segs = placed_key + (uint64_t)sl * p.sample_stride;
        const uint32_t slot24 = (sl - s_begin) << 24;
        uint64_t x1 = ((uint32_t)lane < n) ? segs[lane] : 0;           // batch b0
        uint64_t x2 = ((uint32_t)lane + 32 < n) ? segs[lane + 32] : 0; // batch b0 + 32
        for (uint32_t b0 = 0; b0 < n; b0 += 32) {
            const uint32_t i = b0 + lane;
            const uint64_t x = x1;
            x1 = x2;
            x2 = (i + 64 < n) ? segs[i + 64] : 0;                     // software prefetch, two batches ahead
            const int s = (int)seg_start(x), e = (int)seg_end(x);
            bool flag = false;
            uint32_t j = 0;
            if (i < n) {
                if (INDEXED) {
                    j = idx[min(__umulhi((uint32_t)s, h.inv), h.nbins)];
                } else {
                    uint32_t lo = 0, hi = h.n_union;                   // lower_bound (utils/gat_utils.c:8-32)
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if ((int)uiv[mid].y <= s) lo = mid + 1; else hi = mid;
                    }
                    j = lo;
                }
                const uint2 c0 = uiv[j];
                int ax = (int)c0.x;
                bool more = false;
                if ((int)c0.y <= s) {               // (~1 lane in 6) the candidate ends before s: take the next one
                    const uint2 c1 = uiv[j + 1];
                    ax = (int)c1.x;
                    more = (int)c1.y <= s;          // still not past s: let the exact pass scan
                }
                flag = more || (ax < e);
            }
            const uint32_t m = __ballot_sync(GATB_FULL, flag);
            if (m) {
                if (flag) {
                    QEntry en;
                    en.s = s; en.e = e; en.is = i | slot24; en.j = j;
                    queue[qn + __popc(m & lt_mask)] = en;
                }
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    qn -= 32;
                    resolve_entry<COUNTER>(filt, tile_g, h.uiv_off, h.uoff_off, h.cons_off, queue[qn + lane],
                                           placed_key, p.sample_stride, s_begin, acc_s);
                    __syncwarp();
                }
            }
        }
    }
    if (qn) {                                                          // drain once per key
        if ((uint32_t)lane < qn)
            resolve_entry<COUNTER>(filt, tile_g, h.uiv_off, h.uoff_off, h.cons_off, queue[lane],
                                   placed_key, p.sample_stride, s_begin, acc_s);
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
        if (staged) {
            // header + bin index + the union intervals actually present (+ 2 sentinels): ONE bulk copy
            // by the TMA engine (cp.async.bulk, 1-D), completion signalled on an mbarrier; meanwhile the
            // next key's filter is prefetched into L2
            const uint32_t bytes = (h.uiv_off + (h.n_union + 2) * 8 + 15u) & ~15u;
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
        if (staged) {
            if (h.nbins) count_key<COUNTER, true>(filt_s, tile_g, h, p, k, s_begin, s_end, lane, warp, nwarps, queue, acc_u);
            else count_key<COUNTER, false>(filt_s, tile_g, h, p, k, s_begin, s_end, lane, warp, nwarps, queue, acc_u);
        } else {
            if (h.nbins) count_key<COUNTER, true>(tile_g, tile_g, h, p, k, s_begin, s_end, lane, warp, nwarps, queue, acc_u);
            else count_key<COUNTER, false>(tile_g, tile_g, h, p, k, s_begin, s_end, lane, warp, nwarps, queue, acc_u);
        }
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
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------------
// Tile construction on the device, one CTA per tile (replaces a host loop over every interval):
//   A  merge the <= KMAX sorted track lists by start into cons[] (rank = own index + lower/upper bounds
//      in the other lists: no sort, deterministic ties by slot) and validate the lists
//   B  union: running max of ends over cons[] -> heads, union intervals, CSR offsets
//   C  sentinels + header
//   D  bin index over the union ends
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
    uint2 *uiv = reinterpret_cast<uint2 *>(tp + h.uiv_off);
    uint32_t *uoff = reinterpret_cast<uint32_t *>(tp + h.uoff_off);
    uint4 *cons = reinterpret_cast<uint4 *>(tp + h.cons_off);
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

    // ---- A: rank merge + validation
    uint32_t err = 0;
    for (uint32_t kk = 0; kk < (uint32_t)KMAX; kk++) {
        const uint32_t n = s_n[kk];
        const uint32_t *ls = p.start + s_base[kk], *le = p.end + s_base[kk];
        for (uint32_t i = tid; i < n; i += BT) {
            const uint32_t x = ls[i], y = le[i];
            if (y >= 0x80000000u) err |= 1u;
            if (x >= y) err |= 2u;
            if (i > 0 && le[i - 1] > x) err |= 2u;
            uint32_t rank = i;
            for (uint32_t t = 0; t < (uint32_t)KMAX; t++) {
                const uint32_t m = s_n[t];
                if (t == kk || m == 0) continue;
                const uint32_t *os = p.start + s_base[t];
                uint32_t lo = 0, hi = m;             // elements of list t placed before (x, kk)
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    const uint32_t v = os[mid];
                    if (v < x || (v == x && t < kk)) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            }
            cons[rank] = make_uint4(x, y, kk, 0u);
        }
    }
    if (err) atomicOr(p.error, err);
    __syncthreads();

    // ---- B: union of the merged list
    const uint32_t nc = h.n_cons;
    const uint32_t chunk = (nc + BT - 1) / BT;
    const uint32_t lo = min(tid * chunk, nc), hi = min(lo + chunk, nc);
    int mx = -1;
    for (uint32_t i = lo; i < hi; i++) mx = max(mx, (int)cons[i].y);
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
            const uint4 v = cons[i];
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
            const uint4 v = cons[i];
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
        TileHeader out = h;
        out.n_union = nu;
        *reinterpret_cast<TileHeader *>(tp) = out;
    }
    __syncthreads();

    // ---- D: bin index: idx[b] = first union interval with end > lowest position of bin b
    if (h.nbins) {
        for (uint32_t b = tid; b <= h.nbins; b += BT) {
            uint32_t l2 = 0, h2 = nu;
            if (b < h.nbins) {
                const uint64_t pos = (((uint64_t)b << 32) + h.inv - 1) / h.inv;
                while (l2 < h2) {
                    const uint32_t mid = (l2 + h2) >> 1;
                    if ((uint64_t)uiv[mid].y <= pos) l2 = mid + 1; else h2 = mid;
                }
            } else l2 = nu;
            idx[b] = (uint16_t)l2;
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
