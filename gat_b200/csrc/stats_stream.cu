// stats_stream.cu -- K5 column statistics of a uint32 count matrix [sample][column] as HBM-bound streaming passes.  sm_100a.
//
// (gat/Engine.pyx:1635-1718 AnnotatorResult statistics, :1543-1576 getTwoSidedPValue.)  The matrix is S x A uint32,
// 4 GB at the north-star size, and every statistic is a per-column reduction over all rows: the passes are pure
// streaming and the roofline is the HBM copy bandwidth.  The column-tiled kernels of count.cu put only A / 16 CTAs
// on the machine (63 for 1000 columns, 8 for a column block of 125) and ran at a few percent of it.  Here a CTA owns
// whole ROWS: a chunk of `rows_per_stage` rows is one contiguous piece of memory, which the TMA unit copies into a
// ring of shared-memory stages (cp.async.bulk, completion on an mbarrier) while the threads -- thread = column,
// row lanes when there are fewer columns than threads -- work the previous stage off with conflict-free shared
// loads.  The grid is 2 CTAs per SM (1 for the select pass), chunks are dealt round-robin.
//   pass 1   sum x (uint64), sum x^2 (128 bits), #{x < observed}, #{x == observed}, max -- all integers, accumulated
//            with atomics: exact, so the results do not depend on the grid, the GPU or the column layout.  The
//            variance follows on the host from the exact n * sum x^2 - (sum x)^2 (no second pass over the matrix)
//   select   radix select of the two order statistics (CI bounds), 4 bits per pass from the highest non-zero
//            nibble of the matrix maximum: per-CTA histograms [rank][nibble][column] in shared memory, flushed to
//            global counters; stats_pick_kernel narrows prefix and rank between passes
#include "count.cuh"
#include "ptx.cuh"

namespace gatb {

constexpr uint32_t SS_STAGE_BYTES = 32768;          // one stage of the ring
constexpr uint32_t SS_WAIT_SPINS = 1u << 20;        // mbarrier polls before a CTA gives up (flags p.error: the call fails)

__host__ __device__ __forceinline__ uint32_t ss_hist_stride(uint32_t n_cols) { return ((n_cols + 31u) & ~31u) + 1u; }   // = 1 mod 32

// shared memory of a pass: [stages][mbarriers][select: histograms]
size_t stats_stream_smem(const StreamStatsParams &p, int mode, int threads)
{
    size_t b = (size_t)p.n_stages * SS_STAGE_BYTES + 64;
    if (mode == 1) b += (size_t)2 * 16 * ss_hist_stride(p.n_cols) * sizeof(uint32_t);
    (void)threads;
    return b;
}

// the CTA's j-th chunk is chunk blockIdx.x + j * gridDim.x of the matrix
struct ChunkSeq {
    uint32_t n_local;           // chunks of this CTA
    __device__ __forceinline__ uint64_t row0(const StreamStatsParams &p, uint32_t q) const
    {
        return ((uint64_t)blockIdx.x + (uint64_t)q * gridDim.x) * p.rows_per_stage;
    }
    __device__ __forceinline__ uint32_t rows(const StreamStatsParams &p, uint32_t q) const
    {
        return (uint32_t)min((uint64_t)p.rows_per_stage, p.n_samples - row0(p, q));
    }
};

// MODE 0: pass 1, 1: select pass.  KC = columns per thread (thread (c, r) owns columns c, c + CW, ...)
template <int MODE, int KC>
__global__ void __launch_bounds__(512) stats_stream_kernel(StreamStatsParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t T = blockDim.x, tid = threadIdx.x;
    const uint32_t CW = p.col_width, RW = T / CW;           // threads across columns, row lanes
    const uint32_t c = tid % CW, r = tid / CW;
    const uint32_t stage0 = smem_addr(smem);
    const uint32_t mbar0 = stage0 + p.n_stages * SS_STAGE_BYTES;
    uint8_t *extra = smem + (size_t)p.n_stages * SS_STAGE_BYTES + 64;
    const uint32_t row_bytes = p.n_cols * 4u;
    const uint32_t full_bytes = p.rows_per_stage * row_bytes;

    ChunkSeq seq;
    seq.n_local = blockIdx.x < p.n_chunks ? (p.n_chunks - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;
    const uint32_t n_chunks = seq.n_local;

    if (tid == 0) {
        for (uint32_t s = 0; s < p.n_stages; s++) mbar_init(mbar0 + 8u * s, 1u);
        fence_mbar_init();
    }
    uint32_t *hist = reinterpret_cast<uint32_t *>(extra);   // MODE 1: [2][16][HS]
    const uint32_t HS = ss_hist_stride(p.n_cols);
    if (MODE == 1)
        for (uint32_t i = tid; i < 2u * 16u * HS; i += T) hist[i] = 0u;
    __syncthreads();

    // the TMA producer: one thread; a full chunk is ONE bulk copy (rows are contiguous), the partial chunk at the
    // very end of the matrix is copied by the threads themselves
    auto issue = [&](uint32_t q) {
        if (q < n_chunks && seq.rows(p, q) == p.rows_per_stage) {
            const uint32_t s = q % p.n_stages;
            mbar_expect_tx(mbar0 + 8u * s, full_bytes);
            bulk_copy_g2s(stage0 + s * SS_STAGE_BYTES, p.counts + seq.row0(p, q) * p.n_cols, full_bytes, mbar0 + 8u * s);
        }
    };
    if (tid == 0)
        for (uint32_t q = 0; q < p.n_stages; q++) issue(q);

    // per-column state
    double obs[KC];
    unsigned long long isum[KC], sq_lo[KC];
    uint32_t sq_hi[KC], nlt[KC], neq[KC], vmax = 0, pre0[KC], pre1[KC];
#pragma unroll
    for (int k = 0; k < KC; k++) {
        const uint32_t col = c + (uint32_t)k * CW;
        const bool live = col < p.n_cols;
        obs[k] = (MODE == 0 && live) ? p.observed[col] : 0.0;
        pre0[k] = (MODE == 1 && live) ? p.prefix[col] : 0u;
        pre1[k] = (MODE == 1 && live) ? p.prefix[p.n_cols + col] : 0u;
        isum[k] = sq_lo[k] = 0ull; sq_hi[k] = nlt[k] = neq[k] = 0u;
    }
    const uint32_t hmask = p.shift >= 28 ? 0u : (0xffffffffu << (p.shift + 4));      // bits already decided

    for (uint32_t q = 0; q < n_chunks; q++) {
        const uint32_t s = q % p.n_stages, rows = seq.rows(p, q);
        const uint32_t *stage = reinterpret_cast<const uint32_t *>(smem + (size_t)s * SS_STAGE_BYTES);
        int timed_out = 0;
        if (rows == p.rows_per_stage) {
            uint32_t spins = 0;
            while (!mbar_try_wait(mbar0 + 8u * s, (q / p.n_stages) & 1u))
                if (++spins > SS_WAIT_SPINS) { timed_out = 1; break; }
        } else if (rows) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(smem + (size_t)s * SS_STAGE_BYTES);
            const uint32_t *src = p.counts + seq.row0(p, q) * p.n_cols;
            for (uint32_t i = tid; i < rows * p.n_cols; i += T) dst[i] = src[i];
            __syncthreads();
        }
        for (uint32_t row = r; row < rows; row += RW) {
            const uint32_t *rowp = stage + row * p.n_cols;
#pragma unroll
            for (int k = 0; k < KC; k++) {
                const uint32_t col = c + (uint32_t)k * CW;
                if (col < p.n_cols) {
                    const uint32_t v = rowp[col];
                    if (MODE == 0) {
                        const double x = (double)v;
                        const unsigned long long v2 = (unsigned long long)v * v;
                        isum[k] += v;
                        sq_lo[k] += v2;
                        sq_hi[k] += (sq_lo[k] < v2) ? 1u : 0u;   // carry out of the low 64 bits
                        nlt[k] += (x < obs[k]) ? 1u : 0u;        // searchargsorted + cmpDouble (gat/Engine.pyx:122-127, 1549-1557)
                        neq[k] += (x == obs[k]) ? 1u : 0u;
                        vmax = max(vmax, v);
                    } else {
                        const uint32_t bin = (v >> p.shift) & 15u;
                        const bool m0 = (v & hmask) == pre0[k], m1 = (v & hmask) == pre1[k];
                        // while both ranks still share their prefix one histogram serves both
                        if (m0) atomicAdd(&hist[bin * HS + col], 1u);
                        if (m1 && pre0[k] != pre1[k]) atomicAdd(&hist[(16u + bin) * HS + col], 1u);
                    }
                }
            }
        }
        // everyone is done with stage s: refill it.  (A copy that never completed -- it cannot, short of a broken
        // address -- ends the CTA's work for all of its threads at once and fails the call.)
        if (__syncthreads_or(timed_out)) {
            if (tid == 0) *p.error = 1u;
            break;
        }
        if (tid == 0) issue(q + p.n_stages);
    }

    if (MODE == 0) {
        // integers: exact and order-independent, one atomic per column, row lane and CTA.  The 128-bit sum of
        // squares: every atomic on the low word reports its own carry, whatever the order
#pragma unroll
        for (int k = 0; k < KC; k++) {
            const uint32_t col = c + (uint32_t)k * CW;
            if (col < p.n_cols && n_chunks) {
                if (isum[k]) atomicAdd(p.isum + col, isum[k]);
                unsigned long long hi = sq_hi[k];
                if (sq_lo[k]) {
                    const unsigned long long old = atomicAdd(p.sq_lo + col, sq_lo[k]);
                    hi += (old + sq_lo[k] < old) ? 1ull : 0ull;
                }
                if (hi) atomicAdd(p.sq_hi + col, hi);
                if (nlt[k]) atomicAdd(p.n_lt + col, (unsigned long long)nlt[k]);
                if (neq[k]) atomicAdd(p.n_eq + col, (unsigned long long)neq[k]);
            }
        }
        vmax = __reduce_max_sync(GATB_FULL, vmax);
        if ((tid & 31u) == 0 && vmax) atomicMax(p.vmax, vmax);
    }
    if (MODE == 1) {
        __syncthreads();
        for (uint32_t i = tid; i < 2u * 16u * HS; i += T) {
            const uint32_t v = hist[i], col = i % HS;
            if (v && col < p.n_cols) atomicAdd(p.hist + (uint64_t)(i / HS) * p.n_cols + col, v);
        }
    }
}

// between two select passes: the nibble that holds the wanted rank extends the prefix; the histograms are cleared
__global__ void __launch_bounds__(256) stats_pick_kernel(StreamStatsParams p)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2u * p.n_cols) return;
    const uint32_t w = i / p.n_cols, col = i % p.n_cols;
    const bool shared_hist = p.prefix[col] == p.prefix[p.n_cols + col];      // (before either of them is updated)
    uint32_t *h = p.hist + (uint64_t)((w && !shared_hist) ? 16u : 0u) * p.n_cols + col;
    unsigned long long rank = p.rank[i], cum = 0;
    uint32_t b = 0;
    for (; b < 16u; b++) {
        const uint32_t v = h[(uint64_t)b * p.n_cols];
        if (cum + v > rank) break;
        cum += v;
    }
    if (b > 15u) b = 15u;
    // both ranks of a column may read the shared histogram: clear only after both have (second kernel phase below)
    p.rank_out[i] = rank - cum;
    p.prefix_out[i] = p.prefix[i] | (b << p.shift);
}
__global__ void __launch_bounds__(256) stats_pick_commit_kernel(StreamStatsParams p)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2u * p.n_cols) {
        p.rank[i] = p.rank_out[i];
        p.prefix[i] = p.prefix_out[i];
        if (p.shift == 0) (i < p.n_cols ? p.q_lo : p.q_hi)[i % p.n_cols] = (double)p.prefix_out[i];
    }
    for (uint64_t j = i; j < 32ull * p.n_cols; j += (uint64_t)gridDim.x * blockDim.x) p.hist[j] = 0u;
}

template <int MODE>
static cudaError_t launch_stream_mode(cudaStream_t st, const StreamStatsParams &p, int sm_count)
{
    const int threads = MODE == 1 ? 512 : 256;
    StreamStatsParams q = p;
    uint32_t cw = 1;
    while (cw < p.n_cols && cw < (uint32_t)threads) cw <<= 1;
    q.col_width = cw;
    const uint32_t kc = (p.n_cols + cw - 1u) / cw;
    const size_t smem = stats_stream_smem(q, MODE, threads);
    const unsigned grid = (unsigned)std::min<uint64_t>(p.n_chunks, (uint64_t)sm_count * (MODE == 1 ? 1u : 2u));
    cudaError_t e;
#define GATB_SS_LAUNCH(KC_)                                                                                                   \
    e = cudaFuncSetAttribute(stats_stream_kernel<MODE, KC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    if (e != cudaSuccess) return e;                                                                                            \
    stats_stream_kernel<MODE, KC_><<<grid, threads, smem, st>>>(q);
    if (kc <= 1) { GATB_SS_LAUNCH(1) }
    else if (kc <= 2) { GATB_SS_LAUNCH(2) }
    else if (kc <= 4) { GATB_SS_LAUNCH(4) }
    else { GATB_SS_LAUNCH(8) }
#undef GATB_SS_LAUNCH
    return cudaGetLastError();
}

// can the streaming passes take this matrix?  (uint32, rows short enough for a stage, histograms that fit)
bool stats_stream_fits(const void *counts, uint64_t n_samples, uint32_t n_cols, size_t smem_optin)
{
    if (((uintptr_t)counts & 15u) != 0 || n_samples == 0 || n_samples >= (1ull << 32)) return false;
    if ((uint64_t)n_cols * 4u * 4u > SS_STAGE_BYTES) return false;                 // at least 4 rows per stage
    StreamStatsParams p;
    p.n_cols = n_cols; p.n_stages = 2;
    return stats_stream_smem(p, 1, 512) <= smem_optin;
}

void stats_stream_geometry(StreamStatsParams &p)
{
    const uint32_t row_bytes = p.n_cols * 4u;
    p.rows_per_stage = std::max(4u, (SS_STAGE_BYTES / row_bytes) & ~3u);            // multiple of 4: chunks start 16-byte aligned
    p.n_chunks = (uint32_t)((p.n_samples + p.rows_per_stage - 1u) / p.rows_per_stage);
}

cudaError_t launch_stats_stream_pass1(cudaStream_t st, StreamStatsParams p, int sm_count)
{
    p.n_stages = 3;
    return launch_stream_mode<0>(st, p, sm_count);
}

// one 4-bit select pass at p.shift (prefix / rank / hist as the previous pass left them)
cudaError_t launch_stats_stream_select(cudaStream_t st, StreamStatsParams p, int sm_count)
{
    p.n_stages = 2;
    cudaError_t e = launch_stream_mode<1>(st, p, sm_count);
    if (e != cudaSuccess) return e;
    const unsigned nb = (2u * p.n_cols + 255u) / 256u;
    stats_pick_kernel<<<nb, 256, 0, st>>>(p);
    stats_pick_commit_kernel<<<nb, 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace gatb
