// tests/emu/include/cub -- TEST INFRASTRUCTURE: host stand-ins for the few cub entry points the library calls
// (see tests/emu/include/cuda_runtime.h).  Same contracts: a null temp-storage pointer asks for the size; scans may
// run in place; SortPairs is a stable LSD sort on the key bits [begin_bit, end_bit) that leaves its result in the
// buffer DoubleBuffer::Current() names afterwards.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace cub {

struct Max {
    template <class T> T operator()(const T &a, const T &b) const { return a > b ? a : b; }
};
struct Sum {
    template <class T> T operator()(const T &a, const T &b) const { return a + b; }
};

struct DeviceScan {
    template <class In, class Out>
    static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, In in, Out out, int n, cudaStream_t = nullptr)
    {
        if (tmp == nullptr) { bytes = 16; return cudaSuccess; }
        typedef typename std::remove_cv<typename std::remove_reference<decltype(out[0])>::type>::type T;
        T run = 0;
        for (int i = 0; i < n; i++) { const T v = (T)in[i]; out[i] = run; run = (T)(run + v); }
        return cudaSuccess;
    }
    template <class In, class Out, class Op>
    static cudaError_t InclusiveScan(void *tmp, size_t &bytes, In in, Out out, Op op, int n, cudaStream_t = nullptr)
    {
        if (tmp == nullptr) { bytes = 16; return cudaSuccess; }
        typedef typename std::remove_cv<typename std::remove_reference<decltype(out[0])>::type>::type T;
        T run = 0;
        for (int i = 0; i < n; i++) { const T v = (T)in[i]; run = i ? op(run, v) : v; out[i] = run; }
        return cudaSuccess;
    }
};

template <class T> struct DoubleBuffer {
    T *d_buffers[2];
    int selector;
    DoubleBuffer(T *a, T *b) : selector(0) { d_buffers[0] = a; d_buffers[1] = b; }
    T *Current() { return d_buffers[selector]; }
    T *Alternate() { return d_buffers[selector ^ 1]; }
};

struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void *tmp, size_t &bytes, DoubleBuffer<K> &keys, DoubleBuffer<V> &vals, int n,
                                 int begin_bit = 0, int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr)
    {
        if (tmp == nullptr) { bytes = 16; return cudaSuccess; }
        const int bits = end_bit - begin_bit;
        const K mask = bits >= (int)sizeof(K) * 8 ? ~(K)0 : (K)((((K)1) << bits) - 1);
        std::vector<int> order((size_t)n);
        std::iota(order.begin(), order.end(), 0);
        const K *k = keys.Current();
        const V *v = vals.Current();
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ((k[a] >> begin_bit) & mask) < ((k[b] >> begin_bit) & mask); });
        K *ko = keys.Alternate();
        V *vo = vals.Alternate();
        for (int i = 0; i < n; i++) { ko[i] = k[order[(size_t)i]]; vo[i] = v[order[(size_t)i]]; }
        keys.selector ^= 1;
        vals.selector ^= 1;
        return cudaSuccess;
    }
};

// block-wide exclusive sum: every thread of the block calls it (thread = element)
template <class T, int BLOCK> struct BlockScan {
    struct TempStorage { T v[BLOCK]; };
    TempStorage &t;
    explicit BlockScan(TempStorage &t_) : t(t_) {}
    void ExclusiveSum(T in, T &out, T &total)
    {
        const unsigned me = threadIdx.x;
        t.v[me] = in;
        __syncthreads();
        T run = 0, all = 0;
        for (unsigned i = 0; i < (unsigned)BLOCK; i++) { if (i == me) run = all; all = (T)(all + t.v[i]); }
        out = run;
        total = all;
        __syncthreads();
    }
};

}  // namespace cub
