// tests/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE, not part of the product.
//
// A stand-in for the CUDA runtime header that lets g++ compile the library's .cu sources UNCHANGED for the host, with
// the SIMT execution model emulated (tests/emu/emu_runtime.cpp): every CUDA thread of a block is a fiber, warp
// collectives (__shfl_sync, __ballot_sync, __reduce_*_sync, __syncwarp) and __syncthreads are rendezvous points of
// those fibers, blocks run one after the other, "device" memory is host memory, streams and events are no-ops.
// tests/emu/build_emu.py rewrites only the `kernel<<<grid, block, smem, stream>>>(args)` launch statements.
//
// What it is for: the CPU test suite (tests/test_emu_parity.py) runs the kernels' own source code against the oracle
// on a machine without a GPU, bit for bit, through the same C ABI as the GPU tests.  It is orders of magnitude slower
// than the oracle and is never built, loaded or referenced by gat_b200/ (tests/test_abi.py checks that); the product
// has no CPU path.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <cmath>
#include <type_traits>

#ifndef GATB_EMU
#define GATB_EMU 1
#endif

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

using std::isfinite;

// ------------------------------------------------------------------------------------------------ vector types
struct alignas(8) uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 v; v.x = x; v.y = y; return v; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

// ------------------------------------------------------------------------------------------------ runtime API
typedef enum cudaError {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorNotSupported = 801
} cudaError_t;

typedef struct gatb_emu_stream *cudaStream_t;
typedef struct gatb_emu_event *cudaEvent_t;
typedef struct gatb_emu_pool *cudaMemPool_t;
struct cudaIpcMemHandle_t { char reserved[64]; };

enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4, cudaMemPoolAttrReservedMemCurrent = 5, cudaMemPoolAttrUsedMemCurrent = 7 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

struct cudaDeviceProp {
    char name[256];
    size_t totalGlobalMem;
    size_t sharedMemPerBlockOptin;
    int multiProcessorCount;
    int major, minor;
};

cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t st);
cudaError_t cudaFree(void *p);
cudaError_t cudaFreeAsync(void *p, cudaStream_t st);
cudaError_t cudaMallocHost(void **p, size_t n);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind kind, cudaStream_t st = nullptr);
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind kind, cudaStream_t st = nullptr);
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t st = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t *ev);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *ev, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t ev);
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t st = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t ev);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *pool, int device);
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t pool, cudaMemPoolAttr attr, void *value);
cudaError_t cudaMemPoolGetAttribute(cudaMemPool_t pool, cudaMemPoolAttr attr, void *value);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)(void *)p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMallocHost((void **)(void *)p, n); }
template <class T> static inline cudaError_t cudaMallocAsync(T **p, size_t n, cudaStream_t st) { return cudaMallocAsync((void **)(void *)p, n, st); }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ------------------------------------------------------------------------------------------------ SIMT emulation
namespace gatb_emu {

struct WarpState;
struct Thread {                 // one CUDA thread = one fiber
    uint3 tid;
    uint32_t lane;
    WarpState *warp;
};
struct Block {
    uint3 bid;
    dim3 bdim, gdim;
    uint8_t *dyn_smem;          // extern __shared__ array of the running block
};
extern Thread *cur;
extern Block blk;

// rendezvous of the calling thread's warp: deposits v, returns the 32 deposited values (valid until the caller's next
// collective); *alive = lanes that took part (lanes that have left the kernel do not)
const uint64_t *warp_exchange(uint64_t v, uint32_t *alive);
void block_sync();
int block_sync_or(int pred);    // __syncthreads_or
void yield();                   // let the other threads of the block run (spin loops on shared state)

// run fn() as every thread of a grid (blocks one after the other)
struct Body { virtual void run() = 0; virtual ~Body() {} };
void launch_body(dim3 grid, dim3 block, size_t smem, Body &body);
template <class F> struct BodyOf : Body { F &f; explicit BodyOf(F &f_) : f(f_) {} void run() override { f(); } };
template <class F> static inline void launch(dim3 grid, dim3 block, size_t smem, F f)
{
    BodyOf<F> b(f);
    launch_body(grid, block, smem, b);
}

// shared-memory "addresses" (the 32-bit window the kernels' ld.shared / st.shared helpers use)
static inline uint32_t shared_addr(const void *p) { return (uint32_t)((const uint8_t *)p - blk.dyn_smem) + 1024u; }
static inline void *shared_ptr(uint32_t addr) { return blk.dyn_smem + (addr - 1024u); }

template <class T> static inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "collective operand wider than 8 bytes");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T> static inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace gatb_emu

#define threadIdx (gatb_emu::cur->tid)
#define blockIdx (gatb_emu::blk.bid)
#define blockDim (gatb_emu::blk.bdim)
#define gridDim (gatb_emu::blk.gdim)

static inline void __syncthreads() { gatb_emu::block_sync(); }
static inline int __syncthreads_or(int pred) { return gatb_emu::block_sync_or(pred); }
static inline void __syncwarp(unsigned = 0xffffffffu)
{
    uint32_t alive;
    gatb_emu::warp_exchange(0, &alive);
}

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    uint32_t alive;
    const uint64_t *all = gatb_emu::warp_exchange(gatb_emu::to_bits(v), &alive);
    const int lane = (int)gatb_emu::cur->lane, base = lane & ~(width - 1);
    return gatb_emu::from_bits<T>(all[base + (src & (width - 1))]);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32)
{
    uint32_t alive;
    const uint64_t *all = gatb_emu::warp_exchange(gatb_emu::to_bits(v), &alive);
    const int lane = (int)gatb_emu::cur->lane, base = lane & ~(width - 1);
    const int s = lane - (int)delta;
    return gatb_emu::from_bits<T>(all[s < base ? lane : s]);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32)
{
    uint32_t alive;
    const uint64_t *all = gatb_emu::warp_exchange(gatb_emu::to_bits(v), &alive);
    const int lane = (int)gatb_emu::cur->lane, base = lane & ~(width - 1);
    const int s = lane + (int)delta;
    return gatb_emu::from_bits<T>(all[s > base + width - 1 ? lane : s]);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
    uint32_t alive;
    const uint64_t *all = gatb_emu::warp_exchange(gatb_emu::to_bits(v), &alive);
    const int lane = (int)gatb_emu::cur->lane, base = lane & ~(width - 1);
    const int s = lane ^ m;
    return gatb_emu::from_bits<T>(all[s > base + width - 1 ? lane : s]);
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
    uint32_t alive;
    const uint64_t *all = gatb_emu::warp_exchange(pred ? 1u : 0u, &alive);
    unsigned m = 0;
    for (int l = 0; l < 32; l++)
        if (((alive >> l) & 1u) && all[l]) m |= 1u << l;
    return m;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
static inline int __all_sync(unsigned mask, int pred)
{
    uint32_t alive;
    const uint64_t *all = gatb_emu::warp_exchange(pred ? 1u : 0u, &alive);
    (void)mask;
    for (int l = 0; l < 32; l++)
        if (((alive >> l) & 1u) && !all[l]) return 0;
    return 1;
}

namespace gatb_emu {
template <class T, class Op> static inline T warp_fold(T v, Op op)
{
    uint32_t alive;
    const uint64_t *all = warp_exchange(to_bits(v), &alive);
    bool first = true;
    T acc = v;
    for (int l = 0; l < 32; l++)
        if ((alive >> l) & 1u) {
            const T x = from_bits<T>(all[l]);
            acc = first ? x : op(acc, x);
            first = false;
        }
    return acc;
}
}  // namespace gatb_emu
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return gatb_emu::warp_fold(v, [](unsigned a, unsigned b) { return a + b; }); }
static inline int __reduce_add_sync(unsigned, int v) { return (int)gatb_emu::warp_fold((unsigned)v, [](unsigned a, unsigned b) { return a + b; }); }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return gatb_emu::warp_fold(v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
static inline int __reduce_min_sync(unsigned, int v) { return gatb_emu::warp_fold(v, [](int a, int b) { return a < b ? a : b; }); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return gatb_emu::warp_fold(v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
static inline int __reduce_max_sync(unsigned, int v) { return gatb_emu::warp_fold(v, [](int a, int b) { return a > b ? a : b; }); }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return gatb_emu::warp_fold(v, [](unsigned a, unsigned b) { return a | b; }); }
static inline unsigned __reduce_and_sync(unsigned, unsigned v) { return gatb_emu::warp_fold(v, [](unsigned a, unsigned b) { return a & b; }); }

// ------------------------------------------------------------------------------------------------ scalar intrinsics
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned __float_as_uint(float f) { return gatb_emu::from_bits<unsigned>(gatb_emu::to_bits(f)); }
static inline long long __double_as_longlong(double d) { return gatb_emu::from_bits<long long>(gatb_emu::to_bits(d)); }
static inline double __longlong_as_double(long long v) { return gatb_emu::from_bits<double>(gatb_emu::to_bits(v)); }
static inline size_t __cvta_generic_to_shared(const void *p) { return gatb_emu::shared_addr(p); }
static inline size_t __cvta_generic_to_global(const void *p) { return (size_t)p; }

// CUDA's min / max take mixed integer types (the signed operand is converted like in a comparison)
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b)
{
    typedef typename std::common_type<A, B>::type T;
    return (T)a < (T)b ? (T)a : (T)b;
}
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b)
{
    typedef typename std::common_type<A, B>::type T;
    return (T)a > (T)b ? (T)a : (T)b;
}

// atomics: one fiber runs at a time, so plain read-modify-write is atomic
template <class T, class U> static inline T atomicAdd(T *p, U v) { const T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { const T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicAnd(T *p, U v) { const T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { const T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { const T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { const T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T *p, U cmp, V v) { const T o = *p; if (o == (T)cmp) *p = (T)v; return o; }
