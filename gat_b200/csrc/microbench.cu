// microbench.cu -- the two machine peaks the counting kernel is measured against (bench.py `roofline`):
//   * L2 read bandwidth: every SM streams a buffer that fits in L2 (default 48 MB) with 16-byte read-only loads,
//     many passes; the first pass warms L2 and is not timed
//   * warp-instruction issue rate: four integer (IADD3) and four FP32 (FFMA) chains per thread, interleaved, no memory,
//     all four schedulers of every SM busy (16 warps per scheduler) -- compared against 4 x SMs x clock (an
//     integer-only chain measures the ALU pipe: half of it)
// HBM copy bandwidth comes from the driver-written MEASURED_PEAKS.json; these two are measured the same way
// (best of several runs, CUDA events, on the GPU the bench runs on) by gatb_microbench().
#include "../../include/gat_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace gatb {

__global__ void __launch_bounds__(1024) l2_read_kernel(const uint4 *__restrict__ buf, uint64_t n_vec, int passes, uint32_t *sink)
{
    uint32_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++) {
        // a different starting offset every pass so that the compiler cannot hoist the loads
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            uint4 v;
            asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;        // (never true for the zero-filled buffer; keeps the loads alive)
}

constexpr int ISSUE_UNROLL = 64;                // instructions per chain and loop iteration

__global__ void __launch_bounds__(1024) issue_kernel(int iters, uint32_t seed, uint32_t *sink)
{
    // four integer chains (IADD3, ALU pipe) interleaved with four FP32 chains (FFMA, FMA pipe): one pipe alone takes
    // only every other issue slot of a scheduler, the two together can fill all of them
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 ^ 11u;
    float f0 = (float)(seed & 7u) + 1.0f, f1 = f0 + 0.5f, f2 = f0 + 0.25f, f3 = f0 + 0.125f;
    const float m = 1.0f + (float)(seed & 1u) * 1e-7f, c = (float)(seed & 3u) * 1e-9f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < ISSUE_UNROLL; u++) {
            a0 += a1; f0 = fmaf(f0, m, c);
            a1 += a2; f1 = fmaf(f1, m, c);
            a2 += a3; f2 = fmaf(f2, m, c);
            a3 += a0; f3 = fmaf(f3, m, c);
        }
    }
    const uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ __float_as_uint(f0 + f1 + f2 + f3);
    if (r == 0x9e3779b9u) *sink = r;
}

}  // namespace gatb

// out[0] = GB/s (which 0: L2 reads) or 1e9 warp instructions/s (which 1: issue), out[1] = milliseconds of the
// best run, out[2] = work of one run (bytes or warp instructions), out[3] = SM count
extern "C" int gatb_microbench(int device, int which, uint64_t bytes, int repeats, double *out)
{
    if (!out || repeats < 1) return GATB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return GATB_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GATB_ERR_CUDA;
    const int sms = prop.multiProcessorCount;
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return GATB_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    uint32_t *sink = nullptr;
    cudaMalloc(&sink, 4);
    double best_ms = 1e30, work = 0;
    int rc = GATB_OK;
    if (which == 0) {
        if (bytes == 0) bytes = 48ull << 20;
        const uint64_t n_vec = bytes / 16;
        uint4 *buf = nullptr;
        if (cudaMalloc(&buf, n_vec * 16) != cudaSuccess) rc = GATB_ERR_CUDA;
        else {
            cudaMemsetAsync(buf, 0, n_vec * 16, st);
            const int passes = 40;
            gatb::l2_read_kernel<<<sms, 1024, 0, st>>>(buf, n_vec, 2, sink);          // warm L2
            for (int r = 0; r < repeats; r++) {
                cudaEventRecord(e0, st);
                gatb::l2_read_kernel<<<sms, 1024, 0, st>>>(buf, n_vec, passes, sink);
                cudaEventRecord(e1, st);
                if (cudaEventSynchronize(e1) != cudaSuccess) { rc = GATB_ERR_CUDA; break; }
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best_ms) best_ms = ms;
            }
            work = (double)n_vec * 16.0 * passes;
            cudaFree(buf);
        }
    } else if (which == 1) {
        const int iters = 4000;
        const int blocks = sms * 2;                 // 2 x 1024 threads = 64 warps per SM = 16 per scheduler
        gatb::issue_kernel<<<blocks, 1024, 0, st>>>(10, 1u, sink);
        for (int r = 0; r < repeats; r++) {
            cudaEventRecord(e0, st);
            gatb::issue_kernel<<<blocks, 1024, 0, st>>>(iters, (uint32_t)r, sink);
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess) { rc = GATB_ERR_CUDA; break; }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best_ms) best_ms = ms;
        }
        work = (double)blocks * 32.0 * iters * gatb::ISSUE_UNROLL * 8.0;   // warp instructions (4 IADD3 + 4 FFMA per unrolled step; loop overhead not counted)
    } else rc = GATB_ERR_INVALID;
    if (rc == GATB_OK && cudaGetLastError() != cudaSuccess) rc = GATB_ERR_CUDA;
    if (rc == GATB_OK) {
        out[0] = work / (best_ms * 1e-3) / 1e9;
        out[1] = best_ms; out[2] = work; out[3] = (double)sms;
    }
    cudaFree(sink);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return rc;
}
