// tests/emu/include/ptx.cuh -- TEST INFRASTRUCTURE: what gat_b200/csrc/ptx.cuh does with inline PTX, in plain C++ for
// the SIMT emulation (tests/emu/include/cuda_runtime.h).  The emulated build of count.cu finds this file first.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gatb {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <class T> __device__ __forceinline__ T emu_lds(uint32_t addr) { T v; memcpy(&v, gatb_emu::shared_ptr(addr), sizeof(T)); return v; }
template <class T> __device__ __forceinline__ void emu_sts(uint32_t addr, T v) { memcpy(gatb_emu::shared_ptr(addr), &v, sizeof(T)); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) { return emu_lds<uint4>(addr); }
__device__ __forceinline__ uint2 lds64(uint32_t addr) { return emu_lds<uint2>(addr); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { return emu_lds<uint32_t>(addr); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) { emu_sts(addr, v); }
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { emu_sts(addr, v); }
__device__ __forceinline__ void red_add_shared(uint32_t addr, uint32_t v) { emu_sts(addr, emu_lds<uint32_t>(addr) + v); }
__device__ __forceinline__ uint2 ldg_nc_u2(uint64_t addr) { uint2 v; memcpy(&v, (const void *)addr, sizeof(v)); return v; }
__device__ __forceinline__ uint4 ldg_nc_u4(uint64_t addr) { uint4 v; memcpy(&v, (const void *)addr, sizeof(v)); return v; }
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n) { return n >= 32u ? 0u : v << n; }
__device__ __forceinline__ uint32_t acc_cell(uint32_t wy, uint32_t acc_addr) { return (wy & 0xfffu) * 4u + acc_addr; }
__device__ __forceinline__ void discard_l2_line(uint64_t addr)
{
    memset((void *)addr, 0xDD, 128);                // "undefined afterwards": make a later read of it visible
}

}  // namespace gatb
