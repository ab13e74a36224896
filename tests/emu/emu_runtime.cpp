// tests/emu/emu_runtime.cpp -- TEST INFRASTRUCTURE (see tests/emu/include/cuda_runtime.h): the SIMT emulator behind the
// shim header.  One fiber per CUDA thread of the running block; a fiber runs until it reaches a rendezvous (warp
// collective or __syncthreads) that is not complete yet, then the scheduler resumes the next one.  Any schedule of
// that kind is one the hardware could produce, so a kernel that is correct under CUDA's rules computes the same
// result here; a kernel whose lanes disagree about a collective (one of them never arrives) is reported as a
// deadlock instead of hanging.
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>

#include <mutex>
#include <vector>

namespace gatb_emu {

Thread *cur = nullptr;
Block blk;

// ------------------------------------------------------------------------------------------------ context switch
struct Ctx { void *rsp; };
extern "C" void gatb_emu_switch(Ctx *from, Ctx *to);
#if defined(__x86_64__)
asm(R"(
.text
.globl gatb_emu_switch
.type gatb_emu_switch,@function
gatb_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq (%rsi), %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size gatb_emu_switch,.-gatb_emu_switch
)");
#else
#error "the SIMT emulator's context switch is written for x86-64"
#endif

enum Wait { RUNNABLE = 0, WAIT_WARP = 1, WAIT_BLOCK = 2 };

struct WarpState {
    uint32_t alive_mask = 0, alive = 0, arrived = 0;
    uint64_t gen = 0;
    uint64_t slot[2][32];
    uint32_t took_part[2] = {0, 0};
};

struct Fiber {
    Thread th;
    Ctx ctx;
    uint8_t *stack = nullptr;
    bool done = false;
    int wait = RUNNABLE;
    uint64_t wait_gen = 0;
};

static const size_t STACK_BYTES = 256u << 10;
static std::vector<Fiber> fibers;
static std::vector<WarpState> warps;
static Ctx sched_ctx;
static Fiber *cur_fiber = nullptr;
static Body *cur_body = nullptr;
static uint32_t blk_alive = 0, blk_arrived = 0;
static uint64_t blk_gen = 0;
static int blk_or_acc = 0, blk_or_result[2] = {0, 0};
static std::mutex launch_mutex;

static void yield_to_scheduler() { gatb_emu_switch(&cur_fiber->ctx, &sched_ctx); }

static void complete_warp(WarpState &w)
{
    w.took_part[w.gen & 1u] = w.alive_mask;
    w.arrived = 0;
    w.gen++;
}

const uint64_t *warp_exchange(uint64_t v, uint32_t *alive)
{
    Fiber *f = cur_fiber;
    WarpState &w = *f->th.warp;
    const uint64_t g = w.gen;
    uint64_t *buf = w.slot[g & 1u];
    buf[f->th.lane] = v;
    w.arrived++;
    if (w.arrived == w.alive) complete_warp(w);
    else {
        f->wait = WAIT_WARP;
        f->wait_gen = g;
        while (w.gen == g) yield_to_scheduler();
        f->wait = RUNNABLE;
    }
    *alive = w.took_part[g & 1u];
    return buf;
}

static void complete_block()
{
    blk_or_result[blk_gen & 1u] = blk_or_acc;
    blk_or_acc = 0;
    blk_arrived = 0;
    blk_gen++;
}

int block_sync_or(int pred)
{
    const uint64_t g = blk_gen;
    if (pred) blk_or_acc = 1;
    block_sync();
    return blk_or_result[g & 1u];
}

void block_sync()
{
    Fiber *f = cur_fiber;
    const uint64_t g = blk_gen;
    blk_arrived++;
    if (blk_arrived == blk_alive) complete_block();
    else {
        f->wait = WAIT_BLOCK;
        f->wait_gen = g;
        while (blk_gen == g) yield_to_scheduler();
        f->wait = RUNNABLE;
    }
}

void yield() { yield_to_scheduler(); }

static void fiber_exit(Fiber *f)
{
    // a thread that has left the kernel takes no part in later collectives; one that was the last missing makes
    // them complete
    f->done = true;
    WarpState &w = *f->th.warp;
    w.alive--;
    w.alive_mask &= ~(1u << f->th.lane);
    if (w.alive > 0 && w.arrived == w.alive) complete_warp(w);
    blk_alive--;
    if (blk_alive > 0 && blk_arrived == blk_alive) complete_block();
}

extern "C" void gatb_emu_fiber_main()
{
    Fiber *f = cur_fiber;
    cur_body->run();
    fiber_exit(f);
    yield_to_scheduler();
    abort();                                        // a finished fiber is never resumed
}

static bool runnable(const Fiber &f)
{
    if (f.done) return false;
    if (f.wait == WAIT_WARP) return f.th.warp->gen != f.wait_gen;
    if (f.wait == WAIT_BLOCK) return blk_gen != f.wait_gen;
    return true;
}

static void run_block(uint32_t nthreads)
{
    if (fibers.size() < nthreads) fibers.resize(nthreads);
    const uint32_t nwarps = (nthreads + 31u) / 32u;
    warps.assign(nwarps, WarpState());
    for (uint32_t t = 0; t < nthreads; t++) {
        Fiber &f = fibers[t];
        if (f.stack == nullptr) {
            void *m = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (m == MAP_FAILED) { fprintf(stderr, "gatb_emu: cannot map a fiber stack\n"); abort(); }
            f.stack = (uint8_t *)m;
        }
        f.done = false;
        f.wait = RUNNABLE;
        f.th.tid.x = t % blk.bdim.x;
        f.th.tid.y = (t / blk.bdim.x) % blk.bdim.y;
        f.th.tid.z = t / (blk.bdim.x * blk.bdim.y);
        f.th.lane = t & 31u;
        f.th.warp = &warps[t >> 5];
        warps[t >> 5].alive++;
        warps[t >> 5].alive_mask |= 1u << (t & 31u);
        // initial frame: six callee-saved registers, then the entry point as the return address, then a slot that
        // leaves the stack aligned as after a call
        uint64_t *sp = (uint64_t *)(f.stack + STACK_BYTES);
        *--sp = 0;
        *--sp = (uint64_t)(uintptr_t)&gatb_emu_fiber_main;
        for (int r = 0; r < 6; r++) *--sp = 0;
        f.ctx.rsp = sp;
    }
    blk_alive = nthreads;
    blk_arrived = 0;
    blk_or_acc = 0;
    uint32_t live = nthreads;
    while (live > 0) {
        bool progress = false;
        for (uint32_t t = 0; t < nthreads; t++) {
            Fiber &f = fibers[t];
            if (!runnable(f)) continue;
            cur_fiber = &f;
            cur = &f.th;
            gatb_emu_switch(&sched_ctx, &f.ctx);
            progress = true;
            if (f.done) live--;
        }
        if (!progress) {
            fprintf(stderr, "gatb_emu: deadlock in block (%u,%u,%u): %u threads wait for a collective that the others never reach\n",
                    blk.bid.x, blk.bid.y, blk.bid.z, live);
            for (uint32_t t = 0; t < nthreads && t < 64; t++)
                if (!fibers[t].done) fprintf(stderr, "  thread %u waits on %s\n", t, fibers[t].wait == WAIT_WARP ? "its warp" : "the block");
            abort();
        }
    }
    cur = nullptr;
    cur_fiber = nullptr;
}

void launch_body(dim3 grid, dim3 block, size_t smem, Body &body)
{
    std::lock_guard<std::mutex> lock(launch_mutex);
    const uint32_t nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024u || (uint64_t)grid.x * grid.y * grid.z == 0) {
        fprintf(stderr, "gatb_emu: invalid launch configuration\n");
        abort();
    }
    std::vector<uint8_t> dyn(smem + 64u);
    cur_body = &body;
    blk.bdim = block;
    blk.gdim = grid;
    blk.dyn_smem = (uint8_t *)(((uintptr_t)dyn.data() + 15u) & ~(uintptr_t)15u);
    for (uint32_t z = 0; z < grid.z; z++)
        for (uint32_t y = 0; y < grid.y; y++)
            for (uint32_t x = 0; x < grid.x; x++) {
                memset(dyn.data(), 0xA5, dyn.size());           // shared memory starts out as garbage
                blk.bid.x = x; blk.bid.y = y; blk.bid.z = z;
                run_block(nthreads);
            }
    cur_body = nullptr;
}

}  // namespace gatb_emu

// ---------------------------------------------------------------------------------------------------- runtime API
struct gatb_emu_stream { int unused; };
struct gatb_emu_event { int unused; };
struct gatb_emu_pool { int unused; };
static gatb_emu_pool the_pool;

static int emu_sms()
{
    const char *v = getenv("GATB_EMU_SMS");
    const int n = v ? atoi(v) : 0;
    return n > 0 ? n : 2;
}

cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "SIMT emulation (tests only)");
    p->totalGlobalMem = (size_t)8 << 30;
    p->sharedMemPerBlockOptin = 227u << 10;
    p->multiProcessorCount = emu_sms();
    p->major = 10;
    p->minor = 0;
    return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorNotSupported ? "not supported by the emulation" : "emulated CUDA error"; }

static cudaError_t emu_alloc(void **p, size_t n)
{
    void *m = nullptr;
    if (posix_memalign(&m, 256, n ? n : 1) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
    memset(m, 0xCD, n);                             // device allocations start out as garbage
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void **p, size_t n) { return emu_alloc(p, n); }
cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return emu_alloc(p, n); }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaFreeAsync(void *p, cudaStream_t) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { return emu_alloc(p, n); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t)
{
    if (n) memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < height; r++) memmove((uint8_t *)dst + r * dpitch, (const uint8_t *)src + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { if (n) memset(p, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned) { *st = new gatb_emu_stream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t st) { delete st; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *ev) { *ev = new gatb_emu_event(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *ev, unsigned) { *ev = new gatb_emu_event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t ev) { delete ev; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *pool, int) { *pool = &the_pool; return cudaSuccess; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
cudaError_t cudaMemPoolGetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *value) { *(uint64_t *)value = 0; return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaErrorNotSupported; }

// (the hardware micro-benchmarks have nothing to emulate)
extern "C" int gatb_microbench(int, int, uint64_t, int, double *) { return -2; }

// tells a loader what it got: the product's loader (gat_b200/_lib.py) refuses a library that exports this symbol
extern "C" int gatb_emulation_marker() { return 1; }
