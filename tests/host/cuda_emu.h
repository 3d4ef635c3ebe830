// cuda_emu.h -- TEST INFRASTRUCTURE: runs the CUDA source of the product's SIMT kernels on the CPU.
//
// tests/host/emu_build.py copies a product source (csrc/dq_env.cu, the fp32 half of csrc/dq_qnet.cu), rewrites only the
// `<<<...>>>` launch syntax and the dynamic shared-memory declaration, and compiles the copy with
// `g++ -include tests/host/cuda_emu.h`: the SAME kernel code the GPU runs (templates, shared-memory structs, warp shuffles,
// ballots, barriers) becomes a host library whose
// "device pointers" are host pointers.  Every CUDA thread of a CTA is a cooperative fiber (ucontext); a warp-level
// primitive (__shfl*_sync, __ballot_sync, __any_sync, __syncwarp) or __syncthreads parks the fiber until every
// participating lane / thread has arrived, exactly the rendez-vous the hardware performs, so the kernel's control
// flow, lane roles and shared-memory hand-offs execute unchanged.  CTAs run one after another.  A rendez-vous that
// can never complete (a barrier under divergent control flow) is reported as a deadlock instead of hanging.
//
// This exists so that kernel changes can be checked bit-for-bit against the oracle without a GPU
// (tests/test_env_emulated.py).  It is never part of the product: deepq_decoding_b200/ does not reference it, and
// nothing here is timed or shipped.  It says nothing about performance.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ucontext.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
// (every standard header the kernel source includes is pulled in above: libstdc++ spells some attributes
//  __noinline__ / __forceinline__-style, so the specifier macros below must come after them)

// ---- declaration specifiers ---------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static          // CTAs run one at a time, so one static copy is the CTA's copy

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace dq_emu {

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 256 * 1024;

struct WarpSync {
    uint64_t vals[32];
    uint64_t snap[2][32];
    uint32_t arrived;
    int gen;
};

struct Fiber {
    ucontext_t ctx;
    bool done;
    char* stack;
};

struct Cta {
    Fiber th[kMaxThreads];
    WarpSync warps[kMaxThreads / 32];
    ucontext_t sched;
    int nthreads, ndone, cur;
    int bar_count, bar_gen;
    int nbar_count[16], nbar_gen[16];      // named barriers (bar.sync id, count)
    long progress;
    dim3 bid, bdim, gdim;
    unsigned char* smem;
    const std::function<void()>* body;
};

inline Cta*& cta() { static Cta* c = nullptr; return c; }

inline void yield() { Cta& c = *cta(); swapcontext(&c.th[c.cur].ctx, &c.sched); }

inline void entry() {
    Cta& c = *cta();
    (*c.body)();
    c.th[c.cur].done = true;
    c.ndone++;
    c.progress++;
}   // uc_link returns to the scheduler

inline uint32_t lanes_of_warp(const Cta& c, int w) {
    const int n = c.nthreads - w * 32;
    return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}

// all lanes named in `mask` deposit a value and get the warp's 32 deposited values back
inline const uint64_t* warp_collect(uint32_t mask, uint64_t v) {
    Cta& c = *cta();
    const int t = c.cur, w = t >> 5, lane = t & 31;
    WarpSync& ws = c.warps[w];
    mask &= lanes_of_warp(c, w);
    if (!((mask >> lane) & 1u)) { fprintf(stderr, "dq_emu: lane %d calls a warp primitive with mask %08x that excludes it\n", lane, mask); abort(); }
    const int gen = ws.gen;
    ws.vals[lane] = v;
    ws.arrived |= 1u << lane;
    if ((ws.arrived & mask) == mask) {
        memcpy(ws.snap[gen & 1], ws.vals, sizeof(ws.vals));
        ws.arrived &= ~mask;
        ws.gen++;
        c.progress++;
    } else {
        while (ws.gen == gen) yield();
    }
    return ws.snap[gen & 1];
}

inline void block_barrier() {
    Cta& c = *cta();
    const int gen = c.bar_gen;
    c.bar_count++;
    for (;;) {
        if (c.bar_gen != gen) return;
        if (c.bar_count >= c.nthreads - c.ndone) { c.bar_count = 0; c.bar_gen++; c.progress++; return; }
        yield();
    }
}

// bar.sync id, count / bar.arrive id, count (id 0 is the barrier __syncthreads uses: a kernel may address it by number as long as the two
// uses do not overlap in time; the emulation keeps separate counters for them).  The hardware executes barrier instructions per WARP, as if all its threads were active:
// a warp arrives as one unit of 32.  The fibers of a warp therefore first meet (a warp-wide rendez-vous), one of them adds the warp's
// 32 arrivals, and -- for bar.sync -- all of them wait for the phase to complete.  Callers must be warp-converged, as on the GPU.
inline void named_barrier_impl(int id, int count, bool wait) {
    Cta& c = *cta();
    if (id < 0 || id > 15 || count % 32) { fprintf(stderr, "dq_emu: bad named barrier (%d, %d)\n", id, count); abort(); }
    const int w = c.cur >> 5, lane = c.cur & 31;
    const uint32_t lanes = lanes_of_warp(c, w);
    int gen = c.nbar_gen[id];
    const uint64_t* v = warp_collect(lanes, (uint64_t)(unsigned)gen);       // every lane sees the phase the warp's LOWEST lane sampled
    int low = 0;
    while (!((lanes >> low) & 1u)) ++low;
    gen = (int)v[low];
    if (lane == low) {
        c.nbar_count[id] += 32;
        if (c.nbar_count[id] >= count) { c.nbar_count[id] = 0; c.nbar_gen[id]++; }
        c.progress++;
    }
    warp_collect(lanes, 0);                                                 // the arrival is registered before any lane moves on
    if (wait)
        while (c.nbar_gen[id] == gen) yield();
}
inline void named_barrier(int id, int count) { named_barrier_impl(id, count, true); }
inline void named_barrier_arrive(int id, int count) { named_barrier_impl(id, count, false); }

inline int idle_limit() { static int m = -1; if (m < 0) { const char* e = getenv("DQ_EMU_IDLE"); m = e ? atoi(e) : 2; } return m; }
inline int sched_mode() { static int m = -1; if (m < 0) { const char* e = getenv("DQ_EMU_SCHED"); m = e ? atoi(e) : 0; } return m; }
inline unsigned sched_rand() { static unsigned s = 0; if (!s) s = 2654435761u * (unsigned)(sched_mode() + 1); s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    static Cta* c = nullptr;
    if (!c) {
        c = new Cta();
        for (int i = 0; i < kMaxThreads; ++i) c->th[i].stack = nullptr;
    }
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads > kMaxThreads) { fprintf(stderr, "dq_emu: block too large\n"); abort(); }
    std::vector<unsigned char> smem(smem_bytes + 256);
    cta() = c;
    c->body = &body;
    c->bdim = block; c->gdim = grid;
    c->smem = (unsigned char*)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        c->bid = dim3(bx, by, bz);
        c->nthreads = nthreads; c->ndone = 0; c->bar_count = 0; c->bar_gen = 0; c->progress = 0;
        memset(c->nbar_count, 0, sizeof(c->nbar_count)); memset(c->nbar_gen, 0, sizeof(c->nbar_gen));
        memset(c->warps, 0, sizeof(c->warps));
        memset(c->smem, 0xA5, smem_bytes);           // shared memory starts undefined on the GPU: poison it
        for (int i = 0; i < nthreads; ++i) {
            Fiber& f = c->th[i];
            if (!f.stack) f.stack = (char*)malloc(kStackBytes);
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStackBytes;
            f.ctx.uc_link = &c->sched;
            makecontext(&f.ctx, (void (*)())entry, 0);
        }
        int idle_passes = 0;
        const int nwarps = (nthreads + 31) / 32;
        while (c->ndone < nthreads) {
            const long before = c->progress;
            // warps of a CTA run in no particular order on the GPU: DQ_EMU_SCHED = 0 ascending (default), 1 descending,
            // >= 2 a fresh pseudo-random warp order every pass (the value seeds it)
            int order[kMaxThreads / 32];
            for (int w = 0; w < nwarps; ++w) order[w] = sched_mode() == 1 ? nwarps - 1 - w : w;
            if (sched_mode() >= 2)
                for (int w = nwarps - 1; w > 0; --w) { const int j = (int)(sched_rand() % (unsigned)(w + 1)); const int t = order[w]; order[w] = order[j]; order[j] = t; }
            for (int wi = 0; wi < nwarps; ++wi)
                for (int i = order[wi] * 32; i < nthreads && i < order[wi] * 32 + 32; ++i) {
                    if (c->th[i].done) continue;
                    c->cur = i;
                    swapcontext(&c->sched, &c->th[i].ctx);
                }
            idle_passes = (c->progress == before) ? idle_passes + 1 : 0;
            if (idle_passes > idle_limit()) {
                fprintf(stderr, "dq_emu: deadlock in block (%u,%u,%u): %d of %d threads finished, the rest wait at a barrier or warp "
                                "primitive that can never complete\n", bx, by, bz, c->ndone, nthreads);
                abort();
            }
        }
    }
    cta() = nullptr;
}

inline dim3 thread_idx() {
    const Cta& c = *cta();
    const unsigned t = (unsigned)c.cur;
    return dim3(t % c.bdim.x, (t / c.bdim.x) % c.bdim.y, t / (c.bdim.x * c.bdim.y));
}

}  // namespace dq_emu

#define threadIdx (dq_emu::thread_idx())
#define blockIdx (dq_emu::cta()->bid)
#define blockDim (dq_emu::cta()->bdim)
#define gridDim (dq_emu::cta()->gdim)
#define DQ_EMU_DYNAMIC_SMEM (dq_emu::cta()->smem)

// ---- warp / block primitives ----------------------------------------------------------------------------------
static inline void __syncthreads() { dq_emu::block_barrier(); }
static inline void __threadfence_block() {}          // fibers of one OS thread: program order is memory order
static inline void __nanosleep(unsigned) { dq_emu::yield(); }
static inline long long clock64() { static long long t = 0; return t += 1000; }          // 4e6 polls of a spin loop trip the kernel's watchdog
static inline void __trap() { fprintf(stderr, "dq_emu: __trap() (a spin-wait watchdog of the kernel fired)\n"); abort(); }
static inline void __syncwarp(uint32_t mask = 0xffffffffu) { dq_emu::warp_collect(mask, 0); }
static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    const uint64_t* v = dq_emu::warp_collect(mask, pred ? 1 : 0);
    const dq_emu::Cta& c = *dq_emu::cta();
    mask &= dq_emu::lanes_of_warp(c, c.cur >> 5);
    uint32_t b = 0;
    for (int i = 0; i < 32; ++i) if (((mask >> i) & 1u) && v[i]) b |= 1u << i;
    return b;
}
static inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v) {
    const uint64_t* a = dq_emu::warp_collect(mask, v);
    const dq_emu::Cta& c = *dq_emu::cta();
    mask &= dq_emu::lanes_of_warp(c, c.cur >> 5);
    uint32_t m = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) m = std::max(m, (uint32_t)a[i]);
    return m;
}
static inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(uint32_t mask, int pred) { return __ballot_sync(mask, !pred) == 0; }

template <typename T> static inline uint64_t dq_emu_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T dq_emu_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
static inline int dq_emu_lane() { return dq_emu::cta()->cur & 31; }

template <typename T> static inline T __shfl_sync(uint32_t mask, T v, int src, int width = 32) {
    const uint64_t* a = dq_emu::warp_collect(mask, dq_emu_bits(v));
    const int lane = dq_emu_lane(), base = lane & ~(width - 1);
    return dq_emu_unbits<T>(a[base + (src & (width - 1))]);
}
template <typename T> static inline T __shfl_up_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
    const uint64_t* a = dq_emu::warp_collect(mask, dq_emu_bits(v));
    const int lane = dq_emu_lane(), base = lane & ~(width - 1), src = lane - (int)delta;
    return dq_emu_unbits<T>(a[src < base ? lane : src]);
}
template <typename T> static inline T __shfl_down_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
    const uint64_t* a = dq_emu::warp_collect(mask, dq_emu_bits(v));
    const int lane = dq_emu_lane(), base = lane & ~(width - 1), src = lane + (int)delta;
    return dq_emu_unbits<T>(a[src >= base + width ? lane : src]);
}
template <typename T> static inline T __shfl_xor_sync(uint32_t mask, T v, int lm, int width = 32) {
    const uint64_t* a = dq_emu::warp_collect(mask, dq_emu_bits(v));
    const int lane = dq_emu_lane(), base = lane & ~(width - 1), src = lane ^ lm;
    return dq_emu_unbits<T>(a[src >= base + width ? lane : src]);
}

// ---- scalar intrinsics ----------------------------------------------------------------------------------------
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline uint32_t __brev(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31)); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)((((((uint64_t)hi) << 32) | lo) << (sh & 31)) >> 32); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t ab = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(ab >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, const T v) { *p = v; }      // cache hint only
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
// explicitly rounded fp32 operations (the build uses -ffp-contract=off, so plain operators round once as well)
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }

template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicXor(T* p, T v) { T o = *p; *p = o ^ v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ---- the slice of the runtime API the host side of dq_env.cu uses: "device memory" is host memory ------------
typedef int cudaError_t;
typedef struct dq_emu_stream* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "out of memory (emulated)"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
// every buffer of the emulation is host memory the "device" can address: callers' buffers count as pinned
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) { a->type = cudaMemoryTypeHost; a->device = 0; a->devicePointer = const_cast<void*>(p); a->hostPointer = const_cast<void*>(p); return cudaSuccess; }
static inline cudaError_t cudaHostGetDevicePointer(void** dev, void* host, unsigned) { *dev = host; return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
typedef struct dq_emu_event* cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }      // launches run to completion in order
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; ++r) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}
