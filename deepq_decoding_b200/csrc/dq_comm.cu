// Data-parallel training exchange: the gradient all-reduce FUSED with the Adam update, over NVLink peer memory.
//
// The reference trains one agent per process and has no exchange step.  Here the lattices of one run are sharded over
// the GPUs of a box (one process per GPU) and each update averages the flat gradient (193 283 fp32 = 773 KB for the
// d=5 DP network) over the ranks before an identical Adam step on every rank.  Instead of an NCCL
// all-reduce followed by an optimizer kernel, ONE kernel per rank does both:
//
//   * every rank owns one cudaMalloc'd "exchange region" [flags | grads(even) | grads(odd)], mapped into every other
//     rank's address space through CUDA IPC (NVLink P2P loads);
//   * backward writes the local gradient straight into the region (buffer = update parity);
//   * the kernel signals "my gradient of update e is complete" into every peer's flag slot (st.release.sys), waits for
//     the peers' signals on its own flags (ld.acquire.sys), then each thread loads its float4 of every rank's
//     gradient IN RANK ORDER (so the sum is bit-identical on all ranks), scales by 1/world and applies the Keras-2
//     Adam update to the local parameters and moments.
//
// Double buffering by update parity makes the single barrier sufficient: a rank writes buffer (e+1)&1 only after its
// update-e kernel returned, and that kernel waited for every peer's "ready e", which a peer sends only after its
// update-(e-1) kernel (the last reader of that buffer) has finished in stream order.
// A peer that never arrives cannot hang the GPU: the wait gives up after kSpinLimit clocks, raises a sticky
// status flag that the host reads with dq_comm_status, and the update is skipped (no Adam step on stale gradients).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/dq_decoding.h"
#include "dq_adam.cuh"

namespace dq {
extern thread_local std::string g_err_q;
void count_launch();
}  // namespace dq

namespace {

constexpr int kMaxWorld = 16;
constexpr int kFlagBytes = 256;                          // kMaxWorld u32 epochs, padded so the gradients stay 256 B aligned
constexpr long long kSpinLimit = 40000000000ll;          // ~20 s of SM clocks

struct Peers {
    char* base[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {      // L1-bypassing 128-bit load (peer data changes every update)
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}


__global__ void __launch_bounds__(256) allreduce_adam_kernel(Peers peers, int rank, int world, uint32_t epoch, long long grad_off_bytes,
                                                             float* __restrict__ params, float* __restrict__ am, float* __restrict__ av,
                                                             long long n, float lr_t, float b1, float b2, float eps, float inv_world,
                                                             int* __restrict__ status) {
    // ---- 1. "my gradient of this update is complete" -> every rank; wait for theirs -------------------------------
    if (blockIdx.x == 0 && threadIdx.x < world)
        st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x]) + rank, epoch);
    __shared__ int s_timed_out;
    if (threadIdx.x == 0) s_timed_out = 0;
    __syncthreads();
    if (threadIdx.x < world) {
        const uint32_t* f = reinterpret_cast<const uint32_t*>(peers.base[rank]) + threadIdx.x;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
            if (clock64() - t0 > kSpinLimit) { atomicExch(status, 1); s_timed_out = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (s_timed_out) return;          // a peer never arrived: its gradient is stale or incomplete, so no update is applied (the host sees status)
    // ---- 2. sum in rank order, mean, Adam ------------------------------------------------------------------------------
    const long long n4 = (n + 3) >> 2;                  // the exchange buffers are padded to a multiple of 4 floats (zeros)
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float4 g = ld_peer_f4(reinterpret_cast<const float*>(peers.base[r] + grad_off_bytes) + 4 * q);
            s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
        }
        const float g4[4] = {__fmul_rn(s.x, inv_world), __fmul_rn(s.y, inv_world), __fmul_rn(s.z, inv_world), __fmul_rn(s.w, inv_world)};
        const long long i0 = 4 * q;
        if (i0 + 3 < n) {
            float4 p = *reinterpret_cast<float4*>(params + i0), m = *reinterpret_cast<float4*>(am + i0), v = *reinterpret_cast<float4*>(av + i0);
            dq::adam_update(p.x, m.x, v.x, g4[0], lr_t, b1, b2, eps);
            dq::adam_update(p.y, m.y, v.y, g4[1], lr_t, b1, b2, eps);
            dq::adam_update(p.z, m.z, v.z, g4[2], lr_t, b1, b2, eps);
            dq::adam_update(p.w, m.w, v.w, g4[3], lr_t, b1, b2, eps);
            *reinterpret_cast<float4*>(params + i0) = p; *reinterpret_cast<float4*>(am + i0) = m; *reinterpret_cast<float4*>(av + i0) = v;
        } else {
            for (int k = 0; k < 4 && i0 + k < n; ++k) dq::adam_update(params[i0 + k], am[i0 + k], av[i0 + k], g4[k], lr_t, b1, b2, eps);
        }
    }
}

int cfail(int code, const std::string& msg) { dq::g_err_q = msg; return code; }
#define CCUDA(expr)                                                                            \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return cfail(DQ_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

}  // namespace

struct dq_comm {
    int rank, world, device;
    int64_t n, n_pad;
    char* region;                    // this rank's exchange region (cudaMalloc)
    size_t region_bytes;
    Peers peers;                     // every rank's region as mapped here (peers.base[rank] == region)
    bool connected;
    uint32_t epoch;                  // updates completed
    int* status;                     // device flag: 1 = a wait timed out
    int sm_count;
};

extern "C" int dq_comm_create(dq_comm** out, int rank, int world, int64_t n_floats, int device) {
    if (!out) return cfail(DQ_EINVAL, "out is NULL");
    *out = nullptr;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return cfail(DQ_EINVAL, "need 0 <= rank < world <= 16");
    if (n_floats < 1) return cfail(DQ_EINVAL, "n_floats must be positive");
    int ndev = 0;
    CCUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return cfail(DQ_EINVAL, "no such CUDA device");
    CCUDA(cudaSetDevice(device));
    dq_comm* c = new dq_comm();
    memset(c, 0, sizeof(*c));
    c->rank = rank; c->world = world; c->device = device; c->n = n_floats; c->n_pad = (n_floats + 63) / 64 * 64;
    c->region_bytes = kFlagBytes + 2 * (size_t)c->n_pad * sizeof(float);
    cudaError_t e = cudaMalloc(&c->region, c->region_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->region, 0, c->region_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->status, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(c->status, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(c->region); cudaFree(c->status); delete c; return cfail(DQ_ECUDA, std::string("exchange region: ") + cudaGetErrorString(e)); }
    c->peers.base[rank] = c->region;
    c->connected = (world == 1);
    *out = c;
    return DQ_OK;
}

extern "C" int dq_comm_destroy(dq_comm* c) {
    if (!c) return DQ_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peers.base[r]) cudaIpcCloseMemHandle(c->peers.base[r]);
    cudaFree(c->region);
    cudaFree(c->status);
    delete c;
    return DQ_OK;
}

extern "C" int dq_comm_handle(dq_comm* c, void* handle64) {
    if (!c || !handle64) return cfail(DQ_EINVAL, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == DQ_COMM_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CCUDA(cudaSetDevice(c->device));
    CCUDA(cudaIpcGetMemHandle(&h, c->region));
    memcpy(handle64, &h, sizeof(h));
    return DQ_OK;
}

extern "C" int dq_comm_connect(dq_comm* c, const void* handles) {
    if (!c || !handles) return cfail(DQ_EINVAL, "bad argument");
    if (c->connected) return cfail(DQ_ESTATE, "already connected");
    CCUDA(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        CCUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peers.base[r] = static_cast<char*>(p);
    }
    c->connected = true;
    return DQ_OK;
}

extern "C" int dq_comm_next_grads(dq_comm* c, float** grads) {
    if (!c || !grads) return cfail(DQ_EINVAL, "bad argument");
    *grads = reinterpret_cast<float*>(c->region + kFlagBytes) + (size_t)((c->epoch + 1) & 1) * c->n_pad;
    return DQ_OK;
}

extern "C" int dq_comm_allreduce_adam(dq_comm* c, float* params, float* m, float* v, float lr, float beta1, float beta2, float eps,
                                      int64_t t, dq_stream stream) {
    if (!c || !params || !m || !v || t < 1) return cfail(DQ_EINVAL, "bad argument");
    if (!c->connected) return cfail(DQ_ESTATE, "dq_comm_connect has not been called");
    if (((uintptr_t)params | (uintptr_t)m | (uintptr_t)v) & 15) return cfail(DQ_EINVAL, "params / m / v must be 16-byte aligned");
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
    const uint32_t epoch = c->epoch + 1;
    const long long off = kFlagBytes + (long long)(epoch & 1) * c->n_pad * (long long)sizeof(float);
    const long long n4 = (c->n + 3) / 4;
    long long grid = (n4 + 255) / 256;
    if (grid > 2LL * c->sm_count) grid = 2LL * c->sm_count;
    allreduce_adam_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(c->peers, c->rank, c->world, epoch, off, params, m, v, c->n,
                                                                           (float)lr_t, beta1, beta2, eps, 1.0f / (float)c->world, c->status);
    dq::count_launch();
    CCUDA(cudaGetLastError());
    c->epoch = epoch;
    return DQ_OK;
}

extern "C" int dq_comm_status(dq_comm* c, int* timed_out) {
    if (!c || !timed_out) return cfail(DQ_EINVAL, "bad argument");
    CCUDA(cudaSetDevice(c->device));
    CCUDA(cudaMemcpy(timed_out, c->status, sizeof(int), cudaMemcpyDeviceToHost));
    return DQ_OK;
}
