// dq_qnet.cu -- Q-network forward / backward, Keras-Adam, DQN targets + loss, policies, replay ring
// (sm_100a, fp32 SIMT path; the bf16 tcgen05 path for the dense contractions lives in dq_gemm_tc.cu).
//
// Replaces, for N lattices / a batch of B transitions at a time:
//   build_convolutional_nn + the keras-rl dueling head ... example_notebooks/Function_Library.py:338-377
//   DQNAgent.forward / backward (double DQN, hard target copy, Adam) ... keras-rl 0.4.x semantics,
//       call sites cluster_scripts/d5_dp/0.001/Single_Point_Training_Script.py:109-152
//   EpsGreedyQPolicy / GreedyQPolicy with legal-action masking ... same script :110-115, 166-167
//   SequentialMemory(limit, window_length=1) ... same script :109
//
// Data layout.  Observations never exist as bytes on this path: a network input is the PACKED
// observation, one (2d+1)^2-cell bitmap per layer in PW uint64 words, stored row-major as
// [layer*PW + word][sample] -- exactly rows ROW_BM.. of the environment's state matrix, so acting reads
// the env state in place.  Activations are fp32, channels-last ([sample][position][channel]); every
// layer after the first is then a GEMM whose A rows are gathered patches (implicit im2col), and the
// first layer, whose input is binary, is a sparse sum of weight rows over the set taps.
// Parameters / gradients / Adam moments are caller-owned flat fp32 buffers (layout: dq_qnet_param_layout).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/dq_decoding.h"
#include "dq_lattice.cuh"
#include "dq_ptx.cuh"
#include "dq_adam.cuh"

namespace dq {

extern thread_local std::string g_err_q;
void count_launch();

// ---- Programmatic dependent launch for the chains of small kernels (a bf16 forward is 5 launches, a bf16 update 45): a kernel launched
// through DQ_LAUNCH_PDL may be SCHEDULED while its predecessor in the stream is still running; it must call pdl_wait() before its first
// access to global memory (nothing before that point may read or write anything another kernel touches), and calls
// pdl_launch_dependents() once it holds every resource a successor could compete for (tensor memory above all: a successor that took
// the last TMEM columns and then waited for this grid would starve this grid's own CTAs).  In a launch without the attribute both are
// no-ops, so the same kernels serve the plain launches of the fp32 path.  DQ_QNET_PDL=0 launches everything plainly.
__device__ __forceinline__ void pdl_launch_dependents() {
#ifndef DQ_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#ifndef DQ_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
#ifdef DQ_EMU
#define DQ_LAUNCH_PDL(kernel, grid, block, smem, st, ...) kernel<<<grid, block, smem, st>>>(__VA_ARGS__)
#else
inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("DQ_QNET_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define DQ_LAUNCH_PDL(kernel, grid, block, smem, st, ...) dq::launch_pdl(kernel, dim3(grid), dim3(block), smem, st, __VA_ARGS__)
#endif

constexpr int kMaxConv = 4, kMaxDense = 4;

struct ConvL { int cin, ih, oh, ksz, stride, filters, K, P; };     // square maps; K = ksz*ksz*cin; P = oh*oh
struct QCfg {
    int C, H, PW, rows;                 // input layers, side, words per layer, packed rows = C*PW
    int n_conv; ConvL conv[kMaxConv];
    int n_fc;   int fc_in[kMaxDense + 2], fc_out[kMaxDense + 2]; float drop[kMaxDense + 2];
    int A, dueling, n_hidden;
    long long w_off[kMaxConv + kMaxDense + 2], b_off[kMaxConv + kMaxDense + 2];   // offsets in the flat buffer
    long long n_params;
};

// ------------------------------------------------------------------------------------------------ layer 1
// conv on binary input: out[b][pos][co] = relu(bias[co] + sum over set taps k of W[k][co]),
// k = (ky*ksz + kx)*C + ci  (Keras HWIO flattened).  One warp per (sample, position); lane = output channel(s).
__device__ __forceinline__ u32 layer_taps(const u64* __restrict__ packed, long long stride, long long b, int ci, int PW,
                                          int H, int y0, int x0, int ksz) {
    // ksz x ksz window of layer ci at (y0, x0) as a bit mask, tap t = ky*ksz + kx
    u32 taps = 0;
    for (int ky = 0; ky < ksz; ++ky) {
        const int bit = (y0 + ky) * H + x0, w = bit >> 6, s = bit & 63;
        u64 v = packed[(long long)(ci * PW + w) * stride + b] >> s;
        if (s + ksz > 64 && w + 1 < PW) v |= packed[(long long)(ci * PW + w + 1) * stride + b] << (64 - s);
        taps |= (u32)(v & ((1u << ksz) - 1)) << (ky * ksz);
    }
    return taps;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(256)
conv1_bits_kernel(const u64* __restrict__ packed, long long stride, long long batch, ConvL L, int C, int PW, int H,
                  const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ out,
                  const float* __restrict__ dY, float* __restrict__ dW) {
    extern __shared__ float sW[];                      // forward: W [K][F]; backward: dW accumulator [K][F]
    const int F = L.filters, K = L.K;
    if (BACKWARD) {
        for (int i = threadIdx.x; i < K * F; i += blockDim.x) sW[i] = 0.f;
    } else {
        const int n = K * F, bd = blockDim.x;           // 8 loads in flight per thread, then the shared stores
        for (int i0 = 0; i0 < n; i0 += 8 * bd) {
            float w8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w8[j] = __ldg(W + min(i0 + j * bd + (int)threadIdx.x, n - 1));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = i0 + j * bd + (int)threadIdx.x;
                if (i < n) sW[i] = w8[j];
            }
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long total = batch * L.P;
    for (long long job = (long long)blockIdx.x * nwarp + warp; job < total; job += (long long)gridDim.x * nwarp) {
        const long long b = job / L.P;
        const int pos = (int)(job - b * L.P), oy = pos / L.oh, ox = pos - oy * L.oh;
        u32 taps = 0;
        if (lane < C) taps = layer_taps(packed, stride, b, lane, PW, H, oy * L.stride, ox * L.stride, L.ksz);
        if (!BACKWARD) {
            float acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = (lane + 32 * j < F) ? bias[lane + 32 * j] : 0.f;
            for (int ci = 0; ci < C; ++ci) {
                u32 m = __shfl_sync(0xffffffffu, taps, ci);
                while (m) {
                    const int t = __ffs(m) - 1; m &= m - 1;
                    const float* wr = sW + (t * C + ci) * F;
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (lane + 32 * j < F) acc[j] += wr[lane + 32 * j];
                }
            }
            float* o = out + job * F;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (lane + 32 * j < F) o[lane + 32 * j] = fmaxf(acc[j], 0.f);
        } else {
            float g[4];
            const float* gy = dY + job * F;
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = (lane + 32 * j < F) ? gy[lane + 32 * j] : 0.f;
            for (int ci = 0; ci < C; ++ci) {
                u32 m = __shfl_sync(0xffffffffu, taps, ci);
                while (m) {
                    const int t = __ffs(m) - 1; m &= m - 1;
                    float* wr = sW + (t * C + ci) * F;
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (lane + 32 * j < F && g[j] != 0.f) atomicAdd(wr + lane + 32 * j, g[j]);
                }
            }
        }
    }
    if (BACKWARD) {
        __syncthreads();
        for (int i = threadIdx.x; i < K * F; i += blockDim.x) if (sW[i] != 0.f) atomicAdd(dW + i, sW[i]);
    }
}

// ------------------------------------------------------------------------------------------------ patch GEMMs
// A "patch matrix" A[m][k]: m = (sample, oy, ox), k = (ky, kx, c) over a channels-last input [B][ih][ih][cin].
// Dense layers are the degenerate case ih = oh = ksz = 1, cin = K.
struct Patch { int P, oh, ih, cin, ksz, stride; };
__device__ __forceinline__ long long patch_row(const Patch& g, long long m) {
    const long long b = m / g.P;
    const int pos = (int)(m - b * g.P), oy = pos / g.oh, ox = pos - oy * g.oh;
    return ((b * g.ih + oy * g.stride) * g.ih + ox * g.stride) * (long long)g.cin;
}
__device__ __forceinline__ int patch_col(const Patch& g, int k) {
    const int seg = k / g.cin, c = k - seg * g.cin, ky = seg / g.ksz, kx = seg - ky * g.ksz;
    return (ky * g.ih + kx) * g.cin + c;
}

constexpr int TB = 64, TK = 16;        // 64x64 output tile, 16-deep steps, 256 threads, 4x4 per thread

// Y[m][n] = act(sum_k A[m][k] W[k][n] + bias[n]);  act: 0 linear, 1 ReLU.  64 x 64 outputs per CTA, 256 threads, 4x4 per thread.
// KS = depth of one k step: 16 for big grids (many CTAs hide each other's load latency), 64 when the grid is small and
// every (load -> sync -> FMA -> sync) step is an exposed memory latency.
template <int KS>
__global__ void __launch_bounds__(256, 2)
gemm_fwd_kernel(const float* __restrict__ X, Patch g, const float* __restrict__ W, const float* __restrict__ bias,
                float* __restrict__ Y, long long M, int N, int K, int act) {
    __shared__ float As[KS][TB + 4], Bs[KS][TB + 4];
    __shared__ long long rowoff[TB];
    __shared__ int coloff[2][KS];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * TB;
    const int n0 = blockIdx.y * TB;
    constexpr int E = TB * KS / 256;                    // elements of each tile per thread
    if (tid < TB) rowoff[tid] = (m0 + tid < M) ? patch_row(g, m0 + tid) : -1;
    if (tid < KS) coloff[0][tid] = (tid < K) ? patch_col(g, tid) : -1;       // one division per k per CTA, not per element
    __syncthreads();
    float acc[4][4] = {};
    float va[E], vb[E];
    // All global loads of a k step are issued together (unconditional loads from a safe address), one step AHEAD of the
    // math: the registers are stored to shared memory at the top of the next iteration, so the L2 round trip of step s+1
    // hides behind the FMAs of step s.
    auto load_tiles = [&](int k0, const int* co_tab) {
#pragma unroll
        for (int i = 0; i < E; ++i) {                   // A tile: 64 rows x KS k
            const int e = tid + i * 256, r = e / KS, kk = e % KS;
            const long long ro = rowoff[r];
            const int co = co_tab[kk];
            const bool ok = co >= 0 && ro >= 0;
            const float v = __ldg(X + (ok ? ro + co : 0));
            va[i] = ok ? v : 0.f;
        }
#pragma unroll
        for (int i = 0; i < E; ++i) {                   // W tile: KS k x 64 n
            const int e = tid + i * 256, kk = e >> 6, c = e & 63, k = k0 + kk, n = n0 + c;
            const bool ok = k < K && n < N;
            const float v = __ldg(W + (ok ? (long long)k * N + n : 0));
            vb[i] = ok ? v : 0.f;
        }
    };
    load_tiles(0, coloff[0]);
    int it = 0;
    for (int k0 = 0; k0 < K; k0 += KS, ++it) {
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const int e = tid + i * 256;
            As[e % KS][e / KS] = va[i];
            Bs[e >> 6][e & 63] = vb[i];
        }
        const bool more = k0 + KS < K;
        if (more && tid < KS) coloff[(it + 1) & 1][tid] = (k0 + KS + tid < K) ? patch_col(g, k0 + KS + tid) : -1;
        __syncthreads();
        if (more) load_tiles(k0 + KS, coloff[(it + 1) & 1]);
#pragma unroll 16
        for (int kk = 0; kk < KS; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (act == 1) v = fmaxf(v, 0.f);
            Y[m * N + n] = v;
        }
    }
}

// dW[k][n] += sum_m A[m][k] dY[m][n]   over this CTA's slice of m (grid.z), atomically
__global__ void __launch_bounds__(256)
gemm_dw_kernel(const float* __restrict__ X, Patch g, const float* __restrict__ dY, float* __restrict__ dW,
               long long M, int N, int K, long long m_chunk) {
    pdl_launch_dependents(); pdl_wait();
    __shared__ float As[TK][TB + 4], Bs[TK][TB + 4];
    __shared__ int coloff[TB];
    __shared__ long long rowoff[TK];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int k0 = blockIdx.x * TB, n0 = blockIdx.y * TB;
    const long long mb = (long long)blockIdx.z * m_chunk, me = min(M, mb + m_chunk);
    if (tid < TB) coloff[tid] = (k0 + tid < K) ? patch_col(g, k0 + tid) : -1;
    float acc[4][4] = {};
    for (long long ms = mb; ms < me; ms += TK) {
        if (tid < TK) rowoff[tid] = (ms + tid < me) ? patch_row(g, ms + tid) : -1;
        __syncthreads();
        float va[4], vb[4];                            // loads first, shared stores after (see gemm_fwd_kernel)
#pragma unroll
        for (int i = 0; i < 4; ++i) {                  // A^T tile: 16 m x 64 k
            const int e = tid + i * 256, mm = e >> 6, c = e & 63;
            const long long ro = rowoff[mm];
            const int co = coloff[c];
            const bool ok = ro >= 0 && co >= 0;
            const float v = __ldg(X + (ok ? ro + co : 0));
            va[i] = ok ? v : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {                  // dY tile: 16 m x 64 n
            const int e = tid + i * 256, mm = e >> 6, c = e & 63;
            const long long m = ms + mm;
            const bool ok = m < me && n0 + c < N;
            const float v = __ldg(dY + (ok ? m * N + n0 + c : 0));
            vb[i] = ok ? v : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            As[e >> 6][e & 63] = va[i];
            Bs[e >> 6][e & 63] = vb[i];
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < TK; ++mm) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[mm][ty * 4 + i]; b[i] = Bs[mm][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N && acc[i][j] != 0.f) atomicAdd(dW + (long long)k * N + n, acc[i][j]);
        }
    }
}

// dX[patch(m,k)] (+)= sum_n dY[m][n] W[k][n];  overlapping patches (conv) accumulate atomically into a zeroed dX
template <int NS>
__global__ void __launch_bounds__(256)
gemm_dx_kernel(const float* __restrict__ dY, const float* __restrict__ W, float* __restrict__ dX, Patch g,
               long long M, int N, int K, int overlap) {
    pdl_launch_dependents(); pdl_wait();
    __shared__ float As[NS][TB + 4], Bs[NS][TB + 4];
    __shared__ long long rowoff[TB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * TB;
    const int k0 = blockIdx.y * TB;
    constexpr int E = TB * NS / 256;
    if (tid < TB) rowoff[tid] = (m0 + tid < M) ? patch_row(g, m0 + tid) : -1;
    __syncthreads();
    float acc[4][4] = {};
    float va[E], vb[E];
    auto load_tiles = [&](int ns) {                     // one reduction step ahead of the math (see gemm_fwd_kernel)
#pragma unroll
        for (int i = 0; i < E; ++i) {                   // dY tile: 64 m x NS n
            const int e = tid + i * 256, r = e / NS, nn = e % NS;
            const bool ok = m0 + r < M && ns + nn < N;
            const float v = __ldg(dY + (ok ? (m0 + r) * N + ns + nn : 0));
            va[i] = ok ? v : 0.f;
        }
#pragma unroll
        for (int i = 0; i < E; ++i) {                   // W^T tile: NS n x 64 k
            const int e = tid + i * 256, c = e / NS, nn = e % NS;
            const bool ok = k0 + c < K && ns + nn < N;
            const float v = __ldg(W + (ok ? (long long)(k0 + c) * N + ns + nn : 0));
            vb[i] = ok ? v : 0.f;
        }
    };
    load_tiles(0);
    for (int ns = 0; ns < N; ns += NS) {
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const int e = tid + i * 256;
            As[e % NS][e / NS] = va[i];
            Bs[e % NS][e / NS] = vb[i];
        }
        __syncthreads();
        if (ns + NS < N) load_tiles(ns + NS);
#pragma unroll 16
        for (int nn = 0; nn < NS; ++nn) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[nn][ty * 4 + i]; b[i] = Bs[nn][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        if (rowoff[r] < 0) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k >= K) continue;
            float* dst = dX + rowoff[r] + patch_col(g, k);
            if (overlap) atomicAdd(dst, acc[i][j]); else *dst = acc[i][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------ elementwise
// dY *= (Y > 0) [* mask];  Y is the layer's post-activation output
__global__ void relu_bwd_kernel(float* __restrict__ dY, const float* __restrict__ Y, const float* __restrict__ mask, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float g = (Y[i] > 0.f) ? dY[i] : 0.f;
        if (mask) g *= mask[i];
        dY[i] = g;
    }
}
// training-mode dropout after a ReLU layer: mask = keep ? 1/(1-rate) : 0 (Keras inverted dropout); Y *= mask
__global__ void dropout_kernel(float* __restrict__ Y, float* __restrict__ mask, long long n, float rate, u32 k0, u32 k1, u32 tag) {
    const u32 thr = (u32)fminf(rate * 4294967296.f, 4294967295.f);
    const float scale = 1.f / (1.f - rate);
    for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 * 4 < n; i4 += (long long)gridDim.x * blockDim.x) {
        const Philox4 u = philox4x32_10((u32)i4, (u32)(i4 >> 32), tag, 3u, k0, k1);      // domain 3: dropout
        const u32 uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i4 * 4 + j;
            if (i < n) { const float mk = (uu[j] >= thr) ? scale : 0.f; mask[i] = mk; Y[i] *= mk; }
        }
    }
}
__global__ void colsum_kernel(const float* __restrict__ dY, float* __restrict__ db, long long M, int N) {
    pdl_launch_dependents(); pdl_wait();
    // grid.x covers columns in blocks of 32; each block reduces a slice of rows (grid.y) and adds atomically
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    float s = 0.f;
    if (c < N)
        for (long long m = (long long)blockIdx.y * 8 + w; m < M; m += (long long)gridDim.y * 8) s += dY[m * N + c];
    part[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        if (c < N && t != 0.f) atomicAdd(db + c, t);
    }
}
// dueling head (dueling_type='avg'): Q[b][a] = y[b][0] + y[b][1+a] - mean_a y[b][1+a]
__global__ void dueling_fwd_kernel(const float* __restrict__ y, float* __restrict__ q, long long B, int A) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* r = y + b * (A + 1);
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += r[1 + a];
    const float base = r[0] - s / (float)A;
    for (int a = 0; a < A; ++a) q[b * A + a] = base + r[1 + a];
}
__global__ void dueling_bwd_kernel(const float* __restrict__ dq, float* __restrict__ dy, long long B, int A) {
    pdl_launch_dependents(); pdl_wait();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += dq[b * A + a];
    dy[b * (A + 1)] = s;
    for (int a = 0; a < A; ++a) dy[b * (A + 1) + 1 + a] = dq[b * A + a] - s / (float)A;
}
// The whole backward step of the dueling layer Dense(A -> A+1) + 'avg' combine in ONE launch (it is 4 of the 45 launches of a bf16
// update, all on the chain every other layer waits for): per CTA of 32 samples, dY = combine^T(dq) in shared memory, then
//   dX[b][k] = sum_n dY[b][n] W[k][n]      (the gradient at the output of Dense(A), what the tensor-core layers consume)
//   dW[k][n] += sum_b x[b][k] dY[b][n],  db[n] += sum_b dY[b][n]      (fp32 atomics: one partial per CTA)
// x = the layer's input (Dense(A) output, fp32 [B][K]), W [K][N = A+1], K = A.
constexpr int kHeadBwdSamples = 32;
__global__ void __launch_bounds__(128)
head_dueling_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ x, const float* __restrict__ W, long long B, int K, int N, int A,
                        float* __restrict__ dX, float* __restrict__ dW, float* __restrict__ db) {
    pdl_launch_dependents(); pdl_wait();
    extern __shared__ __align__(16) float hb[];          // W [K][N], dY [32][N], x [32][K]
    float* sW = hb;
    float* sY = sW + K * N;
    float* sX = sY + kHeadBwdSamples * N;
    const int tid = threadIdx.x;
    const long long b0 = (long long)blockIdx.x * kHeadBwdSamples;
    const int nrow = (int)min((long long)kHeadBwdSamples, B - b0);
    for (int i = tid; i < K * N; i += 128) sW[i] = W[i];
    for (int i = tid; i < kHeadBwdSamples * K; i += 128) sX[i] = (i / K < nrow) ? x[b0 * K + i] : 0.f;
    if (tid < kHeadBwdSamples) {                            // dueling_bwd_kernel for one sample
        float* y = sY + tid * N;
        if (tid < nrow) {
            const float* g = dq + (b0 + tid) * A;
            float sum = 0.f;
            for (int a = 0; a < A; ++a) sum += g[a];
            y[0] = sum;
            for (int a = 0; a < A; ++a) y[1 + a] = g[a] - sum / (float)A;
        } else {
            for (int n = 0; n < N; ++n) y[n] = 0.f;
        }
    }
    __syncthreads();
    for (int i = tid; i < nrow * K; i += 128) {
        const int r = i / K, k = i - r * K;
        const float* y = sY + r * N;
        const float* w = sW + k * N;
        float acc = 0.f;
        for (int n = 0; n < N; ++n) acc = fmaf(y[n], w[n], acc);
        dX[(b0 + r) * K + k] = acc;
    }
    for (int i = tid; i < K * N; i += 128) {
        const int k = i / N, n = i - k * N;
        float acc = 0.f;
        for (int r = 0; r < kHeadBwdSamples; ++r) acc = fmaf(sX[r * K + k], sY[r * N + n], acc);
        if (acc != 0.f) atomicAdd(dW + i, acc);
    }
    if (tid < N) {
        float acc = 0.f;
        for (int r = 0; r < kHeadBwdSamples; ++r) acc += sY[r * N + tid];
        if (acc != 0.f) atomicAdd(db + tid, acc);
    }
}
// Keras-2 Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; p -= lr_t*m/(sqrt(v)+eps).  g is scaled by gscale first
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                            long long n, float lr_t, float b1, float b2, float eps, float gscale) {
    pdl_launch_dependents(); pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        dq::adam_update(pi, mi, vi, __fmul_rn(g[i], gscale), lr_t, b1, b2, eps);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}
// double-DQN target: y = r + gamma*(1-terminal)*Qt[argmax_a Qo[a]]   (ties -> lowest index)
__global__ void dqn_target_kernel(const float* __restrict__ qo, const float* __restrict__ qt, const float* __restrict__ r,
                                  const uint8_t* __restrict__ term, float gamma, long long B, int A, float* __restrict__ y) {
    pdl_launch_dependents(); pdl_wait();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int best = 0; float bv = qo[b * A];
    for (int a = 1; a < A; ++a) { const float v = qo[b * A + a]; if (v > bv) { bv = v; best = a; } }
    y[b] = r[b] + (term[b] ? 0.f : gamma * qt[b * A + best]);
}
// loss = mean_b 0.5*(y - Q[b][a_b])^2;  dQ = d loss / dQ;  stats += {sum of 0.5*err^2, sum of max_a Q}
__global__ void dqn_loss_grad_kernel(const float* __restrict__ q, const int32_t* __restrict__ act, const float* __restrict__ y,
                                     long long B, int A, float* __restrict__ dq, float* __restrict__ stats) {
    pdl_launch_dependents(); pdl_wait();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f, mq = 0.f;
    if (b < B) {
        const int a = act[b];
        float mx = q[b * A];
        for (int j = 0; j < A; ++j) { dq[b * A + j] = 0.f; mx = fmaxf(mx, q[b * A + j]); }
        const float err = q[b * A + a] - y[b];
        dq[b * A + a] = err / (float)B;
        l = 0.5f * err * err; mq = mx;
    }
    for (int o = 16; o; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); mq += __shfl_xor_sync(0xffffffffu, mq, o); }
    if ((threadIdx.x & 31) == 0 && stats) { atomicAdd(stats, l); atomicAdd(stats + 1, mq); }
}
// eps-greedy over legal actions (EpsGreedyQPolicy: with probability eps uniform over env.legal_actions, else argmax
// over all actions, or over the legal ones when masked_greedy).  eps = 0 + masked_greedy is GreedyQPolicy(masked).
// Draws: Philox(stream id, step, 0, domain 1): word 0 -> the uniform pick (same as dq_policy_random_legal),
// word 1 -> the eps test (u < floor(eps*2^32)).
__global__ void eps_greedy_kernel(const float* __restrict__ q, const u64* __restrict__ legal, int n, int W, int A, u32 env_id_base,
                                  u32 step, u32* __restrict__ ctr, u32 k0, u32 k1, u32 eps_thr, int masked_greedy,
                                  int32_t* __restrict__ actions) {
    pdl_launch_dependents(); pdl_wait();
    __shared__ u32 s_step;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctr) {
        if (threadIdx.x == 0) s_step = *reinterpret_cast<volatile u32*>(ctr);
        __syncthreads();
        step = s_step;
    }
    if (e < n) {
        u64 m[3] = {0, 0, 0};
        int cnt = 0;
        for (int i = 0; i < W; ++i) { m[i] = legal[(size_t)e * W + i]; cnt += popc64(m[i]); }
        const Philox4 u = philox4x32_10(env_id_base + (u32)e, step, 0u, 1u, k0, k1);
        int act = A - 1;
        if (u.y < eps_thr) {
            int pick = (int)mulhi32(u.x, (u32)cnt);
            for (int i = 0; i < W; ++i) {
                const int c = popc64(m[i]);
                if (pick < c) { act = i * 64 + select64(m[i], pick); break; }
                pick -= c;
            }
        } else {
            float bv = -INFINITY; int best = -1;
            for (int a = 0; a < A; ++a) {
                if (masked_greedy && !((m[a >> 6] >> (a & 63)) & 1)) continue;
                const float v = q[(size_t)e * A + a];
                if (best < 0 || v > bv) { bv = v; best = a; }
            }
            act = best < 0 ? A - 1 : best;
        }
        actions[e] = act;
    }
    if (ctr && threadIdx.x == 0 && atomicAdd(ctr + 1, 1u) == gridDim.x - 1) { ctr[1] = 0; atomicAdd(ctr, 1u); }
}
// bytes [B][C][H][H] -> packed [C*PW][stride]
__global__ void pack_obs_kernel(const uint8_t* __restrict__ obs, u64* __restrict__ packed, long long stride, long long B,
                                int C, int PW, int cells) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per (row, sample)
    if (t >= B * C * PW) return;
    const long long b = t % B; const int row = (int)(t / B), layer = row / PW, w = row - layer * PW;
    const uint8_t* src = obs + (b * C + layer) * cells;
    u64 v = 0;
    for (int i = 0; i < 64; ++i) { const int c = w * 64 + i; if (c < cells && src[c]) v |= 1ull << i; }
    packed[(long long)row * stride + b] = v;
}
// replay sample: uniform (slot, lattice) pairs -> packed s / s' batches + action, reward, terminal
__global__ void replay_sample_kernel(const u64* __restrict__ ring_obs, const int32_t* __restrict__ ring_act,
                                     const float* __restrict__ ring_rew, const uint8_t* __restrict__ ring_term,
                                     int rows, long long npad, int n, int cap, int head, int filled, long long batch,
                                     u32 k0, u32 k1, u32 draw, u64* __restrict__ s0, u64* __restrict__ s1,
                                     int32_t* __restrict__ act, float* __restrict__ rew, uint8_t* __restrict__ term,
                                     int32_t* __restrict__ picked) {
    pdl_launch_dependents(); pdl_wait();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    // transitions live in the `filled` most recent completed slots; slot `head` is the one being written next
    const Philox4 u = philox4x32_10((u32)b, (u32)(b >> 32), draw, 2u, k0, k1);                  // domain 2: replay
    const int age = (int)mulhi32(u.x, (u32)filled);                // 0 = most recent completed transition
    const int i = (int)mulhi32(u.y, (u32)n);
    const int t = (head - 1 - age + 2 * cap) % cap, t1 = (t + 1) % cap;
    for (int r = 0; r < rows; ++r) {
        s0[(long long)r * batch + b] = ring_obs[((long long)t * rows + r) * npad + i];
        s1[(long long)r * batch + b] = ring_obs[((long long)t1 * rows + r) * npad + i];
    }
    act[b] = ring_act[(long long)t * n + i];
    rew[b] = ring_rew[(long long)t * n + i];
    term[b] = ring_term[(long long)t * n + i];
    if (picked) { picked[2 * b] = t; picked[2 * b + 1] = i; }
}

}  // namespace dq

// ================================================================================================ host / C ABI
using namespace dq;

static int qfail(int code, const std::string& msg) { dq::g_err_q = msg; return code; }
#define QCUDA(expr)                                                                          \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) return qfail(DQ_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct dq_qnet {
    QCfg c;
    int device;
    long long max_batch;
    // activations (fp32, channels-last) of the last forward; gradient scratch of the same shapes
    float* act_conv[kMaxConv]; float* dact_conv[kMaxConv];
    float* act_fc[kMaxDense + 2]; float* dact_fc[kMaxDense + 2]; float* mask_fc[kMaxDense + 2];
    u64* pack_scratch;
    int last_train;
    void* tc;                       // bf16 buffers of the tensor-core path (dq_qnet_tc), allocated on first use
};
static void tc_free(dq_qnet* h);

static int grid_for(long long n, int block, int cap = 148 * 16) {
    long long g = (n + block - 1) / block;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int dq_qnet_create(dq_qnet** out, int in_channels, int in_side, int n_conv, const int* filters, const int* kernels,
                              const int* strides, int n_dense, const int* units, const float* dropout, int num_actions,
                              int dueling, int64_t max_batch, int device) {
    if (!out) return qfail(DQ_EINVAL, "out is NULL");
    *out = nullptr;
    if (n_conv < 1 || n_conv > kMaxConv || n_dense < 0 || n_dense > kMaxDense) return qfail(DQ_EINVAL, "1..4 conv layers and 0..4 hidden dense layers are supported");
    if (in_channels < 1 || in_channels > 32 || in_side < 3 || in_side > 15) return qfail(DQ_EINVAL, "input must be [<=32][<=15][<=15]");
    if (num_actions < 2 || max_batch < 1) return qfail(DQ_EINVAL, "bad num_actions / max_batch");
    int ndev = 0;
    QCUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return qfail(DQ_EINVAL, "no such CUDA device");
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(device);
    dq_qnet* h = new dq_qnet();
    memset(h, 0, sizeof(*h));
    QCfg& c = h->c;
    c.C = in_channels; c.H = in_side; c.PW = (in_side * in_side + 63) / 64; c.rows = c.C * c.PW;
    c.n_conv = n_conv; c.A = num_actions; c.dueling = dueling ? 1 : 0; c.n_hidden = n_dense;
    int cin = in_channels, side = in_side;
    long long off = 0;
    int t = 0;
    for (int l = 0; l < n_conv; ++l) {
        ConvL& L = c.conv[l];
        L.cin = cin; L.ih = side; L.ksz = kernels[l]; L.stride = strides[l]; L.filters = filters[l];
        if (L.ksz < 1 || L.ksz > 4 || L.stride < 1 || L.ksz > side || L.filters < 1 || L.filters > 128) { delete h; cudaSetDevice(prev); return qfail(DQ_EINVAL, "unsupported conv layer"); }
        L.oh = (side - L.ksz) / L.stride + 1; L.K = L.ksz * L.ksz * cin; L.P = L.oh * L.oh;
        c.w_off[t] = off; off += (long long)L.K * L.filters; c.b_off[t] = off; off += L.filters; ++t;
        cin = L.filters; side = L.oh;
    }
    int fin = cin * side * side;
    c.n_fc = n_dense + 1 + (dueling ? 1 : 0);
    for (int i = 0; i < c.n_fc; ++i) {
        const int fo = i < n_dense ? units[i] : (i == n_dense ? num_actions : num_actions + 1);
        c.fc_in[i] = fin; c.fc_out[i] = fo; c.drop[i] = (i < n_dense && dropout) ? dropout[i] : 0.f;
        c.w_off[t] = off; off += (long long)fin * fo; c.b_off[t] = off; off += fo; ++t;
        fin = fo;
    }
    c.n_params = off;
    h->device = device; h->max_batch = max_batch;
    cudaError_t err = cudaSuccess;
    auto alloc = [&](float** p, long long n) { if (err == cudaSuccess) err = cudaMalloc(p, (size_t)n * sizeof(float)); };
    for (int l = 0; l < n_conv; ++l) {
        const long long n = max_batch * c.conv[l].P * c.conv[l].filters;
        alloc(&h->act_conv[l], n); alloc(&h->dact_conv[l], n);
    }
    for (int i = 0; i < c.n_fc; ++i) {
        const long long n = max_batch * c.fc_out[i];
        alloc(&h->act_fc[i], n); alloc(&h->dact_fc[i], n);
        if (c.drop[i] > 0.f) alloc(&h->mask_fc[i], n);
    }
    if (err == cudaSuccess) err = cudaMalloc(&h->pack_scratch, (size_t)c.rows * max_batch * 8);
    cudaSetDevice(prev);
    if (err != cudaSuccess) { return qfail(DQ_ECUDA, std::string("cudaMalloc(activations): ") + cudaGetErrorString(err)); }
    *out = h;
    return DQ_OK;
}

extern "C" int dq_qnet_destroy(dq_qnet* h) {
    if (!h) return DQ_OK;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(h->device);
    for (int l = 0; l < kMaxConv; ++l) { cudaFree(h->act_conv[l]); cudaFree(h->dact_conv[l]); }
    for (int i = 0; i < kMaxDense + 2; ++i) { cudaFree(h->act_fc[i]); cudaFree(h->dact_fc[i]); cudaFree(h->mask_fc[i]); }
    cudaFree(h->pack_scratch);
    tc_free(h);
    cudaSetDevice(prev);
    delete h;
    return DQ_OK;
}

extern "C" int dq_qnet_info(const dq_qnet* h, int what, int64_t* out) {
    if (!h || !out) return qfail(DQ_EINVAL, "NULL argument");
    switch (what) {
        case DQ_QINFO_NUM_PARAMS: *out = h->c.n_params; break;
        case DQ_QINFO_PACKED_ROWS: *out = h->c.rows; break;
        case DQ_QINFO_NUM_TENSORS: *out = 2 * (h->c.n_conv + h->c.n_fc); break;
        case DQ_QINFO_FLOPS_PER_SAMPLE: {
            long long f = 0;
            for (int l = 0; l < h->c.n_conv; ++l) f += 2ll * h->c.conv[l].K * h->c.conv[l].filters * h->c.conv[l].P;
            for (int i = 0; i < h->c.n_fc; ++i) f += 2ll * h->c.fc_in[i] * h->c.fc_out[i];
            *out = f; break;
        }
        default: return qfail(DQ_EINVAL, "unknown info selector");
    }
    return DQ_OK;
}

// offsets[2*t] = kernel offset, offsets[2*t+1] = bias offset of tensor pair t (conv layers first, then dense);
// shapes[2*t], shapes[2*t+1] = (rows K, cols N) of the kernel.  Conv kernels are Keras HWIO flattened to [K][N];
// the first dense kernel's rows are in (position, channel) order (see qnet.py for the permutation from Keras' C,H,W).
extern "C" int dq_qnet_param_layout(const dq_qnet* h, int64_t* offsets, int64_t* shapes) {
    if (!h || !offsets || !shapes) return qfail(DQ_EINVAL, "NULL argument");
    const QCfg& c = h->c;
    int t = 0;
    for (int l = 0; l < c.n_conv; ++l, ++t) { offsets[2 * t] = c.w_off[t]; offsets[2 * t + 1] = c.b_off[t]; shapes[2 * t] = c.conv[l].K; shapes[2 * t + 1] = c.conv[l].filters; }
    for (int i = 0; i < c.n_fc; ++i, ++t) { offsets[2 * t] = c.w_off[t]; offsets[2 * t + 1] = c.b_off[t]; shapes[2 * t] = c.fc_in[i]; shapes[2 * t + 1] = c.fc_out[i]; }
    return DQ_OK;
}

extern "C" int dq_qnet_pack_obs(dq_qnet* h, const uint8_t* obs, uint64_t* packed, int64_t stride, int64_t batch, dq_stream stream) {
    if (!h || !obs || !packed || batch < 1 || stride < batch) return qfail(DQ_EINVAL, "bad argument");
    const QCfg& c = h->c;
    const long long n = batch * c.rows;
    pack_obs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(obs, (u64*)packed, stride, batch, c.C, c.PW, c.H * c.H);
    count_launch();
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

static void launch_gemm_fwd(const float* X, const Patch& g, const float* W, const float* bias, float* Y, long long M, int N, int K,
                            int act, cudaStream_t st) {
    dim3 grid((unsigned)((M + 63) / 64), (N + 63) / 64);
    if ((long long)grid.x * grid.y >= 2 * 148) gemm_fwd_kernel<16><<<grid, 256, 0, st>>>(X, g, W, bias, Y, M, N, K, act);
    else gemm_fwd_kernel<64><<<grid, 256, 0, st>>>(X, g, W, bias, Y, M, N, K, act);
    count_launch();
}

// rows per CTA of gemm_dw_kernel: enough slices of M that (k tiles x n tiles x slices) fills the GPU about four times over
static long long dw_chunk(long long M, int K, int N) {
    const long long tiles = (long long)((K + TB - 1) / TB) * ((N + TB - 1) / TB);
    long long slices = std::max<long long>(1, std::min<long long>((M + 63) / 64, (592 + tiles - 1) / tiles));
    long long chunk = (M + slices - 1) / slices;
    return (chunk + TK - 1) / TK * TK;
}

static Patch conv_patch(const ConvL& L) { Patch g; g.P = L.P; g.oh = L.oh; g.ih = L.ih; g.cin = L.cin; g.ksz = L.ksz; g.stride = L.stride; return g; }
static Patch dense_patch(int K) { Patch g; g.P = 1; g.oh = 1; g.ih = 1; g.cin = K; g.ksz = 1; g.stride = 1; return g; }

extern "C" int dq_qnet_forward(dq_qnet* h, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                               float* q_out, int train, uint64_t dropout_seed, dq_stream stream) {
    if (!h || !params || !packed || !q_out) return qfail(DQ_EINVAL, "NULL argument");
    if (batch < 1 || batch > h->max_batch) return qfail(DQ_EINVAL, "batch exceeds max_batch of the handle");
    const QCfg& c = h->c;
    cudaStream_t st = (cudaStream_t)stream;
    int t = 0;
    {   // layer 1 on packed bits
        const ConvL& L = c.conv[0];
        const size_t smem = (size_t)L.K * L.filters * sizeof(float);
        if (smem > 48 * 1024) QCUDA(cudaFuncSetAttribute(conv1_bits_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = grid_for(batch * L.P, 8, 148 * 8);
        conv1_bits_kernel<false><<<grid, 256, smem, st>>>((const u64*)packed, stride, batch, L, c.C, c.PW, c.H,
                                                           params + c.w_off[0], params + c.b_off[0], h->act_conv[0], nullptr, nullptr);
        count_launch(); ++t;
    }
    const float* x = h->act_conv[0];
    for (int l = 1; l < c.n_conv; ++l, ++t) {
        const ConvL& L = c.conv[l];
        const long long M = batch * L.P;
        launch_gemm_fwd(x, conv_patch(L), params + c.w_off[t], params + c.b_off[t], h->act_conv[l], M, L.filters, L.K, 1, st);
        x = h->act_conv[l];
    }
    for (int i = 0; i < c.n_fc; ++i, ++t) {
        const int K = c.fc_in[i], N = c.fc_out[i];
        launch_gemm_fwd(x, dense_patch(K), params + c.w_off[t], params + c.b_off[t], h->act_fc[i], batch, N, K, i < c.n_hidden ? 1 : 0, st);
        if (train && c.drop[i] > 0.f) {
            const long long n = batch * N;
            dropout_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(h->act_fc[i], h->mask_fc[i], n, c.drop[i], (u32)dropout_seed, (u32)(dropout_seed >> 32), (u32)i);
            count_launch();
        }
        x = h->act_fc[i];
    }
    if (c.dueling) {
        dueling_fwd_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, st>>>(x, q_out, batch, c.A);
        count_launch();
    } else {
        QCUDA(cudaMemcpyAsync(q_out, x, (size_t)batch * c.A * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    h->last_train = train;
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

// Gradients of sum_b sum_a dq[b][a]*Q[b][a] w.r.t. every parameter, for the batch of the LAST dq_qnet_forward call
// (same packed input, train flag honoured: dropout masks are reused).  grads (flat, n_params) is overwritten.
extern "C" int dq_qnet_backward(dq_qnet* h, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                                const float* dq, float* grads, dq_stream stream) {
    if (!h || !params || !packed || !dq || !grads) return qfail(DQ_EINVAL, "NULL argument");
    if (batch < 1 || batch > h->max_batch) return qfail(DQ_EINVAL, "batch exceeds max_batch of the handle");
    const QCfg& c = h->c;
    cudaStream_t st = (cudaStream_t)stream;
    QCUDA(cudaMemsetAsync(grads, 0, (size_t)c.n_params * sizeof(float), st));
    const int nt = c.n_conv + c.n_fc;
    // head
    float* dy = h->dact_fc[c.n_fc - 1];
    if (c.dueling) { dueling_bwd_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, st>>>(dq, dy, batch, c.A); count_launch(); }
    else QCUDA(cudaMemcpyAsync(dy, dq, (size_t)batch * c.A * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // dense stack, last to first
    for (int i = c.n_fc - 1; i >= 0; --i) {
        const int t = c.n_conv + i, K = c.fc_in[i], N = c.fc_out[i];
        float* dY = h->dact_fc[i];
        if (i < c.n_hidden) {       // ReLU (+ dropout) layer: mask the incoming gradient
            const long long n = batch * N;
            relu_bwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(dY, h->act_fc[i], (h->last_train && c.drop[i] > 0.f) ? h->mask_fc[i] : nullptr, n);
            count_launch();
        }
        const float* xin = i > 0 ? h->act_fc[i - 1] : h->act_conv[c.n_conv - 1];
        float* dxin = i > 0 ? h->dact_fc[i - 1] : h->dact_conv[c.n_conv - 1];
        const long long chunk = dw_chunk(batch, K, N);
        dim3 gw((K + TB - 1) / TB, (N + TB - 1) / TB, (unsigned)((batch + chunk - 1) / chunk));
        gemm_dw_kernel<<<gw, 256, 0, st>>>(xin, dense_patch(K), dY, grads + c.w_off[t], batch, N, K, chunk);
        colsum_kernel<<<dim3((N + 31) / 32, (unsigned)std::min<long long>(64, (batch + 7) / 8)), 256, 0, st>>>(dY, grads + c.b_off[t], batch, N);
        dim3 gx((unsigned)((batch + TB - 1) / TB), (K + TB - 1) / TB);
        if ((long long)gx.x * gx.y >= 2 * 148 || N <= 16) gemm_dx_kernel<16><<<gx, 256, 0, st>>>(dY, params + c.w_off[t], dxin, dense_patch(K), batch, N, K, 0);
        else gemm_dx_kernel<64><<<gx, 256, 0, st>>>(dY, params + c.w_off[t], dxin, dense_patch(K), batch, N, K, 0);
        count_launch(); count_launch(); count_launch();
    }
    // conv stack, last to second
    for (int l = c.n_conv - 1; l >= 1; --l) {
        const ConvL& L = c.conv[l];
        const long long M = batch * L.P, n = M * L.filters;
        float* dY = h->dact_conv[l];
        relu_bwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(dY, h->act_conv[l], nullptr, n);
        const long long chunk = dw_chunk(M, L.K, L.filters);
        dim3 gw((L.K + TB - 1) / TB, (L.filters + TB - 1) / TB, (unsigned)((M + chunk - 1) / chunk));
        gemm_dw_kernel<<<gw, 256, 0, st>>>(h->act_conv[l - 1], conv_patch(L), dY, grads + c.w_off[l], M, L.filters, L.K, chunk);
        colsum_kernel<<<dim3((L.filters + 31) / 32, (unsigned)std::min<long long>(64, (M + 7) / 8)), 256, 0, st>>>(dY, grads + c.b_off[l], M, L.filters);
        const ConvL& Lp = c.conv[l - 1];
        QCUDA(cudaMemsetAsync(h->dact_conv[l - 1], 0, (size_t)batch * Lp.P * Lp.filters * sizeof(float), st));
        dim3 gx((unsigned)((M + TB - 1) / TB), (L.K + TB - 1) / TB);
        gemm_dx_kernel<16><<<gx, 256, 0, st>>>(dY, params + c.w_off[l], h->dact_conv[l - 1], conv_patch(L), M, L.filters, L.K, 1);
        count_launch(); count_launch(); count_launch(); count_launch();
    }
    {   // layer 1
        const ConvL& L = c.conv[0];
        const long long M = batch * L.P, n = M * L.filters;
        float* dY = h->dact_conv[0];
        relu_bwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(dY, h->act_conv[0], nullptr, n);
        colsum_kernel<<<dim3((L.filters + 31) / 32, (unsigned)std::min<long long>(64, (M + 7) / 8)), 256, 0, st>>>(dY, grads + c.b_off[0], M, L.filters);
        const size_t smem = (size_t)L.K * L.filters * sizeof(float);
        if (smem > 48 * 1024) QCUDA(cudaFuncSetAttribute(conv1_bits_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv1_bits_kernel<true><<<grid_for(M, 8, 148 * 4), 256, smem, st>>>((const u64*)packed, stride, batch, L, c.C, c.PW, c.H,
                                                                           nullptr, nullptr, nullptr, dY, grads + c.w_off[0]);
        count_launch(); count_launch(); count_launch();
    }
    (void)nt;
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

extern "C" int dq_qnet_activation(dq_qnet* h, int index, float** dev_ptr, int64_t* per_sample) {
    if (!h || !dev_ptr) return qfail(DQ_EINVAL, "NULL argument");
    const QCfg& c = h->c;
    if (index < 0 || index >= c.n_conv + c.n_fc) return qfail(DQ_EINVAL, "no such layer");
    if (index < c.n_conv) { *dev_ptr = h->act_conv[index]; if (per_sample) *per_sample = (int64_t)c.conv[index].P * c.conv[index].filters; }
    else { *dev_ptr = h->act_fc[index - c.n_conv]; if (per_sample) *per_sample = c.fc_out[index - c.n_conv]; }
    return DQ_OK;
}

extern "C" int dq_adam_step(float* params, float* m, float* v, const float* grads, int64_t n, float lr, float beta1, float beta2,
                            float eps, int64_t t, float grad_scale, dq_stream stream) {
    if (!params || !m || !v || !grads || n < 1 || t < 1) return qfail(DQ_EINVAL, "bad argument");
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
    DQ_LAUNCH_PDL(adam_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, params, m, v, grads, n, (float)lr_t, beta1, beta2, eps, grad_scale);
    count_launch();
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

extern "C" int dq_dqn_targets(const float* q_online_next, const float* q_target_next, const float* reward, const uint8_t* terminal,
                              float gamma, int64_t batch, int num_actions, float* y, dq_stream stream) {
    if (!q_online_next || !q_target_next || !reward || !terminal || !y || batch < 1) return qfail(DQ_EINVAL, "bad argument");
    DQ_LAUNCH_PDL(dqn_target_kernel, (unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream, q_online_next, q_target_next, reward, terminal, gamma, batch, num_actions, y);
    count_launch();
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

extern "C" int dq_dqn_loss_grad(const float* q, const int32_t* actions, const float* y, int64_t batch, int num_actions, float* dq,
                                float* stats, dq_stream stream) {
    if (!q || !actions || !y || !dq || batch < 1) return qfail(DQ_EINVAL, "bad argument");
    DQ_LAUNCH_PDL(dqn_loss_grad_kernel, (unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream, q, actions, y, batch, num_actions, dq, stats);
    count_launch();
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

extern "C" int dq_policy_eps_greedy(const float* q, const uint64_t* legal, int64_t n, int mask_words, int num_actions, uint32_t env_id_base,
                                    uint64_t seed, uint32_t step_index, uint32_t* dev_step_counter, double eps, int masked_greedy,
                                    int32_t* actions, dq_stream stream) {
    if (!q || !legal || !actions || n < 1 || mask_words < 1 || mask_words > 3) return qfail(DQ_EINVAL, "bad argument");
    double t = floor(eps * 4294967296.0);
    const u32 thr = eps <= 0.0 ? 0u : (t >= 4294967295.0 ? 0xFFFFFFFFu : (u32)t);
    DQ_LAUNCH_PDL(eps_greedy_kernel, (unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream, q, (const u64*)legal, (int)n, mask_words, num_actions, env_id_base,
                                                                                      step_index, dev_step_counter, (u32)seed, (u32)(seed >> 32), thr,
                                                                                      masked_greedy, actions);
    count_launch();
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

extern "C" int dq_replay_sample(const uint64_t* ring_obs, const int32_t* ring_act, const float* ring_rew, const uint8_t* ring_term,
                                int rows, int64_t npad, int64_t n, int capacity, int head, int filled, int64_t batch, uint64_t seed,
                                uint32_t draw_index, uint64_t* s0, uint64_t* s1, int32_t* act, float* rew, uint8_t* term,
                                int32_t* picked, dq_stream stream) {
    if (!ring_obs || !ring_act || !ring_rew || !ring_term || !s0 || !s1 || !act || !rew || !term) return qfail(DQ_EINVAL, "NULL argument");
    if (filled < 1 || filled >= capacity || batch < 1) return qfail(DQ_EINVAL, "replay ring holds no complete transition yet");
    DQ_LAUNCH_PDL(replay_sample_kernel, (unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream, (const u64*)ring_obs, ring_act, ring_rew, ring_term, rows, npad, (int)n,
                                                                                             capacity, head, filled, batch, (u32)seed, (u32)(seed >> 32), draw_index,
                                                                                             (u64*)s0, (u64*)s1, act, rew, term, picked);
    count_launch();
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

// ------------------------------------------------------------------------------------------------ folded head (acting)
// The layers after the last hidden dense layer are all linear: Dense(K -> A) (Function_Library.py:338-377), keras-rl's dueling
// Dense(A -> A+1) and the 'avg' combine Q_a = y_0 + y_{a+1} - mean_a' y_{a'+1}.  For inference they are ONE affine map
// Q = h * Wf + bf.
namespace dq {
__global__ void __launch_bounds__(128)
fold_head_kernel(const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ W3,
                 const float* __restrict__ b3, int K, int A, float* __restrict__ Wf, float* __restrict__ bf) {
    pdl_launch_dependents(); pdl_wait();
    // one warp per row r (r < K: row r of Wf [K][A]; r == K: bf); lane = column c of the dueling layer (c = 0: state value, c = 1 + a: advantage a)
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r > K) return;
    const float* x = r < K ? W2 + (size_t)r * A : b2;          // the row of Dense(K -> A) this warp pushes through the head
    float* out = r < K ? Wf + (size_t)r * A : bf;
    float y0 = 0.f, sum = 0.f;
    for (int c = lane; c <= A; c += 32) {
        float y = r < K ? 0.f : b3[c];
        for (int m = 0; m < A; ++m) y = fmaf(x[m], W3[(size_t)m * (A + 1) + c], y);
        if (c == 0) y0 = y;
        else { out[c - 1] = y; sum += y; }
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    y0 = __shfl_sync(0xffffffffu, y0, 0);
    const float mean = sum / (float)A;
    for (int c = lane; c <= A; c += 32) if (c >= 1) out[c - 1] = y0 + out[c - 1] - mean;      // each lane re-reads only what it wrote
}
}  // namespace dq

extern "C" int dq_qnet_fold_head(const dq_qnet* h, const float* params, float* w_out, float* b_out, dq_stream stream) {
    if (!h || !params || !w_out || !b_out) return qfail(DQ_EINVAL, "NULL argument");
    const QCfg& c = h->c;
    if (!c.dueling || c.n_fc < 2) return qfail(DQ_EINVAL, "the network has no dueling head to fold");
    const int i2 = c.n_fc - 2, i3 = c.n_fc - 1, t2 = c.n_conv + i2, t3 = c.n_conv + i3, K = c.fc_in[i2];
    if (c.fc_out[i2] != c.A || c.fc_in[i3] != c.A || c.fc_out[i3] != c.A + 1) return qfail(DQ_EINVAL, "unexpected head shape");
    DQ_LAUNCH_PDL(fold_head_kernel, (unsigned)((K + 1 + 3) / 4), 128, 0, (cudaStream_t)stream, params + c.w_off[t2], params + c.b_off[t2], params + c.w_off[t3],
                                                                                      params + c.b_off[t3], K, c.A, w_out, b_out);
    count_launch();
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

// ================================================================================================
// bf16 tensor-core inference path (acting): tcgen05.mma with TMEM accumulators.
//
// Every layer after the first is Y[M][N] = act(A[M][K] * W[K][N] + b) with A rows gathered from the previous
// channels-last activation (implicit im2col).  One CTA (128 threads) owns a 128 x BN output tile and the whole
// K extent: the 128 A rows and BN weight rows are copied into shared memory as K-major, 128-byte-swizzled
// tiles (one 1024-byte swizzle atom = 8 rows x 64 bf16) with 16-byte cp.async's, thread r gathering row r;
// one thread then issues K/16 tcgen05.mma (M=128, N=BN, K=16, bf16 x bf16 -> fp32 in TMEM), commits to an
// mbarrier, and the four warps read their 32 TMEM lanes back (tcgen05.ld 32x32b) for the fused
// bias + ReLU + convert epilogue.  Several CTAs are resident per SM, so one CTA's copies overlap another's MMAs.
// SASS evidence: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (commit), LDGSTS (cp.async).
#include <cuda_bf16.h>
#include <cuda.h>           // CUtensorMap (types only: the encoder is looked up through the runtime)

namespace dq {

struct TcArgs {
    const __nv_bfloat16* X; Patch g;            // AMODE 0: input activation (channels-last bf16) and its patch geometry
    const u64* packed; long long pstride;       // AMODE 1: packed observation rows [C*PW][pstride]
    int C, PW, H, T;                            //          input layers, words per layer, side, taps per layer (ksz*ksz)
    const __nv_bfloat16* Wt;                    // weights, transposed + zero-padded: [Npad][Kpad]
    const float* bias;
    void* Y; int ldy, out_bf16, relu;           // output rows of ldy elements
    long long M; int N, K, Kpad;
};

// [tcgen05 kernels: begin]  (tests/emu_qnet.py swaps the marked regions for plain loops with the same arguments; the product builds them as written)
__device__ __forceinline__ void cp_async16(u32 smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint64_t umma_smem_desc(u32 saddr) {
    // K-major, SWIZZLE_128B: 8-row groups 1024 B apart (SBO), LBO unused; version 1 (sm_100)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_bf16(u32 tmem_d, uint64_t adesc, uint64_t bdesc, u32 idesc, u32 accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}


// Epilogue of a 128 x BN accumulator tile: warp w owns TMEM lanes 32w..32w+31 = output rows; thread = one row.
template <int BN>
__device__ __forceinline__ void tc_epilogue(const TcArgs& a, u32 tmem, int warp, long long m, bool valid, int n0, const float* sbias,
                                            u32 stage_smem) {
    const u32 taddr = tmem + ((u32)(warp * 32) << 16);
    const bool vec_ok = (a.ldy & 7) == 0;
    const int lane = threadIdx.x & 31;
    // bf16 full-width tiles: a thread owns a ROW, so direct stores would touch 32 lines per warp instruction.  The warp's
    // 32 x BN tile goes through shared memory (the operand stages are free once the MMAs have completed; 16-byte pieces
    // XOR-swizzled by row) and leaves as full lines: BN/8 consecutive lanes write one row.
    const bool staged = a.out_bf16 && vec_ok && n0 + BN <= a.N;
    // fp32 output whose rows are dense (ldy == N, one N tile, e.g. the Dense -> num_actions layer): staged the same way
    const bool dense32 = !a.out_bf16 && a.ldy == a.N && a.N <= BN && gridDim.y == 1;
    const u32 wbase = stage_smem + (u32)warp * (u32)(32 * BN * (dense32 ? 4 : 2));      // this warp's 32-row staging tile
    constexpr int PR = BN / 8;                          // 16-byte pieces per row
    constexpr int SW = (PR < 8 ? PR : 8) - 1;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
        u32 v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr + (u32)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid || staged || dense32) {
            float f[16];
            const float4* sb4 = reinterpret_cast<const float4*>(sbias + c0);       // four 128-bit broadcast loads
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 bq = sb4[j4];
                const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float x = __uint_as_float(v[4 * j4 + u]) + bb[u];
                    f[4 * j4 + u] = a.relu ? fmaxf(x, 0.f) : x;
                }
            }
            if (a.out_bf16) {
                __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(a.Y) + m * a.ldy + n0 + c0;
                if (n0 + c0 + 16 <= a.N && vec_ok) {
                    uint4 p0, p1;
                    __nv_bfloat162 t;
#define DQ_PACK(dst, i) t = __floats2bfloat162_rn(f[i], f[i + 1]); dst = *reinterpret_cast<u32*>(&t);
                    DQ_PACK(p0.x, 0) DQ_PACK(p0.y, 2) DQ_PACK(p0.z, 4) DQ_PACK(p0.w, 6)
                    DQ_PACK(p1.x, 8) DQ_PACK(p1.y, 10) DQ_PACK(p1.z, 12) DQ_PACK(p1.w, 14)
#undef DQ_PACK
                    if (staged) {
                        const int q0 = c0 >> 3;
                        const u32 d0 = wbase + (u32)lane * (u32)(BN * 2) + (u32)(((q0) ^ (lane & SW)) << 4);
                        const u32 d1 = wbase + (u32)lane * (u32)(BN * 2) + (u32)(((q0 + 1) ^ (lane & SW)) << 4);
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(d0), "r"(p0.x), "r"(p0.y), "r"(p0.z), "r"(p0.w) : "memory");
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(d1), "r"(p1.x), "r"(p1.y), "r"(p1.z), "r"(p1.w) : "memory");
                    } else {
                        reinterpret_cast<uint4*>(y)[0] = p0; reinterpret_cast<uint4*>(y)[1] = p1;
                    }
                } else {
                    for (int j = 0; j < 16; ++j) if (n0 + c0 + j < a.N) y[j] = __float2bfloat16(f[j]);
                }
            } else if (dense32) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < a.N) asm volatile("st.shared.f32 [%0], %1;" ::"r"(wbase + (u32)((lane * a.N + c0 + j) * 4)), "f"(f[j]) : "memory");
            } else {
                float* y = reinterpret_cast<float*>(a.Y) + m * a.ldy + n0 + c0;
                for (int j = 0; j < 16; ++j) if (n0 + c0 + j < a.N) y[j] = f[j];
            }
        }
    }
    if (dense32) {                                      // the warp's 32 rows are one contiguous span of the fp32 output
        __syncwarp();
        const long long mw = m - lane;
        const long long rows = min(32ll, a.M - mw);
        float* yb = reinterpret_cast<float*>(a.Y) + mw * a.N;
        const int tot = rows > 0 ? (int)rows * a.N : 0;
        for (int i = lane; i < tot; i += 32) {
            float val;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val) : "r"(wbase + (u32)(i * 4)));
            yb[i] = val;
        }
    }
    if (staged) {
        __syncwarp();
        const long long mw = m - lane;                  // first row of this warp
        __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(a.Y);
#pragma unroll
        for (int it = 0; it < PR; ++it) {
            const int idx = it * 32 + lane, row = idx / PR, q = idx % PR;
            uint4 val;
            const u32 src = wbase + (u32)row * (u32)(BN * 2) + (u32)((q ^ (row & SW)) << 4);
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(src));
            if (mw + row < a.M) *reinterpret_cast<uint4*>(yb + (mw + row) * a.ldy + n0 + q * 8) = val;
        }
    }
}

// AMODE 0: A rows gathered from a bf16 activation (implicit im2col), 16-byte cp.async per chunk.
// AMODE 1: A rows expanded from the packed binary observation: K order is (layer, tap), so a row's K bits are the
//          per-layer tap masks concatenated; 8 bits -> 8 bf16 {0, 1.0} per 16-byte chunk.
template <int BN, int AMODE>
__global__ void __launch_bounds__(128)
tc_gemm_kernel(const TcArgs a) {
    extern __shared__ unsigned char tc_raw[];
    __shared__ alignas(8) u64 mbar;
    __shared__ u32 tmem_slot;
    __shared__ int koff[128];                                       // AMODE 0: element offset of every 8-wide k chunk (K <= 1024)
    __shared__ alignas(16) float sbias[BN];
    __shared__ u64 swords[AMODE == 1 ? 36 * 16 : 1];                // AMODE 1: packed words of the tile's samples
    __shared__ alignas(16) uint4 bf16lut[AMODE == 1 ? 256 : 1];     // AMODE 1: byte -> its 8 bits as 8 bf16 {0, 1.0}
    const int tid = threadIdx.x, warp = tid >> 5;
    const int KB = a.Kpad >> 6;                                     // 64-element k blocks
    const u32 s_base = (smem_u32(tc_raw) + 1023u) & ~1023u;         // swizzle atoms need 1024-byte alignment
    const u32 sA = s_base, sB = s_base + (u32)KB * 16384u;
    const long long m0 = (long long)blockIdx.x * 128;
    const int n0 = blockIdx.y * BN;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((u32)(BN < 32 ? 32 : BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) mbar_init(&mbar, 1);
    if (warp == 0) pdl_launch_dependents();                         // by the warp that has just taken the CTA's TMEM columns: a successor cannot starve it
    pdl_wait();                                                     // first global access below
    if (tid < BN) sbias[tid] = (a.bias && n0 + tid < a.N) ? a.bias[n0 + tid] : 0.f;
    {   // B: weight rows, spread over the CTA (independent of A)
        const int total = BN * KB * 8;
        for (int i = tid; i < total; i += 128) {
            const int c = i & 7, n = (i >> 3) % BN, kb = (i >> 3) / BN;
            const u32 dst = sB + (u32)kb * (u32)(BN * 128) + (u32)n * 128u + (u32)((c ^ (n & 7)) << 4);
            cp_async16(dst, a.Wt + (size_t)(n0 + n) * a.Kpad + kb * 64 + c * 8);
        }
    }
    const int r = tid;
    const long long m = m0 + r;
    const bool valid = m < a.M;
    if (AMODE == 0) {
        if (tid < KB * 8) koff[tid] = (tid * 8 < a.K) ? patch_col(a.g, tid * 8) : -1;
        const long long rowoff = valid ? patch_row(a.g, m) : 0;
        __syncthreads();
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int ko = koff[kb * 8 + c];
                const u32 dst = sA + (u32)kb * 16384u + (u32)r * 128u + (u32)((c ^ (r & 7)) << 4);
                if (valid && ko >= 0) cp_async16(dst, a.X + rowoff + ko);
                else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
            }
        }
    } else {
        for (int x = tid; x < 256; x += 128)
            bf16lut[x] = make_uint4(((x & 1u) * 0x3F80u) | ((x & 2u) * 0x1FC00000u), (((x >> 2) & 1u) * 0x3F80u) | (((x >> 2) & 2u) * 0x1FC00000u),
                                    (((x >> 4) & 1u) * 0x3F80u) | (((x >> 4) & 2u) * 0x1FC00000u), (((x >> 6) & 1u) * 0x3F80u) | (((x >> 6) & 2u) * 0x1FC00000u));
        // stage the packed words of the samples this tile touches
        const int P = a.g.P, rows = a.C * a.PW;
        const long long b_lo = m0 / P;
        const long long b_hi = min((m0 + 127) / P, (a.M - 1) / P);
        const int ns = (int)(b_hi - b_lo + 1);
        for (int i = tid; i < ns * rows; i += 128) {
            const int s = i / rows, w = i - s * rows;
            swords[s * 36 + w] = a.packed[(long long)w * a.pstride + b_lo + s];
        }
        __syncthreads();
        const long long b = valid ? m / P : b_lo;
        const int pos = valid ? (int)(m - b * P) : 0, oy = pos / a.g.oh, ox = pos - oy * a.g.oh;
        const u64* wds = swords + (int)(b - b_lo) * 36;
        // this row's K bits: bit k = (layer k / T, tap k % T); per layer the ksz row segments of the patch are cut out of the
        // layer bitmap with one funnel shift each (32-bit view: a segment never needs bits of a third word)
        u64 kbits[3] = {0, 0, 0};
        if (valid) {
            const u32* w32 = reinterpret_cast<const u32*>(wds);
            const int ksz = a.g.ksz, b0 = oy * a.g.stride * a.H + ox * a.g.stride;
            const u32 kmask = (1u << ksz) - 1u;
            for (int ci = 0; ci < a.C; ++ci) {
                const u32* lw = w32 + ci * a.PW * 2;
                u32 taps = 0;
#pragma unroll 4
                for (int ky = 0; ky < ksz; ++ky) {
                    const int bit = b0 + ky * a.H, wi = bit >> 5;
                    taps |= (__funnelshift_r(lw[wi], lw[wi + 1], bit & 31) & kmask) << (ky * ksz);
                }
                const int pos = ci * a.T, wq = pos >> 6, sh = pos & 63;
                const u64 tv = (u64)taps;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (q == wq) kbits[q] |= tv << sh;
                    if (q == wq + 1 && sh) kbits[q] |= tv >> (64 - sh);
                }
            }
        }
        for (int kb = 0; kb < KB; ++kb) {
            const u64 bits = kb == 0 ? kbits[0] : (kb == 1 ? kbits[1] : kbits[2]);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const u32 x = (u32)(bits >> (8 * c)) & 0xFFu;
                const u32 dst = sA + (u32)kb * 16384u + (u32)r * 128u + (u32)((c ^ (r & 7)) << 4);
                const uint4 w = bf16lut[x];                                              // bf16 1.0 = 0x3F80
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_proxy_async();                                            // generic-proxy writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem = tmem_slot;

    if (tid == 0) {
        // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
        const u32 idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(BN >> 3) << 17) | ((u32)(128 >> 4) << 24);
        for (int kb = 0; kb < KB; ++kb) {
            const uint64_t da = umma_smem_desc(sA + (u32)kb * 16384u), db = umma_smem_desc(sB + (u32)kb * (u32)(BN * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)                             // 16 bf16 = 32 bytes along the swizzled row
                umma_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    mbar_wait(&mbar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    tc_epilogue<BN>(a, tmem, warp, m, valid, n0, sbias, s_base);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((u32)(BN < 32 ? 32 : BN)) : "memory");
}

// AMODE 0 with K > 64: the same tile, but K is streamed through a two-stage ring of 64-wide chunks (A chunk 128 x 128 B +
// W chunk BN x 128 B per stage): the cp.async copies of chunk k+1 are in flight while the MMAs of chunk k run, a stage is
// recycled when the tcgen05.commit of the MMAs that read it has arrived on its mbarrier, and the small footprint
// (40-64 KB instead of 83-198 KB) keeps 3-5 CTAs resident per SM.
__device__ __forceinline__ void mbar_wait_or_trap(u64* bar, u32 parity) {
    const long long t0 = clock64();
    while (clock64() - t0 < 4000000000ll) {             // ~2 s of SM clocks
        u32 ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();                                           // a lost arrival must fail the launch, not hang the GPU
}

template <int BN, int S>
__global__ void __launch_bounds__(128)
tc_gemm_pipe_kernel(const TcArgs a) {
    constexpr u32 STAGE = 16384u + (u32)BN * 128u;
    extern __shared__ unsigned char tc_raw[];
    __shared__ alignas(8) u64 mbar_free[S];
    __shared__ u32 tmem_slot;
    __shared__ int koff[128];                                       // element offset of every 8-wide k chunk (K <= 1024)
    __shared__ long long rowoff_s[128];
    __shared__ alignas(16) float sbias[BN];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int KB = a.Kpad >> 6;
    const u32 s_base = (smem_u32(tc_raw) + 1023u) & ~1023u;
    const long long m0 = (long long)blockIdx.x * 128;
    const int n0 = blockIdx.y * BN;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((u32)(BN < 32 ? 32 : BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) for (int s = 0; s < S; ++s) mbar_init(&mbar_free[s], 1);
    if (warp == 0) pdl_launch_dependents();
    pdl_wait();
    if (tid < BN) sbias[tid] = (a.bias && n0 + tid < a.N) ? a.bias[n0 + tid] : 0.f;
    if (tid < KB * 8) koff[tid] = (tid * 8 < a.K) ? patch_col(a.g, tid * 8) : -1;
    const int r = tid;
    const long long m = m0 + r;
    const bool valid = m < a.M;
    rowoff_s[r] = valid ? patch_row(a.g, m) : -1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem = tmem_slot;
    const u32 idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(BN >> 3) << 17) | ((u32)(128 >> 4) << 24);

    auto load_chunk = [&](int kb) {
        const u32 sA = s_base + (u32)(kb % S) * STAGE, sB = sA + 16384u;
        // 8 consecutive lanes copy the 8 x 16 B of one row's chunk (one full 128-byte line per row), 4 rows per warp instruction
        const int c = tid & 7, ko = koff[kb * 8 + c];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int rr = (tid >> 3) + 16 * j;
            const long long ro = rowoff_s[rr];
            const u32 dst = sA + (u32)rr * 128u + (u32)((c ^ (rr & 7)) << 4);
            if (ro >= 0 && ko >= 0) cp_async16(dst, a.X + ro + ko);
            else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
        }
        for (int i = tid; i < BN * 8; i += 128) {
            const int c = i & 7, n = i >> 3;
            cp_async16(sB + (u32)n * 128u + (u32)((c ^ (n & 7)) << 4), a.Wt + (size_t)(n0 + n) * a.Kpad + kb * 64 + c * 8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // S - 1 chunks ahead of the math; one cp.async group is committed per iteration (possibly empty) so that
    // "all but the newest S - 1 groups have landed" always means "chunk kb has landed"
    for (int kb = 0; kb < S - 1; ++kb) {
        if (kb < KB) load_chunk(kb);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kb = 0; kb < KB; ++kb) {
        const int nk = kb + S - 1;                                  // the chunk to prefetch now
        if (nk < KB) {
            if (nk >= S) mbar_wait_or_trap(&mbar_free[nk % S], (u32)((nk / S - 1) & 1));   // its stage's previous readers are done
            load_chunk(nk);
        } else {
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
        fence_proxy_async();                                        // generic-proxy writes -> visible to the MMA (async proxy)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const u32 sA = s_base + (u32)(kb % S) * STAGE;
            const uint64_t da = umma_smem_desc(sA), db = umma_smem_desc(sA + 16384u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar_free[kb % S])) : "memory");
        }
    }
    mbar_wait_or_trap(&mbar_free[(KB - 1) % S], (u32)(((KB - 1) / S) & 1));      // the last commit covers every MMA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tc_epilogue<BN>(a, tmem, warp, m, valid, n0, sbias, s_base);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((u32)(BN < 32 ? 32 : BN)) : "memory");
}


struct Tmap2D { CUtensorMap m; };
// bf16 matrix [rows][ld] (ld % 8 == 0) as a tiled tensor map with boxes of 64 columns x box_rows rows; columns >= cols and rows >= rows read as zero
static bool make_tmap_2d(Tmap2D* out, const __nv_bfloat16* base, long long rows, long long ld, int box_rows, long long cols = -1) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn fn = [] {         // the driver entry point through the runtime: the library does not link libcuda
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (EncodeFn)f;
    }();
    if (!fn) return false;
    // a tensor map is a pure function of these arguments: the last few are kept (an update re-encodes the same ~30 maps every time)
    struct Key { const void* base; long long rows, ld, cols; int box; };
    struct Slot { Key k; Tmap2D m; };
    static thread_local Slot cache[64];
    static thread_local int used = 0, next = 0;
    const Key key{base, rows, ld, cols, box_rows};
    for (int i = 0; i < used; ++i)
        if (cache[i].k.base == key.base && cache[i].k.rows == key.rows && cache[i].k.ld == key.ld && cache[i].k.cols == key.cols && cache[i].k.box == key.box) { *out = cache[i].m; return true; }
    const cuuint64_t dims[2] = {(cuuint64_t)(cols < 0 ? ld : cols), (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    if (fn(&out->m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    const int slot = used < 64 ? used++ : (next = (next + 1) % 64);
    cache[slot].k = key; cache[slot].m = *out;
    return true;
}
__device__ __forceinline__ void tma_load_2d(u32 smem_dst, const Tmap2D* map, int c_inner, int c_outer, u64* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer) : "memory");
}
// Dense layers (A = a plain [M][K] bf16 matrix, K <= 512): every K block of the tile has its own shared-memory stage and "full" mbarrier,
// so ONE producer thread puts all of the tile's loads in flight at once (TMA, SWIZZLE_128B) -- a dense layer of this network is a
// handful of CTAs whose time is the latency of their loads, not bandwidth -- and the MMA thread consumes the stages in order; the
// epilogue is the forward kernels' (bias + ReLU + convert, rows staged through the freed operand memory).
template <int BN>
__global__ void __launch_bounds__(128)
tc_gemm_tma_kernel(const __grid_constant__ Tmap2D ta, const __grid_constant__ Tmap2D tw, const TcArgs a) {
    constexpr u32 STAGE = 16384u + (u32)BN * 128u;
    extern __shared__ unsigned char tc_raw[];
    __shared__ alignas(8) u64 bar_full[8], bar_accum;
    __shared__ u32 tmem_slot;
    __shared__ alignas(16) float sbias[BN];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = a.Kpad >> 6;                                     // <= 8 (host-checked)
    const u32 s_base = (smem_u32(tc_raw) + 1023u) & ~1023u;
    const long long m0 = (long long)blockIdx.x * 128;
    const int n0 = blockIdx.y * BN;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((u32)(BN < 32 ? 32 : BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < KB; ++s) mbar_init(&bar_full[s], 1);
        mbar_init(&bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == 0) pdl_launch_dependents();
    pdl_wait();
    if (tid < BN) sbias[tid] = (a.bias && n0 + tid < a.N) ? a.bias[n0 + tid] : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem = tmem_slot;
    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full[kb])), "r"(STAGE) : "memory");
                const u32 sA = s_base + (u32)kb * STAGE;
                tma_load_2d(sA, &ta, kb * 64, (int)m0, &bar_full[kb]);
                tma_load_2d(sA + 16384u, &tw, kb * 64, n0, &bar_full[kb]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const u32 idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(BN >> 3) << 17) | ((u32)(128 >> 4) << 24);
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait_or_trap(&bar_full[kb], 0u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const u32 sA = s_base + (u32)kb * STAGE;
                const uint64_t da = umma_smem_desc(sA), db = umma_smem_desc(sA + 16384u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_accum)) : "memory");
        }
        __syncwarp();
    }
    mbar_wait_or_trap(&bar_accum, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long m = m0 + tid;
    tc_epilogue<BN>(a, tmem, warp, m, m < a.M, n0, sbias, s_base);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((u32)(BN < 32 ? 32 : BN)) : "memory");
}
// [tcgen05 kernels: end]

// Fused head for the acting path: y = x W + b for the last (tiny) dense layer, then the dueling combination.
// 4 adjacent lanes share a sample; lane p owns the 4-column groups p, p+4, p+8, ... of the N outputs, so every weight read is
// one 128-bit shared load shared by all samples of the warp.  W (rows padded to a multiple of 4), the bias and the CTA's 32
// input rows are staged in shared memory with all their global loads in flight together.
constexpr int kHeadSamples = 32;
template <int MAXG>                                     // 4-column groups per lane: N <= 16 * MAXG
__global__ void __launch_bounds__(128)
head_dueling_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ q, long long B, int K, int N, int A, int dueling) {
    pdl_launch_dependents(); pdl_wait();
    extern __shared__ __align__(16) float hw[];         // W [K][N4], bias [N4], x tile [32][K]
    const int N4 = (N + 3) & ~3, G = N4 >> 2;
    float* hb = hw + K * N4;
    float* hx = hb + N4;
    const int tid = threadIdx.x;
    const long long b0 = (long long)blockIdx.x * kHeadSamples;
    const int nrow = (int)min((long long)kHeadSamples, B - b0);
    {
        const int nw = K * N, tot = nrow * K;
        const float* xs = x + b0 * K;
        // one L2 round trip for the whole prologue: every load (24 of W, 16 of the row tile per thread) is issued before the
        // first shared store; anything beyond 3072 weights / 2048 inputs goes through the batched loops below
        float w24[24], x16[16];
#pragma unroll
        for (int j = 0; j < 24; ++j) w24[j] = __ldg(W + min(j * 128 + tid, nw - 1));
#pragma unroll
        for (int j = 0; j < 16; ++j) x16[j] = __ldg(xs + min(j * 128 + tid, tot - 1));
        const float bv = tid < N ? __ldg(bias + tid) : 0.f;
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            const int i = j * 128 + tid;
            if (i < nw) { const int k = i / N; hw[k * N4 + (i - k * N)] = w24[j]; }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int i = j * 128 + tid;
            if (i < tot) hx[i] = x16[j];
        }
        if (tid < N4) hb[tid] = bv;
        for (int i0 = 24 * 128; i0 < nw; i0 += 16 * 128) {
            float w16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w16[j] = __ldg(W + min(i0 + j * 128 + tid, nw - 1));
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int i = i0 + j * 128 + tid;
                if (i < nw) { const int k = i / N; hw[k * N4 + (i - k * N)] = w16[j]; }
            }
        }
        if (N4 != N) for (int k = tid; k < K; k += 128) for (int c = N; c < N4; ++c) hw[k * N4 + c] = 0.f;
        for (int i0 = 16 * 128; i0 < tot; i0 += 16 * 128) {
            float y16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) y16[j] = __ldg(xs + min(i0 + j * 128 + tid, tot - 1));
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int i = i0 + j * 128 + tid;
                if (i < tot) hx[i] = y16[j];
            }
        }
    }
    __syncthreads();
    const int s = tid >> 2, part = tid & 3;
    const bool live = s < nrow;
    const int cnt = (G - part + 3) >> 2;                // groups part, part+4, ... < G
    float4 y[MAXG];
#pragma unroll
    for (int i = 0; i < MAXG; ++i) y[i] = (i < cnt) ? reinterpret_cast<const float4*>(hb)[part + 4 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        const float* xr = hx + s * K;
        for (int k = 0; k < K; ++k) {
            const float xv = xr[k];
            const float4* wr = reinterpret_cast<const float4*>(hw + k * N4) + part;
#pragma unroll
            for (int i = 0; i < MAXG; ++i) {
                if (i < cnt) {
                    const float4 w = wr[4 * i];
                    y[i].x = fmaf(xv, w.x, y[i].x); y[i].y = fmaf(xv, w.y, y[i].y);
                    y[i].z = fmaf(xv, w.z, y[i].z); y[i].w = fmaf(xv, w.w, y[i].w);
                }
            }
        }
    }
    const long long b = b0 + s;
    if (dueling) {
        float sum = 0.f;                                // advantages are outputs 1..N-1, the state value is output 0
#pragma unroll
        for (int i = 0; i < MAXG; ++i) {
            const int c0 = 4 * (part + 4 * i);
            const float e[4] = {y[i].x, y[i].y, y[i].z, y[i].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i < cnt && c0 + u >= 1 && c0 + u < N) sum += e[u];
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float v0 = __shfl_sync(0xffffffffu, y[0].x, (tid & 31) & ~3);
        const float base = v0 - sum / (float)A;
        if (live) {
#pragma unroll
            for (int i = 0; i < MAXG; ++i) {
                const int c0 = 4 * (part + 4 * i);
                const float e[4] = {y[i].x, y[i].y, y[i].z, y[i].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) if (i < cnt && c0 + u >= 1 && c0 + u < N) q[b * A + c0 + u - 1] = base + e[u];
            }
        }
    } else if (live) {
#pragma unroll
        for (int i = 0; i < MAXG; ++i) {
            const int c0 = 4 * (part + 4 * i);
            const float e[4] = {y[i].x, y[i].y, y[i].z, y[i].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i < cnt && c0 + u < N) q[b * A + c0 + u] = e[u];
        }
    }
}
// ================================================================================================
// bf16 tensor-core TRAINING path (opt-in: DQNAgent(train_precision="bf16")).  Forward = the tcgen05 layers above (unfolded head,
// dropout on the bf16 activation); backward = per tensor-core layer j, with G_j the gradient at the layer's output:
//   prep_dy      G_j (fp32, bf16, or gathered out of layer j+1's column gradient: col2im) x ReLU / dropout mask
//                -> dYb [M][Npad] and its transpose dYT [Npad'][Mpad] (bf16), bias gradient = column sums
//   im2colT      A_j^T [Kpad][Mpad] (bf16): the layer's patch matrix, transposed (layer 1: straight from the packed bits)
//   tc_dw_tma    dW_j[K][N] += A_j^T x dYT^T   -- contraction over the M = batch x positions rows, split over CTAs (grid.z), TMA-fed, fp32
//                atomics into the caller's gradient buffer
//   tc_gemm      dCol_j[M][K] = dYb x W_j^T    -- the forward kernel with W_j's bf16 copy [K][Npad] (staged by dq_qnet_prepare_tc) as the "transposed weight"
// Both GEMM operands are K-major for tcgen05 because the transposes are materialised (bf16, a few tens of MB at batch 4096).
// fp32 master weights, fp32 accumulation in TMEM, fp32 gradients / Adam: the usual mixed-precision recipe.

// [tcgen05 kernels: begin]
// D[R][N] += At[R][k] * Bt[N][k]^T over this CTA's k blocks (grid.z slices of the contraction).  At rows are padded to a multiple of
// 128 (zeros), Bt rows to a multiple of BN, both leading dimensions to a multiple of 64 elements; 128 x BN fp32 accumulator in TMEM.
// Both operand tiles are brought in by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B tensor maps: the hardware writes the
// K-major swizzle atoms the UMMA descriptors expect) and a warp-specialised pipeline: one producer thread arms a stage's "full" mbarrier
// with the stage's byte count and issues the two tile loads, one MMA thread waits for "full", issues the four tcgen05.mma of the chunk
// and commits them to the stage's "empty" mbarrier; all four warps drain the accumulator at the end.  No thread of the CTA touches the
// operand bytes.  SASS: UTMALDG, UTCHMMA, UTCBAR, SYNCS.  Measured against the same GEMM fed by a 128-thread cp.async ring (the forward
// kernel's scheme): 5-9 % less time per dW launch, 529 -> 521 us per batch-4096 update.
template <int BN, int S>
__global__ void __launch_bounds__(128)
tc_dw_tma_kernel(const __grid_constant__ Tmap2D ta, const __grid_constant__ Tmap2D tb, float* __restrict__ D, int ldd, int R, int N,
                 int kb_total, int kb_per_cta) {
    constexpr u32 STAGE = 16384u + (u32)BN * 128u;
    extern __shared__ unsigned char tc_raw[];
    __shared__ alignas(8) u64 bar_full[S], bar_empty[S], bar_accum;
    __shared__ u32 tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 s_base = (smem_u32(tc_raw) + 1023u) & ~1023u;
    const int r0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
    const int kb0 = blockIdx.z * kb_per_cta, KB = min(kb_per_cta, kb_total - kb0);
    if (KB <= 0) return;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((u32)(BN < 32 ? 32 : BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();                                        // the barriers are next touched by the async proxy (TMA, tcgen05.commit)
    }
    if (warp == 0) pdl_launch_dependents();
    pdl_wait();                                                     // the TMA loads and the gradient atomics come after this point
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem = tmem_slot;
    if (warp == 0) {
        if (lane == 0) {                                            // ---- producer
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % S;
                if (kb >= S) mbar_wait_or_trap(&bar_empty[s], (u32)((kb / S - 1) & 1));      // the MMAs that read this stage have completed
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full[s])), "r"(STAGE) : "memory");
                const u32 sA = s_base + (u32)s * STAGE;
                tma_load_2d(sA, &ta, (kb0 + kb) * 64, r0, &bar_full[s]);
                tma_load_2d(sA + 16384u, &tb, (kb0 + kb) * 64, n0, &bar_full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {                                            // ---- MMA issuer
            const u32 idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(BN >> 3) << 17) | ((u32)(128 >> 4) << 24);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % S;
                mbar_wait_or_trap(&bar_full[s], (u32)((kb / S) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const u32 sA = s_base + (u32)s * STAGE;
                const uint64_t da = umma_smem_desc(sA), db = umma_smem_desc(sA + 16384u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_empty[s])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_accum)) : "memory");
        }
        __syncwarp();
    }
    mbar_wait_or_trap(&bar_accum, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 taddr = tmem + ((u32)(warp * 32) << 16);
    const int row = r0 + tid;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
        u32 v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr + (u32)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < R) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float f = __uint_as_float(v[j]);
                if (n0 + c0 + j < N && f != 0.f) atomicAdd(D + (long long)row * ldd + n0 + c0 + j, f);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((u32)(BN < 32 ? 32 : BN)) : "memory");
}
// [tcgen05 kernels: end]

// A^T[k][m] (bf16, rows 0..Kpad-1, columns 0..Mpad-1, zero outside K x M) of a patch matrix over a channels-last bf16 activation.
// 64 x 64 tiles through shared memory.  cin % 8 == 0 (tc_check), so 8 consecutive k of a row are 16 contiguous bytes of the input:
// a thread loads 8 k of one row, then stores 8 m of one k.
__global__ void __launch_bounds__(256)
im2colT_kernel(const __nv_bfloat16* __restrict__ X, Patch g, long long M, int K, __nv_bfloat16* __restrict__ At, long long lda) {
    pdl_launch_dependents(); pdl_wait();
    __shared__ __align__(16) unsigned short tile[64][72];        // [k][m]; rows of 144 bytes keep the 16-byte reads of phase 2 aligned
    __shared__ long long rowoff[64];
    __shared__ int coloff[8];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * 64;
    const int k0 = blockIdx.y * 64;
    if (tid < 64) rowoff[tid] = (m0 + tid < M) ? patch_row(g, m0 + tid) : -1;
    else if (tid < 72) coloff[tid - 64] = (k0 + (tid - 64) * 8 < K) ? patch_col(g, k0 + (tid - 64) * 8) : -1;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int item = tid + i * 256, mm = item >> 3, kc = item & 7;
        const long long ro = rowoff[mm];
        const int co = coloff[kc];
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ro >= 0 && co >= 0) v = *reinterpret_cast<const uint4*>(X + ro + co);
        const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tile[kc * 8 + 2 * j][mm] = (unsigned short)(w[j] & 0xFFFFu);
            tile[kc * 8 + 2 * j + 1][mm] = (unsigned short)(w[j] >> 16);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int item = tid + i * 256, kk = item >> 3, mc = item & 7;
        *reinterpret_cast<uint4*>(At + (long long)(k0 + kk) * lda + m0 + mc * 8) = *reinterpret_cast<const uint4*>(&tile[kk][mc * 8]);
    }
}
// Layer 1: the same matrix straight from the packed observation bits, rows in the parameter order k = tap * C + layer.
__global__ void __launch_bounds__(256)
im2colT_bits_kernel(const u64* __restrict__ packed, long long stride, ConvL L, int C, int PW, int H, long long M,
                    __nv_bfloat16* __restrict__ At, long long lda, int Kpad) {
    pdl_launch_dependents(); pdl_wait();
    const long long m = (long long)blockIdx.x * 256 + threadIdx.x;          // column (< lda)
    if (m >= lda) return;
    const bool live = m < M;
    const long long b = live ? m / L.P : 0;
    const int pos = live ? (int)(m - b * L.P) : 0, oy = pos / L.oh, ox = pos - oy * L.oh;
    const int T = L.ksz * L.ksz;
    for (int ci = 0; ci < C; ++ci) {
        const u32 taps = live ? layer_taps(packed, stride, b, ci, PW, H, oy * L.stride, ox * L.stride, L.ksz) : 0u;
        for (int t = 0; t < T; ++t) At[(long long)(t * C + ci) * lda + m] = __float2bfloat16((float)((taps >> t) & 1u));
    }
    if (blockIdx.y == 0) for (int k = T * C; k < Kpad; ++k) At[(long long)k * lda + m] = __float2bfloat16(0.f);
}

// The gradient at a layer's output, masked and laid out for the two GEMMs of its backward step.
//   src_mode 0: fp32 [M][N];  1: bf16 [M][N];  2: col2im -- G[m = (b, iy, ix)][c] = sum over the taps (ky, kx) of the layer ABOVE that
//   read this cell of dcol[(b, oy, ox)][(ky*ksz + kx)*N + c]  (up: the patch geometry of the layer above, Kup its K)
struct DySrc { const void* p; int mode; Patch up; int Kup; };
__device__ __forceinline__ void bf16x8_to_float(const uint4& v, float* f) {
    const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(w[j] << 16); f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u); }
}
// 64 (rows m) x 64 (channels n) per CTA; a thread owns 8 consecutive channels of a row (N % 8 == 0 on the vector paths: one 16-byte load per
// source, the cell geometry of the col2im gather computed once per 8 channels), then 8 consecutive rows of a channel of the transpose.
__global__ void __launch_bounds__(256)
prep_dy_kernel(DySrc src, const __nv_bfloat16* __restrict__ act, const float* __restrict__ mask, long long M, int N,
               __nv_bfloat16* __restrict__ dYb, int ldyb, __nv_bfloat16* __restrict__ dYT, long long ldyt, int rows_t,
               float* __restrict__ db) {
    pdl_launch_dependents(); pdl_wait();
    __shared__ float tile[64][65];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    const bool vec = (N & 7) == 0;
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
        const int item = tid + i * 256, mm = item >> 3, nn = (item & 7) * 8;
        const long long m = m0 + mm;
        const int n = n0 + nn;
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
        if (m < M && n < N) {
            if (src.mode == 0) {
                const float* sp = reinterpret_cast<const float*>(src.p) + m * N + n;
#pragma unroll
                for (int j = 0; j < 8; ++j) if (n + j < N) g[j] = sp[j];
            } else if (src.mode == 1) {
                const __nv_bfloat16* sp = reinterpret_cast<const __nv_bfloat16*>(src.p) + m * N + n;
                if (vec) bf16x8_to_float(*reinterpret_cast<const uint4*>(sp), g);
                else for (int j = 0; j < 8; ++j) if (n + j < N) g[j] = __bfloat162float(sp[j]);
            } else {
                const Patch& u = src.up;                         // this layer's output map is the upper layer's input map: side u.ih
                const int cells = u.ih * u.ih;
                const long long b = m / cells;
                const int pos = (int)(m - b * cells), iy = pos / u.ih, ix = pos - iy * u.ih;
                const __nv_bfloat16* dc = reinterpret_cast<const __nv_bfloat16*>(src.p);
                for (int ky = 0; ky < u.ksz; ++ky) {
                    const int ty = iy - ky;
                    if (ty < 0 || ty % u.stride) continue;
                    const int oy = ty / u.stride;
                    if (oy >= u.oh) continue;
                    for (int kx = 0; kx < u.ksz; ++kx) {
                        const int tx = ix - kx;
                        if (tx < 0 || tx % u.stride) continue;
                        const int ox = tx / u.stride;
                        if (ox >= u.oh) continue;
                        const __nv_bfloat16* sp = dc + ((b * u.P + oy * u.oh + ox) * (long long)src.Kup) + (ky * u.ksz + kx) * N + n;
                        if (vec) {
                            float t[8];
                            bf16x8_to_float(*reinterpret_cast<const uint4*>(sp), t);
#pragma unroll
                            for (int j = 0; j < 8; ++j) g[j] += t[j];
                        } else {
                            for (int j = 0; j < 8; ++j) if (n + j < N) g[j] += __bfloat162float(sp[j]);
                        }
                    }
                }
            }
            if (act) {
                float a8[8];
                if (vec) bf16x8_to_float(*reinterpret_cast<const uint4*>(act + m * N + n), a8);
                else for (int j = 0; j < 8; ++j) a8[j] = (n + j < N) ? __bfloat162float(act[m * N + n + j]) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) if (!(a8[j] > 0.f)) g[j] = 0.f;
            }
            if (mask) {
                if (vec) {
                    const float4 k0 = *reinterpret_cast<const float4*>(mask + m * N + n), k1 = *reinterpret_cast<const float4*>(mask + m * N + n + 4);
                    g[0] *= k0.x; g[1] *= k0.y; g[2] *= k0.z; g[3] *= k0.w; g[4] *= k1.x; g[5] *= k1.y; g[6] *= k1.z; g[7] *= k1.w;
                } else {
                    for (int j = 0; j < 8; ++j) if (n + j < N) g[j] *= mask[m * N + n + j];
                }
            }
        }
        u32 pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat16 lo = __float2bfloat16(g[2 * j]), hi = __float2bfloat16(g[2 * j + 1]);
            tile[nn + 2 * j][mm] = __bfloat162float(lo);         // the bias gradient sums what the GEMMs see
            tile[nn + 2 * j + 1][mm] = __bfloat162float(hi);
            pk[j] = (u32)__bfloat16_as_ushort(lo) | ((u32)__bfloat16_as_ushort(hi) << 16);
        }
        if (m < M && n < ldyb) *reinterpret_cast<uint4*>(dYb + m * ldyb + n) = make_uint4(pk[0], pk[1], pk[2], pk[3]);      // ldyb % 64 == 0
    }
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
        const int item = tid + i * 256, nn = item >> 3, mm = (item & 7) * 8;
        if (n0 + nn < rows_t && m0 + mm < ldyt) {                // ldyt % 64 == 0: the 8 columns are inside the row
            u32 pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                pk[j] = (u32)__bfloat16_as_ushort(__float2bfloat16(tile[nn][mm + 2 * j])) | ((u32)__bfloat16_as_ushort(__float2bfloat16(tile[nn][mm + 2 * j + 1])) << 16);
            *reinterpret_cast<uint4*>(dYT + (long long)(n0 + nn) * ldyt + m0 + mm) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    if (tid < 64 && n0 + tid < N) {
        float sum = 0.f;
#pragma unroll 8
        for (int mm = 0; mm < 64; ++mm) sum += tile[tid][mm];
        if (sum != 0.f) atomicAdd(db + n0 + tid, sum);
    }
}
// training-mode dropout on a bf16 activation: the masks of dropout_kernel (same Philox words), kept in fp32 for the backward pass
__global__ void dropout_bf16_kernel(__nv_bfloat16* __restrict__ Y, float* __restrict__ mask, long long n, float rate, u32 k0, u32 k1, u32 tag) {
    pdl_launch_dependents(); pdl_wait();
    const u32 thr = (u32)fminf(rate * 4294967296.f, 4294967295.f);
    const float scale = 1.f / (1.f - rate);
    for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 * 4 < n; i4 += (long long)gridDim.x * blockDim.x) {
        const Philox4 u = philox4x32_10((u32)i4, (u32)(i4 >> 32), tag, 3u, k0, k1);
        const u32 uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i4 * 4 + j;
            if (i < n) { const float mk = (uu[j] >= thr) ? scale : 0.f; mask[i] = mk; Y[i] = __float2bfloat16(__bfloat162float(Y[i]) * mk); }
        }
    }
}

// Every bf16 weight copy of the network in ONE launch (grid.y = job), from fp32 W [K][N]:
//   mode 0: transposed, zero padded [d0 = Npad][d1 = Kpad] (the forward GEMMs' B operand); perm_C > 0 (layer 1): our K order is
//           (layer, tap) while W's rows are (tap, layer): k' = ci*T + t  <-  k = t*C + ci
//   mode 1: same orientation, zero padded [d0 = rows][d1 = Npad] (the "transposed weight" of the backward dX GEMM)
struct PrepJob { const float* W; __nv_bfloat16* out; int K, N, d0, d1, mode, perm_C; };
struct PrepJobs { PrepJob j[2 * (kMaxConv + kMaxDense + 2) + 1]; };
__global__ void prep_weights_kernel(const PrepJobs jobs) {
    pdl_launch_dependents(); pdl_wait();
    const PrepJob& jb = jobs.j[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= jb.d0 * jb.d1) return;
    const int r = i / jb.d1, q = i - r * jb.d1;
    float v = 0.f;
    if (jb.mode == 0) {                                         // r = n, q = k
        int ks = q;
        if (jb.perm_C > 0 && q < jb.K) { const int T = jb.K / jb.perm_C, ci = q / T, t = q - ci * T; ks = t * jb.perm_C + ci; }
        if (r < jb.N && q < jb.K) v = jb.W[(size_t)ks * jb.N + r];
    } else if (r < jb.K && q < jb.N) v = jb.W[(size_t)r * jb.N + q];      // r = k, q = n
    jb.out[i] = __float2bfloat16(v);
}

}  // namespace dq

struct dq_qnet_tc {                     // bf16 buffers of the tensor-core path, owned by the handle
    // tensor-core layer j = tensor j of the flat layout: conv1 .. conv_n, hidden dense .., Dense(num_actions)
    __nv_bfloat16* act[kMaxConv + kMaxDense + 2];      // act[j] = bf16 output of layer j (not for the last one)
    __nv_bfloat16* wt[kMaxConv + kMaxDense + 2];
    __nv_bfloat16* wb[kMaxConv + kMaxDense + 2];       // W itself in bf16, [K padded to the dX tile][N padded to 64]: the "transposed weight" of the backward dX GEMM
    int kpad[kMaxConv + kMaxDense + 2], npad[kMaxConv + kMaxDense + 2], bn[kMaxConv + kMaxDense + 2];
    // Dense(num_actions) + dueling head folded into one affine map (dq_qnet_fold_head; DQ_QNET_FOLD_HEAD=0 turns it off); the last
    // tensor-core layer then multiplies with fold_wt / fold_b and writes Q itself
    float* fold_w; float* fold_b; __nv_bfloat16* fold_wt; int folded;
    struct dq_qnet_tcb* bwd;            // scratch of the bf16 backward pass (dq_qnet_backward_tc), allocated on first use
};
static void tcb_free(struct dq_qnet_tcb* b);
static bool tc_fold_enabled() {
    static const bool on = [] { const char* e = getenv("DQ_QNET_FOLD_HEAD"); return !(e && e[0] == '0'); }();      // default on (measured: 0.127 -> 0.112 ms per 16 384 observations, same greedy agreement with fp32); =0 keeps the three layers apart
    return on;
}
static void tc_free(dq_qnet* h) {
    dq_qnet_tc* tc = (dq_qnet_tc*)h->tc;
    if (!tc) return;
    for (int i = 0; i < kMaxConv + kMaxDense + 2; ++i) { cudaFree(tc->act[i]); cudaFree(tc->wt[i]); cudaFree(tc->wb[i]); }
    cudaFree(tc->fold_w); cudaFree(tc->fold_b); cudaFree(tc->fold_wt);
    tcb_free(tc->bwd);
    delete tc;
    h->tc = nullptr;
}
static int tc_layers(const QCfg& c) { return c.n_conv + c.n_hidden + 1; }
static void tc_shape(const QCfg& c, int t, int& K, int& N, long long& rows) {
    if (t < c.n_conv) { K = c.conv[t].K; N = c.conv[t].filters; rows = c.conv[t].P; }
    else { K = c.fc_in[t - c.n_conv]; N = c.fc_out[t - c.n_conv]; rows = 1; }
}
static dq_qnet_tc* tc_of(dq_qnet* h) {
    if (h->tc) return (dq_qnet_tc*)h->tc;
    const QCfg& c = h->c;
    dq_qnet_tc* tc = new dq_qnet_tc();
    memset(tc, 0, sizeof(*tc));
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(h->device);
    cudaError_t err = cudaSuccess;
    const int n_tc = tc_layers(c);
    for (int j = 0; j < n_tc && err == cudaSuccess; ++j) {
        int K, N; long long rows;
        tc_shape(c, j, K, N, rows);
        tc->kpad[j] = (K + 63) / 64 * 64;
        tc->bn[j] = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
        tc->npad[j] = (N + tc->bn[j] - 1) / tc->bn[j] * tc->bn[j];
        err = cudaMalloc(&tc->wt[j], (size_t)tc->npad[j] * tc->kpad[j] * sizeof(__nv_bfloat16));
        if (err == cudaSuccess && j + 1 < n_tc) err = cudaMalloc(&tc->act[j], (size_t)h->max_batch * rows * N * sizeof(__nv_bfloat16));
        if (err == cudaSuccess && j > 0) {
            const int bn_dx = K <= 32 ? 32 : (K <= 64 ? 64 : 128);
            err = cudaMalloc(&tc->wb[j], (size_t)((K + bn_dx - 1) / bn_dx * bn_dx) * ((N + 63) / 64 * 64) * sizeof(__nv_bfloat16));
        }
    }
    if (err == cudaSuccess && tc_fold_enabled() && c.dueling && c.n_fc >= 2) {
        const int j = n_tc - 1, K = c.fc_in[c.n_fc - 2];
        err = cudaMalloc(&tc->fold_w, (size_t)K * c.A * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&tc->fold_b, (size_t)c.A * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&tc->fold_wt, (size_t)tc->npad[j] * tc->kpad[j] * sizeof(__nv_bfloat16));
        tc->folded = err == cudaSuccess;
    }
    cudaSetDevice(prev);
    h->tc = tc;
    if (err != cudaSuccess) { tc_free(h); return nullptr; }
    return tc;
}

static bool tc_pipe_enabled() {
    static const bool on = [] { const char* e = getenv("DQ_TC_PIPE"); return !(e && e[0] == '0'); }();
    return on;
}
// Opt a kernel in to all the dynamic shared memory the device allows beside its static part, once per (kernel, device): the attribute is a cap, not a
// reservation, and a driver call per launch is measurable when an update is 75 launches of a few microseconds.
template <auto Kernel>                  // one instantiation, and one set of flags, per kernel FUNCTION (a type parameter would merge all kernels of one signature)
static cudaError_t allow_big_smem() {
#ifdef DQ_EMU
    return cudaSuccess;
#else
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    cudaFuncAttributes fa;
    int optin = 0;
    cudaError_t e = cudaFuncGetAttributes(&fa, Kernel);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);      // the opt-in limit covers static + dynamic
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
#endif
}

template <int BN, int AMODE>
static int launch_tc(const TcArgs& a, int npad, cudaStream_t st) {
    if (AMODE == 0 && (a.Kpad >> 6) >= 2 && tc_pipe_enabled()) {
        dim3 grid((unsigned)((a.M + 127) / 128), npad / BN);
        // few tiles (at most ~2 per SM) and a long K: four stages keep three chunks in flight per CTA; otherwise two stages
        // and more CTAs per SM
        const bool deep = (long long)grid.x * grid.y <= 2 * 148 && (a.Kpad >> 6) >= 4;
        const size_t smem = (deep ? 4 : 2) * (16384 + (size_t)BN * 128) + 1024;
        if (deep) {
            QCUDA((allow_big_smem<tc_gemm_pipe_kernel<BN, 4>>()));
            DQ_LAUNCH_PDL((tc_gemm_pipe_kernel<BN, 4>), grid, 128, smem, st, a);
        } else {
            QCUDA((allow_big_smem<tc_gemm_pipe_kernel<BN, 2>>()));
            DQ_LAUNCH_PDL((tc_gemm_pipe_kernel<BN, 2>), grid, 128, smem, st, a);
        }
        count_launch();
        return DQ_OK;
    }
    const size_t smem = (size_t)(128 + BN) * (a.Kpad >> 6) * 128 + 1024;
    if (smem > 227 * 1024) return qfail(DQ_EINVAL, "tensor-core tile does not fit shared memory");
    QCUDA((allow_big_smem<tc_gemm_kernel<BN, AMODE>>()));
    dim3 grid((unsigned)((a.M + 127) / 128), npad / BN);
    DQ_LAUNCH_PDL((tc_gemm_kernel<BN, AMODE>), grid, 128, smem, st, a);
    count_launch();
    return DQ_OK;
}
template <int AMODE>
static int launch_tc_bn(const TcArgs& a, int bn, int npad, cudaStream_t st) {
    switch (bn) {
        case 32: return launch_tc<32, AMODE>(a, npad, st);
        case 64: return launch_tc<64, AMODE>(a, npad, st);
        default: return launch_tc<128, AMODE>(a, npad, st);
    }
}

// dense layers through the TMA-fed kernel: A = [M][K] bf16 with leading dimension lda, all K blocks resident (Kpad <= 512)
template <int BN>
static int launch_tc_dense_tma(const TcArgs& a, long long lda, long long cols, int npad, cudaStream_t st) {
    const int KB = a.Kpad >> 6;
    const size_t smem = (size_t)KB * (16384 + (size_t)BN * 128) + 1024;
    Tmap2D ta, tw;
    if (!make_tmap_2d(&ta, a.X, a.M, lda, 128, cols) || !make_tmap_2d(&tw, a.Wt, npad, a.Kpad, BN)) return qfail(DQ_ECUDA, "cuTensorMapEncodeTiled failed");
    QCUDA((allow_big_smem<tc_gemm_tma_kernel<BN>>()));
    DQ_LAUNCH_PDL((tc_gemm_tma_kernel<BN>), dim3((unsigned)((a.M + 127) / 128), npad / BN), 128, smem, st, ta, tw, a);
    count_launch();
    return DQ_OK;
}
static bool tc_dense_tma_ok(const TcArgs& a, int bn, long long lda) {
    static const bool on = [] { const char* e = getenv("DQ_TC_DENSE_TMA"); return !(e && e[0] == '0'); }();
    const int KB = a.Kpad >> 6;
    return on && KB >= 1 && KB <= 8 && (size_t)KB * (16384 + (size_t)bn * 128) + 1024 <= 227 * 1024 && (lda % 8) == 0;
}
// lda = leading dimension of A, cols = its readable columns (anything beyond reads as zero)
static int launch_tc_dense(const TcArgs& a, long long lda, long long cols, int bn, int npad, cudaStream_t st) {
    switch (bn) {
        case 32: return launch_tc_dense_tma<32>(a, lda, cols, npad, st);
        case 64: return launch_tc_dense_tma<64>(a, lda, cols, npad, st);
        default: return launch_tc_dense_tma<128>(a, lda, cols, npad, st);
    }
}

static int tc_check(const dq_qnet* h) {
    const QCfg& c = h->c;
    if (c.n_hidden < 1) return qfail(DQ_EINVAL, "tensor-core path needs >= 1 hidden dense layer");
    for (int l = 0; l < c.n_conv; ++l) if (c.conv[l].filters % 8) return qfail(DQ_EINVAL, "conv filter counts must be multiples of 8 for the tensor-core path");
    for (int i = 0; i < c.n_hidden; ++i) if (c.fc_out[i] % 8) return qfail(DQ_EINVAL, "hidden dense widths must be multiples of 8 for the tensor-core path");
    if (c.C * c.PW > 36 || (128 / c.conv[0].P + 2) > 16) return qfail(DQ_EINVAL, "first layer too large for the tensor-core path");
    for (int l = 1; l < c.n_conv; ++l)
        if (c.conv[l].ksz * c.conv[l].ksz * c.conv[l - 1].filters > 1024) return qfail(DQ_EINVAL, "conv layer K > 1024 is not supported on the tensor-core path");
    for (int i = 0; i < c.n_hidden + 1 && i < c.n_fc; ++i)
        if (c.fc_in[i] > 1024) return qfail(DQ_EINVAL, "dense layer K > 1024 is not supported on the tensor-core path");
    return DQ_OK;
}

// bf16 copies of the weights (transposed, zero-padded, layer 1 with K in (layer, tap) order).  Call after every change of `params`.
extern "C" int dq_qnet_prepare_tc(dq_qnet* h, const float* params, dq_stream stream) {
    if (!h || !params) return qfail(DQ_EINVAL, "NULL argument");
    int rc = tc_check(h);
    if (rc) return rc;
    const QCfg& c = h->c;
    dq_qnet_tc* tc = tc_of(h);
    if (!tc) return qfail(DQ_ECUDA, "allocating the bf16 buffers failed");
    PrepJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    int nj = 0, max_total = 0;
    auto add = [&](const float* W, __nv_bfloat16* out, int K, int N, int d0, int d1, int mode, int perm_C) {
        jobs.j[nj++] = PrepJob{W, out, K, N, d0, d1, mode, perm_C};
        max_total = std::max(max_total, d0 * d1);
    };
    for (int j = 0; j < tc_layers(c); ++j) {
        int K, N; long long rows;
        tc_shape(c, j, K, N, rows);
        add(params + c.w_off[j], tc->wt[j], K, N, tc->npad[j], tc->kpad[j], 0, j == 0 ? c.C : 0);
        if (j > 0) { const int bn_dx = K <= 32 ? 32 : (K <= 64 ? 64 : 128); add(params + c.w_off[j], tc->wb[j], K, N, (K + bn_dx - 1) / bn_dx * bn_dx, (N + 63) / 64 * 64, 1, 0); }
    }
    if (tc->folded) {
        const int j = tc_layers(c) - 1;
        int K, N; long long rows;
        tc_shape(c, j, K, N, rows);
        rc = dq_qnet_fold_head(h, params, tc->fold_w, tc->fold_b, stream);
        if (rc) return rc;
        add(tc->fold_w, tc->fold_wt, K, N, tc->npad[j], tc->kpad[j], 0, 0);
    }
    DQ_LAUNCH_PDL(prep_weights_kernel, dim3((max_total + 255) / 256, nj), 256, 0, (cudaStream_t)stream, jobs);
    count_launch();
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

// Q values through the bf16 tcgen05 path (inference / acting only; training stays fp32).  Uses the weights staged by
// the last dq_qnet_prepare_tc.  Shapes outside the path's coverage return DQ_EINVAL (callers choose the fp32 path).
static int tc_forward(dq_qnet* h, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                      float* q_out, int train, uint64_t dropout_seed, dq_stream stream) {
    if (!h || !params || !packed || !q_out) return qfail(DQ_EINVAL, "NULL argument");
    if (batch < 1 || batch > h->max_batch) return qfail(DQ_EINVAL, "batch exceeds max_batch of the handle");
    if (!h->tc) return qfail(DQ_ESTATE, "call dq_qnet_prepare_tc first");
    const QCfg& c = h->c;
    cudaStream_t st = (cudaStream_t)stream;
    dq_qnet_tc* tc = (dq_qnet_tc*)h->tc;
    const int n_tc = tc_layers(c);
    const bool folded = tc->folded && !train;              // training keeps Dense(num_actions) and the dueling layer apart (their gradients differ)
    for (int j = 0; j < n_tc; ++j) {
        TcArgs a;
        memset(&a, 0, sizeof(a));
        a.Wt = tc->wt[j]; a.bias = params + c.b_off[j]; a.Kpad = tc->kpad[j];
        const bool last = (j == n_tc - 1);
        if (j < c.n_conv) { const ConvL& L = c.conv[j]; a.g = conv_patch(L); a.M = batch * L.P; a.N = L.filters; a.K = L.K; }
        else { const int i = j - c.n_conv; a.g = dense_patch(c.fc_in[i]); a.M = batch; a.N = c.fc_out[i]; a.K = c.fc_in[i]; }
        a.relu = last ? 0 : 1; a.out_bf16 = last ? 0 : 1; a.ldy = a.N;
        a.Y = last ? (void*)h->act_fc[c.n_hidden] : (void*)tc->act[j];
        if (last && folded) { a.Wt = tc->fold_wt; a.bias = tc->fold_b; a.Y = q_out; }      // N = num_actions: the rows are Q itself
        int rc;
        if (j == 0) {
            a.packed = (const u64*)packed; a.pstride = stride; a.C = c.C; a.PW = c.PW; a.H = c.H; a.T = c.conv[0].ksz * c.conv[0].ksz;
            rc = launch_tc_bn<1>(a, tc->bn[j], tc->npad[j], st);
        } else {
            a.X = tc->act[j - 1];
            if (j >= c.n_conv && tc_dense_tma_ok(a, tc->bn[j], a.K)) rc = launch_tc_dense(a, a.K, a.K, tc->bn[j], tc->npad[j], st);      // dense layer: its input rows are contiguous
            else rc = launch_tc_bn<0>(a, tc->bn[j], tc->npad[j], st);
        }
        if (rc) return rc;
        if (train && j >= c.n_conv && !last && c.drop[j - c.n_conv] > 0.f) {     // Dropout after a hidden dense layer (FL:366-370)
            const int i = j - c.n_conv;
            const long long n = batch * c.fc_out[i];
            DQ_LAUNCH_PDL(dropout_bf16_kernel, grid_for((n + 3) / 4, 256), 256, 0, st, tc->act[j], h->mask_fc[i], n, c.drop[i], (u32)dropout_seed, (u32)(dropout_seed >> 32), (u32)i);
            count_launch();
        }
    }
    h->last_train = train;
    if (folded) { QCUDA(cudaGetLastError()); return DQ_OK; }
    // dueling head (tiny) in fp32 on the SIMT path
    const float* xf = h->act_fc[c.n_hidden];
    if (c.dueling) {
        const int i = c.n_fc - 1, K = c.fc_in[i], N = c.fc_out[i], t = c.n_conv + i;
        const int N4 = (N + 3) & ~3;
        const size_t smem = (size_t)(K * N4 + N4 + kHeadSamples * K) * sizeof(float);
        const unsigned hgrid = (unsigned)((batch + kHeadSamples - 1) / kHeadSamples);
        if (N <= 64 && smem <= 48 * 1024) {
            DQ_LAUNCH_PDL((head_dueling_kernel<4>), hgrid, 128, smem, st, xf, params + c.w_off[t], params + c.b_off[t], q_out, batch, K, N, c.A, 1);
            count_launch();
        } else if (N <= 128 && smem <= 200 * 1024) {
            QCUDA(cudaFuncSetAttribute(head_dueling_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DQ_LAUNCH_PDL((head_dueling_kernel<8>), hgrid, 128, smem, st, xf, params + c.w_off[t], params + c.b_off[t], q_out, batch, K, N, c.A, 1);
            count_launch();
        } else {
            launch_gemm_fwd(xf, dense_patch(K), params + c.w_off[t], params + c.b_off[t], h->act_fc[i], batch, N, K, 0, st);
            dueling_fwd_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, st>>>(h->act_fc[i], q_out, batch, c.A);
            count_launch();
        }
    } else {
        QCUDA(cudaMemcpyAsync(q_out, xf, (size_t)batch * c.A * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    QCUDA(cudaGetLastError());
    return DQ_OK;
}
extern "C" int dq_qnet_forward_tc(dq_qnet* h, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                                  float* q_out, dq_stream stream) {
    return tc_forward(h, params, packed, stride, batch, q_out, 0, 0, stream);
}
// The forward pass of a bf16 update: same tensor-core layers, Dense(num_actions) and the dueling layer kept apart, dropout applied to
// the bf16 activations (masks kept for dq_qnet_backward_tc).
extern "C" int dq_qnet_forward_tc_train(dq_qnet* h, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                                        float* q_out, uint64_t dropout_seed, dq_stream stream) {
    return tc_forward(h, params, packed, stride, batch, q_out, 1, dropout_seed, stream);
}

// ---- bf16 backward --------------------------------------------------------------------------------------------------------------
struct dq_qnet_tcb {                    // scratch of dq_qnet_backward_tc, sized for `cap` samples, grown on demand
    long long cap;
    __nv_bfloat16 *at, *dyb, *dcol;
    __nv_bfloat16* dyt[kMaxConv + kMaxDense + 2];       // one transposed gradient per layer: layer j's weight-gradient GEMM reads it on the
                                                        // side stream while the caller's stream is already staging layer j-1
    cudaStream_t side;                                  // the weight-gradient chain (A^T, dW GEMM) of every layer, see dq_qnet_backward_tc
    cudaEvent_t ev_layer[kMaxConv + kMaxDense + 2], ev_start, ev_done;
};
struct TcbGeom { int K, N; long long rows, M, Mpad; int rowsA, bn_dw, rows_t, ldyb, bn_dx, npad_dx; };
static int bn_for(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : 128); }
static TcbGeom tcb_geom(const QCfg& c, int j, long long batch) {
    TcbGeom g;
    tc_shape(c, j, g.K, g.N, g.rows);
    g.M = batch * g.rows; g.Mpad = (g.M + 63) / 64 * 64;
    g.rowsA = (g.K + 127) / 128 * 128;
    g.bn_dw = bn_for(g.N); g.rows_t = (g.N + g.bn_dw - 1) / g.bn_dw * g.bn_dw;
    g.ldyb = (g.N + 63) / 64 * 64;
    g.bn_dx = bn_for(g.K); g.npad_dx = (g.K + g.bn_dx - 1) / g.bn_dx * g.bn_dx;
    return g;
}
static void tcb_free(dq_qnet_tcb* b) {
    if (!b) return;
    cudaFree(b->at); cudaFree(b->dyb); cudaFree(b->dcol);
    for (int j = 0; j < kMaxConv + kMaxDense + 2; ++j) { cudaFree(b->dyt[j]); if (b->ev_layer[j]) cudaEventDestroy(b->ev_layer[j]); }
    if (b->ev_start) cudaEventDestroy(b->ev_start);
    if (b->ev_done) cudaEventDestroy(b->ev_done);
    if (b->side) cudaStreamDestroy(b->side);
    delete b;
}
static dq_qnet_tcb* tcb_of(dq_qnet* h, long long batch) {
    dq_qnet_tc* tc = (dq_qnet_tc*)h->tc;
    if (tc->bwd && tc->bwd->cap >= batch) return tc->bwd;
    tcb_free(tc->bwd);
    tc->bwd = nullptr;
    const QCfg& c = h->c;
    size_t n_at = 0, n_dyb = 0, n_dcol = 0;
    dq_qnet_tcb* b = new dq_qnet_tcb();
    memset(b, 0, sizeof(*b));
    cudaError_t err = cudaSuccess;
    for (int j = 0; j < tc_layers(c); ++j) {
        const TcbGeom g = tcb_geom(c, j, batch);
        n_at = std::max<size_t>(n_at, (size_t)g.rowsA * (size_t)g.Mpad);
        n_dyb = std::max<size_t>(n_dyb, (size_t)g.Mpad * (size_t)g.ldyb);
        if (j > 0) n_dcol = std::max<size_t>(n_dcol, (size_t)g.M * (size_t)g.K);
        if (err == cudaSuccess) err = cudaMalloc(&b->dyt[j], (size_t)g.rows_t * (size_t)g.Mpad * 2);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&b->ev_layer[j], cudaEventDisableTiming);
    }
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&b->side, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&b->ev_start, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaMalloc(&b->at, n_at * 2);
    if (err == cudaSuccess) err = cudaMalloc(&b->dyb, n_dyb * 2);
    if (err == cudaSuccess) err = cudaMalloc(&b->dcol, std::max<size_t>(n_dcol, 8) * 2);
    if (err != cudaSuccess) { tcb_free(b); return nullptr; }
    b->cap = batch;
    tc->bwd = b;
    return b;
}
template <int BN>
static int launch_tc_dw(const __nv_bfloat16* At, long long lda, const __nv_bfloat16* Bt, long long ldb, float* D, int ldd, int R, int N,
                        int rowsA, int rows_t, cudaStream_t st) {
    constexpr int S = 4;
    const size_t smem = (size_t)S * (16384 + (size_t)BN * 128) + 1024;
    QCUDA((allow_big_smem<tc_dw_tma_kernel<BN, S>>()));
    const int kb_total = (int)(lda / 64);
    const int tiles = (rowsA / 128) * (rows_t / BN);
    // the contraction (batch x positions) is cut into slices over grid.z: about two CTAs per SM, at least 16 chunks of 64 per CTA (measured at batch 4096: 4 -> 556 us per update, 8 -> 536, 16 -> 529, 32 -> 554, 64 -> 623)
    static const int min_chunks = [] { const char* e = getenv("DQ_TC_DW_MINCHUNKS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 16; }();
    int z = std::max(1, std::min(kb_total / min_chunks, (2 * 148 + tiles - 1) / tiles));
    const int per = (kb_total + z - 1) / z;
    z = (kb_total + per - 1) / per;
    Tmap2D ta, tb;
    if (!make_tmap_2d(&ta, At, rowsA, lda, 128) || !make_tmap_2d(&tb, Bt, rows_t, ldb, BN)) return qfail(DQ_ECUDA, "cuTensorMapEncodeTiled failed");
    DQ_LAUNCH_PDL((tc_dw_tma_kernel<BN, S>), dim3(rowsA / 128, rows_t / BN, z), 128, smem, st, ta, tb, D, ldd, R, N, kb_total, per);
    count_launch();
    return DQ_OK;
}

// Gradients of sum_b sum_a dq[b][a]*Q[b][a] for the batch of the LAST dq_qnet_forward_tc_train call, through the tensor cores:
// per layer two bf16 GEMMs with fp32 accumulation (dW over the batch, dX over the output channels); the dueling layer (51 x 52)
// stays on the fp32 kernels.  grads (flat fp32, n_params) is overwritten.
extern "C" int dq_qnet_backward_tc(dq_qnet* h, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                                   const float* dq, float* grads, dq_stream stream) {
    if (!h || !params || !packed || !dq || !grads) return qfail(DQ_EINVAL, "NULL argument");
    if (batch < 1 || batch > h->max_batch) return qfail(DQ_EINVAL, "batch exceeds max_batch of the handle");
    if (!h->tc) return qfail(DQ_ESTATE, "call dq_qnet_prepare_tc and dq_qnet_forward_tc_train first");
    const QCfg& c = h->c;
    cudaStream_t st = (cudaStream_t)stream;
    dq_qnet_tc* tc = (dq_qnet_tc*)h->tc;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(h->device);
    dq_qnet_tcb* sb = tcb_of(h, batch);
    cudaSetDevice(prev);
    if (!sb) return qfail(DQ_ECUDA, "allocating the backward scratch failed");
    const int n_tc = tc_layers(c);
    QCUDA(cudaMemsetAsync(grads, 0, (size_t)c.n_params * sizeof(float), st));
    // Two chains per layer share only the staged gradient: the COLUMN-gradient chain (dX GEMM -> the staging of the layer below) is what
    // the next layer waits for and stays on the caller's stream; the WEIGHT-gradient chain (A^T, dW GEMM) of every layer runs on a side
    // stream, released layer by layer by an event and joined at the end (DQ_TC_BWD_STREAMS=0: everything on the caller's stream).
    static const bool two_streams = [] { const char* e = getenv("DQ_TC_BWD_STREAMS"); return !(e && e[0] == '0'); }();
    cudaStream_t sw = two_streams ? sb->side : st;
    if (two_streams) {
        QCUDA(cudaEventRecord(sb->ev_start, st));                   // the gradient buffer is zero, the previous user of the scratch is done
        QCUDA(cudaStreamWaitEvent(sb->side, sb->ev_start, 0));
    }
    // head: G = gradient at the output of Dense(num_actions), fp32 [batch][A]
    const float* G32 = dq;
    if (c.dueling) {
        const int i = c.n_fc - 1, t = c.n_conv + i, K = c.fc_in[i], N = c.fc_out[i];
        const size_t smem = (size_t)(K * N + kHeadBwdSamples * N + kHeadBwdSamples * K) * sizeof(float);
        if (N <= 128 && smem <= 48 * 1024) {               // the dueling layer of every shipped configuration: one fused launch
            DQ_LAUNCH_PDL(head_dueling_bwd_kernel, (unsigned)((batch + kHeadBwdSamples - 1) / kHeadBwdSamples), 128, smem, st, dq, h->act_fc[i - 1], params + c.w_off[t],
                          (long long)batch, K, N, c.A, h->dact_fc[i - 1], grads + c.w_off[t], grads + c.b_off[t]);
            count_launch();
        } else {
            float* dY = h->dact_fc[i];
            DQ_LAUNCH_PDL(dueling_bwd_kernel, (unsigned)((batch + 127) / 128), 128, 0, st, dq, dY, batch, c.A);
            const long long chunk = dw_chunk(batch, K, N);
            dim3 gw((K + TB - 1) / TB, (N + TB - 1) / TB, (unsigned)((batch + chunk - 1) / chunk));
            DQ_LAUNCH_PDL(gemm_dw_kernel, gw, 256, 0, st, h->act_fc[i - 1], dense_patch(K), dY, grads + c.w_off[t], batch, N, K, chunk);
            DQ_LAUNCH_PDL(colsum_kernel, dim3((N + 31) / 32, (unsigned)std::min<long long>(64, (batch + 7) / 8)), 256, 0, st, dY, grads + c.b_off[t], batch, N);
            dim3 gx((unsigned)((batch + TB - 1) / TB), (K + TB - 1) / TB);
            DQ_LAUNCH_PDL((gemm_dx_kernel<16>), gx, 256, 0, st, dY, params + c.w_off[t], h->dact_fc[i - 1], dense_patch(K), batch, N, K, 0);
            count_launch(); count_launch(); count_launch(); count_launch();
        }
        G32 = h->dact_fc[i - 1];
    }
    for (int j = n_tc - 1; j >= 0; --j) {
        const TcbGeom g = tcb_geom(c, j, batch);
        // 1. the masked gradient at this layer's output, in the two layouts the GEMMs read, and the bias gradient
        DySrc src;
        memset(&src, 0, sizeof(src));
        const __nv_bfloat16* act = nullptr;
        const float* mask = nullptr;
        if (j == n_tc - 1) { src.p = G32; src.mode = 0; }
        else {
            act = tc->act[j];                                                       // ReLU: pass where the (bf16) output is positive
            if (j >= c.n_conv && h->last_train && c.drop[j - c.n_conv] > 0.f) mask = h->mask_fc[j - c.n_conv];
            src.p = sb->dcol;
            if (j + 1 < c.n_conv) { src.mode = 2; src.up = conv_patch(c.conv[j + 1]); src.Kup = c.conv[j + 1].K; }
            else src.mode = 1;                                                      // a dense layer above: its column gradient is this layer's map
        }
        const int ncols = std::max(g.ldyb, g.rows_t);
        DQ_LAUNCH_PDL(prep_dy_kernel, dim3((unsigned)(g.Mpad / 64), (ncols + 63) / 64), 256, 0, st, src, act, mask, g.M, g.N, sb->dyb, g.ldyb, sb->dyt[j], g.Mpad, g.rows_t,
                                                                                      grads + c.b_off[j]);
        count_launch();
        if (two_streams) {
            QCUDA(cudaEventRecord(sb->ev_layer[j], st));
            QCUDA(cudaStreamWaitEvent(sb->side, sb->ev_layer[j], 0));
        }
        // 2. A^T, then dW = A^T x dY over the batch
        if (j == 0) {
            const ConvL& L = c.conv[0];
            DQ_LAUNCH_PDL(im2colT_bits_kernel, dim3((unsigned)((g.Mpad + 255) / 256), 1), 256, 0, sw, (const u64*)packed, stride, L, c.C, c.PW, c.H, g.M, sb->at, g.Mpad, g.rowsA);
        } else {
            const Patch pg = j < c.n_conv ? conv_patch(c.conv[j]) : dense_patch(g.K);
            DQ_LAUNCH_PDL(im2colT_kernel, dim3((unsigned)(g.Mpad / 64), g.rowsA / 64), 256, 0, sw, tc->act[j - 1], pg, g.M, g.K, sb->at, g.Mpad);
        }
        count_launch();
        int rc;
        switch (g.bn_dw) {
            case 32: rc = launch_tc_dw<32>(sb->at, g.Mpad, sb->dyt[j], g.Mpad, grads + c.w_off[j], g.N, g.K, g.N, g.rowsA, g.rows_t, sw); break;
            case 64: rc = launch_tc_dw<64>(sb->at, g.Mpad, sb->dyt[j], g.Mpad, grads + c.w_off[j], g.N, g.K, g.N, g.rowsA, g.rows_t, sw); break;
            default: rc = launch_tc_dw<128>(sb->at, g.Mpad, sb->dyt[j], g.Mpad, grads + c.w_off[j], g.N, g.K, g.N, g.rowsA, g.rows_t, sw); break;
        }
        if (rc) return rc;
        if (j == 0) break;
        // 3. column gradient dCol[M][K] = dY x W^T: the forward GEMM kernel with W's bf16 copy [K][N] as its "transposed weight"
        {
            TcArgs a;
            memset(&a, 0, sizeof(a));
            a.X = sb->dyb; a.g = dense_patch(g.ldyb); a.Wt = tc->wb[j]; a.bias = nullptr; a.Y = sb->dcol; a.ldy = g.K; a.out_bf16 = 1; a.relu = 0;
            a.M = g.M; a.N = g.K; a.K = g.N; a.Kpad = g.ldyb;
            // (the staged gradient is a plain [M][ldyb] matrix whatever the layer: the dense-layer kernel, all K blocks in flight)
            rc = tc_dense_tma_ok(a, g.bn_dx, g.ldyb) ? launch_tc_dense(a, g.ldyb, g.ldyb, g.bn_dx, g.npad_dx, st) : launch_tc_bn<0>(a, g.bn_dx, g.npad_dx, st);
            if (rc) return rc;
        }
    }
    if (two_streams) {
        QCUDA(cudaEventRecord(sb->ev_done, sb->side));
        QCUDA(cudaStreamWaitEvent(st, sb->ev_done, 0));             // the caller's stream sees the whole gradient
    }
    QCUDA(cudaGetLastError());
    return DQ_OK;
}

// bf16 activation j of the last dq_qnet_forward_tc call (0 = conv1 output, j+1 = output of tensor-core layer j); tests
extern "C" int dq_qnet_tc_activation(dq_qnet* h, int index, void** dev_ptr, int64_t* per_sample) {
    if (!h || !dev_ptr || !h->tc) return qfail(DQ_EINVAL, "no tensor-core forward has run on this handle");
    const QCfg& c = h->c;
    const int n_tc = c.n_conv + c.n_hidden;                 // layers with a bf16 output
    if (index < 0 || index >= n_tc) return qfail(DQ_EINVAL, "no such activation");
    *dev_ptr = ((dq_qnet_tc*)h->tc)->act[index];
    if (per_sample) {
        if (index < c.n_conv) *per_sample = (int64_t)c.conv[index].P * c.conv[index].filters;
        else *per_sample = c.fc_out[index - c.n_conv];
    }
    return DQ_OK;
}
