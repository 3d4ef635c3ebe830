// dq_env.cu -- sm_100a environment-step kernel + C ABI (include/dq_decoding.h).
//
// One CTA advances 32 independent lattices.  Per-lattice state is bit-packed into
// STATE_WORDS uint64 rows of a [row][lattice] matrix in HBM; a CTA's tile
// is one 256-byte segment per row, staged through shared memory with 1-D TMA bulk copies
// (cp.async.bulk + mbarrier in, cp.async.bulk.bulk_group out).  Phases:
//   A  warp 0, lane = lattice: apply the action to the Pauli frame, true syndrome by shifted
//      XORs, homology label, referee table lookup, reward / done, heavy(identity|repeat) flag
//   B  warp per heavy lattice: a volume attempt = R rounds of one Philox4x32-10 block per lane;
//      threshold compares become __ballot_sync words which ARE the per-slice error /
//      measurement-flip bit streams; lanes < volume_depth each rebuild one slice, a warp
//      prefix-XOR gives the frame after every slice; repeat until the summed volume is non-trivial
//   C  thread per (lattice, layer): legal-move mask, and the layer's (2d+1)^2-cell bitmap OR-ed
//      into one contiguous bit stream for the CTA
//   D  all threads: 16 stream bits -> 16 observation bytes, one aligned 128-bit store each
//      (the CTA's 32 observations are one contiguous, 16-byte aligned span of HBM)
//
// Replaces (reference paths relative to example_notebooks/): Environments.py:99-115 (reset),
// :118-204 (step), :206-235, :238-314 and the Function_Library.py helpers they call.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <string>

#include "../../include/dq_decoding.h"
#include "dq_lattice.cuh"

namespace dq {

constexpr int kEnvsPerCta = 32;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxVd = 8;
constexpr int kMaxChunks = 4 * 7 + 2;     // d=7, vd=8: 776 items -> 7 rounds -> 28 ballot words (+2 pad)

// rows of the packed state matrix
constexpr int ROW_XB = 0, ROW_ZB = 1, ROW_META = 2, ROW_ACT = 3, ROW_SYN = 6;
// meta word: [lifetime:32][attempt counter:31][done:1]
DQ_HD u64 meta_pack(u32 life, u32 attempts, u32 done) { return (u64)life | ((u64)(attempts & 0x7FFFFFFFu) << 32) | ((u64)done << 63); }

struct EnvParams {
    int d, model, use_y, vd, layers, A, W;      // W = mask words
    int n, npad;                                // lattices, padded to 32
    int rounds;                                 // Philox rounds per volume attempt (B = 32*rounds)
    int obs_bits;                               // C*H*H
    u32 T, T1, T2, Tm;                          // thresholds (RNG contract)
    u32 k0, k1;                                 // Philox key
    u32 env_id_base;
    int ref_mode;
    const uint8_t* lut_a;
    const uint8_t* lut_b;
    u64* state;                                 // [STATE_WORDS][npad]
};

// ---------------------------------------------------------------- PTX wrappers (TMA bulk copy + mbarrier)
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ int lut2(const uint8_t* lut, u32 idx) { return (__ldg(lut + (idx >> 2)) >> ((idx & 3) * 2)) & 3; }

template <int D>
__device__ __forceinline__ int referee_class(const EnvParams& p, u64 syn) {
    if (p.ref_mode == DQ_REFEREE_JOINT) return lut2(p.lut_a, (u32)stabs_grid_to_compact<D>(syn));
    int c = lut2(p.lut_a, stabs_grid_to_type_index<D, 1>(syn)) & 1;
    if (p.model == DQ_MODEL_DP && p.lut_b) c |= (lut2(p.lut_b, stabs_grid_to_type_index<D, 0>(syn)) & 1) << 1;
    return c;
}

// OR a 64-bit value into a little-endian u32 bit stream in shared memory at a run-time bit offset
__device__ __forceinline__ void stream_or64(u32* s, int off, u64 v) {
    int w = off >> 5, sh = off & 31;
    u32 v0 = (u32)v, v1 = (u32)(v >> 32);
    u32 x0 = v0 << sh, x1 = __funnelshift_l(v0, v1, sh), x2 = __funnelshift_l(v1, 0u, sh);
    if (x0) atomicOr(s + w, x0);
    if (x1) atomicOr(s + w + 1, x1);
    if (x2) atomicOr(s + w + 2, x2);
}

struct Smem {
    u64 st[ROW_SYN + kMaxVd][kEnvsPerCta];   // state tile, [row][lattice]
    u32 chA[kWarps][kMaxChunks + 2];         // ballot streams: x-flip / measurement-flip bits
    u32 chB[kWarps][kMaxChunks + 2];         // ballot streams: z-flip bits (DP)
    int32_t life_out[kEnvsPerCta];
    uint8_t task[kEnvsPerCta];               // lattices needing volume generation
    uint8_t task_flags[kEnvsPerCta];         // bit0: heavy volume, bit1: reset afterwards
    int ntask;
    alignas(8) u64 bar;
};

// One volume (Environments.py:158-176 / :216-235) for lattice `slot`, executed by a full warp.
// Updates xb, zb (frame), life, attempts; leaves the vd faulty slices in sm.st[ROW_SYN + j][slot].
template <int D>
__device__ __forceinline__ void generate_volume(const EnvParams& p, Smem& sm, int warp, int lane, int slot,
                                                u32 env_id, u64& xb, u64& zb, u32& life, u32& attempts) {
    typedef Lat<D> L;
    const int vd = p.vd, R = p.rounds, nq_items = vd * L::NQ;
    const u32 TA = (p.model == DQ_MODEL_DP) ? p.T2 : p.T;
    u32* chA = sm.chA[warp];
    u32* chB = sm.chB[warp];
    bool nontrivial;
    u64 f = 0;
    do {
        for (int r = 0; r < R; ++r) {
            const int blk = r * 32 + lane;
            Philox4 u = philox4x32_10(env_id, attempts, (u32)blk, 0u, p.k0, p.k1);
            const u32 uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int item = (w * R + r) * 32 + lane;            // = w*B + blk
                const bool isq = item < nq_items;
                const bool a = uu[w] < (isq ? TA : p.Tm);
                const bool b = isq && (p.model == DQ_MODEL_DP) && uu[w] >= p.T1 && uu[w] < p.T;
                const u32 ba = __ballot_sync(0xffffffffu, a), bb = __ballot_sync(0xffffffffu, b);
                if (lane == 0) { chA[w * R + r] = ba; chB[w * R + r] = bb; }
            }
        }
        if (lane < 2) { chA[4 * R + lane] = 0; chB[4 * R + lane] = 0; }
        __syncwarp();
        u64 ex = 0, ez = 0, m = 0;
        if (lane < vd) {
            ex = qubits_compact_to_grid<D>(extract_bits(chA, lane * L::NQ, L::NQ));
            if (p.model == DQ_MODEL_DP) ez = qubits_compact_to_grid<D>(extract_bits(chB, lane * L::NQ, L::NQ));
            m = stabs_compact_to_grid<D>(extract_bits(chA, nq_items + lane * L::NS, L::NS));
        }
        __syncwarp();
        // inclusive prefix XOR over slices: frame delta after slice `lane`
#pragma unroll
        for (int off = 1; off < kMaxVd; off <<= 1) {
            u64 tx = __shfl_up_sync(0xffffffffu, ex, off), tz = __shfl_up_sync(0xffffffffu, ez, off);
            if (lane >= off) { ex ^= tx; ez ^= tz; }
        }
        const u64 fx = xb ^ ex, fz = zb ^ ez;
        f = true_syndrome<D>(fx, fz) ^ m;
        nontrivial = __ballot_sync(0xffffffffu, lane < vd && f != 0) != 0;
        xb = __shfl_sync(0xffffffffu, fx, vd - 1);
        zb = __shfl_sync(0xffffffffu, fz, vd - 1);
        life += (u32)vd;
        attempts += 1;
    } while (!nontrivial);
    if (lane < vd) sm.st[ROW_SYN + lane][slot] = f;
}

template <int D, bool RESET>
__global__ void __launch_bounds__(kThreads)
env_step_kernel(const EnvParams p, const int32_t* __restrict__ actions, uint8_t* __restrict__ obs,
                float* __restrict__ reward, uint8_t* __restrict__ done_out, int32_t* __restrict__ lifetime,
                u64* __restrict__ legal, int auto_reset) {
    typedef Lat<D> L;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    u32* bits = reinterpret_cast<u32*>(smem_raw + ((sizeof(Smem) + 15) & ~size_t(15)));   // obs_bits u32 words (+2 pad)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int env0 = blockIdx.x * kEnvsPerCta;
    const int nrows_used = ROW_SYN + p.vd;
    const int C = p.vd + p.layers;

    // ---- stage the state tile: one 256 B TMA bulk copy per used row
    if (tid == 0) { mbar_init(&sm.bar, 1); sm.ntask = 0; }
    __syncthreads();
    if (tid == 0) {
        const int nrows = 3 + p.layers + p.vd;
        mbar_expect_tx(&sm.bar, (u32)(nrows * kEnvsPerCta * 8));
        for (int r = 0; r < nrows_used; ++r) {
            if (r >= ROW_ACT + p.layers && r < ROW_SYN) continue;
            bulk_load(&sm.st[r][0], p.state + (size_t)r * p.npad + env0, kEnvsPerCta * 8, &sm.bar);
        }
    }
    for (int i = tid; i < p.obs_bits + 2; i += kThreads) bits[i] = 0;
    mbar_wait(&sm.bar, 0);

    // ---- phase A: lane = lattice
    if (warp == 0) {
        const int e = env0 + lane;
        u64 xb = 0, zb = 0, meta = sm.st[ROW_META][lane], act[3] = {0, 0, 0};
        u32 flags = 0;
        if (!RESET) {
            xb = sm.st[ROW_XB][lane]; zb = sm.st[ROW_ZB][lane];
#pragma unroll
            for (int l = 0; l < 3; ++l) if (l < p.layers) act[l] = sm.st[ROW_ACT + l][lane];
            int a = (e < p.n) ? actions[e] : p.A - 1;
            if (a < 0 || a >= p.A) a = p.A - 1;
            const bool ident = (a == p.A - 1);
            const int layer = ident ? 0 : a / L::NQ, q = ident ? 0 : a % L::NQ;
            const u64 bit = ident ? 0ull : (1ull << (q + q / D));
            const u64 cur = layer == 0 ? act[0] : (layer == 1 ? act[1] : act[2]);
            const bool heavy = ident || (cur & bit) != 0;
            // Pauli applied by this action layer (Function_Library.py:253-304)
            bool fx, fz;
            if (p.model == DQ_MODEL_X) { fx = true; fz = false; }
            else if (p.use_y) { fx = layer <= 1; fz = layer >= 1; }
            else { fx = layer == 0; fz = layer == 1; }
            if (fx) xb ^= bit;
            if (fz) zb ^= bit;
            const u64 syn = true_syndrome<D>(xb, zb);
            const int label = homology_label<D>(xb, zb);
            u32 dn = (u32)(meta >> 63);
            float rw = 0.f;
            if (label == 0 && syn == 0) rw = 1.f;
            else if (referee_class<D>(p, syn) != label) dn = 1;
            if (!heavy) {
                if (layer == 0) act[0] |= bit; else if (layer == 1) act[1] |= bit; else act[2] |= bit;
            }
            meta = (meta & ~(1ull << 63)) | ((u64)dn << 63);
            if (e < p.n) {
                if (reward) reward[e] = rw;
                if (done_out) done_out[e] = (uint8_t)dn;
            }
            flags = (heavy ? 1u : 0u) | ((dn && auto_reset) ? 2u : 0u);
            sm.life_out[lane] = (int32_t)(u32)meta;
        } else {
            flags = 2u;           // reset keeps only the attempt counter (the RNG position)
        }
        sm.st[ROW_XB][lane] = xb; sm.st[ROW_ZB][lane] = zb; sm.st[ROW_META][lane] = meta;
#pragma unroll
        for (int l = 0; l < 3; ++l) if (l < p.layers) sm.st[ROW_ACT + l][lane] = act[l];
        const u32 tmask = __ballot_sync(0xffffffffu, flags != 0);
        if (flags) {
            const int pos = __popc(tmask & ((1u << lane) - 1));
            sm.task[pos] = (uint8_t)lane; sm.task_flags[pos] = (uint8_t)flags;
        }
        if (lane == 0) sm.ntask = __popc(tmask);
    }
    __syncthreads();

    // ---- phase B: warp per lattice that needs a new volume
    for (int t = warp; t < sm.ntask; t += kWarps) {
        const int slot = sm.task[t], flags = sm.task_flags[t];
        const u32 env_id = p.env_id_base + (u32)(env0 + slot);
        u64 xb = sm.st[ROW_XB][slot], zb = sm.st[ROW_ZB][slot], meta = sm.st[ROW_META][slot];
        u32 life = (u32)meta, attempts = (u32)(meta >> 32) & 0x7FFFFFFFu, dn = (u32)(meta >> 63);
        if (flags & 1) generate_volume<D>(p, sm, warp, lane, slot, env_id, xb, zb, life, attempts);
        if (!RESET && lane == 0) sm.life_out[slot] = (int32_t)life;
        if (flags & 2) {               // reset: zero frame, fresh counters, one more volume
            __syncwarp();
            xb = 0; zb = 0; life = 0; dn = 0;
            generate_volume<D>(p, sm, warp, lane, slot, env_id, xb, zb, life, attempts);
        }
        if (lane == 0) {
            sm.st[ROW_XB][slot] = xb; sm.st[ROW_ZB][slot] = zb;
            sm.st[ROW_META][slot] = meta_pack(life, attempts, dn);
            for (int l = 0; l < p.layers; ++l) sm.st[ROW_ACT + l][slot] = 0;
        }
    }
    __syncthreads();

    // ---- phase C: thread per (lattice, layer): legal mask + layer bitmap into the CTA bit stream
    for (int t = tid; t < kEnvsPerCta * C; t += kThreads) {
        const int slot = t % kEnvsPerCta, layer = t / kEnvsPerCta;
        const int e = env0 + slot;
        u64 w[L::PW];
        if (layer < p.vd) syndrome_layer_bitmap<D>(sm.st[ROW_SYN + layer][slot], w);
        else action_layer_bitmap<D>(sm.st[ROW_ACT + layer - p.vd][slot], w);
        const int off = slot * p.obs_bits + layer * L::P;
#pragma unroll
        for (int i = 0; i < L::PW; ++i) stream_or64(bits, off + 64 * i, w[i]);
        if (layer == 0 && e < p.n) {
            if (lifetime && !RESET) lifetime[e] = sm.life_out[slot];
            if (legal) {
                u64 summed = 0, acted = 0;
                for (int j = 0; j < p.vd; ++j) summed |= sm.st[ROW_SYN + j][slot];
                for (int l = 0; l < p.layers; ++l) acted |= sm.st[ROW_ACT + l][slot];
                const u64 lq = qubits_grid_to_compact<D>(qubits_adjacent_to<D>(summed) | qubits_neighbours_of<D>(acted));
                u64 mw[3] = {0, 0, 0};
                for (int l = 0; l < p.layers; ++l) {
                    const int o = l * L::NQ, i = o >> 6, s = o & 63;
                    mw[i] |= lq << s;
                    if (s && i + 1 < 3) mw[i + 1] |= lq >> (64 - s);
                }
                mw[(p.A - 1) >> 6] |= 1ull << ((p.A - 1) & 63);
                for (int i = 0; i < p.W; ++i) legal[(size_t)e * p.W + i] = mw[i];
            }
        }
    }
    fence_proxy_async();      // phase A/B wrote the state tile through the generic proxy
    __syncthreads();

    // ---- write the state tile back (TMA bulk store), overlapped with phase D
    if (tid == 0) {
        for (int r = 0; r < nrows_used; ++r) {
            if (r >= ROW_ACT + p.layers && r < ROW_SYN) continue;
            bulk_store(p.state + (size_t)r * p.npad + env0, &sm.st[r][0], kEnvsPerCta * 8);
        }
    }

    // ---- phase D: 16 stream bits -> 16 observation bytes per 128-bit store
    if (obs) {
        const int nvalid = min(kEnvsPerCta, p.n - env0);
        const long long vbytes = (long long)nvalid * p.obs_bits;
        uint8_t* out = obs + (size_t)env0 * p.obs_bits;
        const int units = (int)(vbytes >> 4);
        for (int u = tid; u < units; u += kThreads) {
            const u32 word = bits[u >> 1];
            const u32 h = (u & 1) ? (word >> 16) : (word & 0xFFFFu);
            uint4 v;
            v.x = ((h & 0xFu) * 0x00204081u) & 0x01010101u;
            v.y = (((h >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
            v.z = (((h >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
            v.w = (((h >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
            *reinterpret_cast<uint4*>(out + ((size_t)u << 4)) = v;
        }
        for (long long b = ((long long)units << 4) + tid; b < vbytes; b += kThreads)
            out[b] = (uint8_t)((bits[b >> 5] >> (b & 31)) & 1u);
    }
    if (tid == 0) bulk_commit_wait_read();
}

// Uniform pick over the sorted legal actions.  `ctr` (optional) = {step index, finished-CTA count} in
// device memory: when given, the step index is read from it and advanced by the last CTA to finish,
// so the launch can sit in a CUDA graph and still draw fresh words on every replay.
__global__ void policy_random_legal_kernel(const u64* __restrict__ legal, int n, int W, int A, u32 env_id_base,
                                           u32 step, u32* __restrict__ ctr, u32 k0, u32 k1,
                                           int32_t* __restrict__ actions) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctr) step = *reinterpret_cast<volatile u32*>(ctr);
    if (e < n) {
        u64 m[3] = {0, 0, 0};
        int cnt = 0;
        for (int i = 0; i < W; ++i) { m[i] = legal[(size_t)e * W + i]; cnt += popc64(m[i]); }
        const Philox4 u = philox4x32_10(env_id_base + (u32)e, step, 0u, 1u, k0, k1);
        int pick = (int)mulhi32(u.x, (u32)cnt), act = A - 1;
        for (int i = 0; i < W; ++i) {
            const int c = popc64(m[i]);
            if (pick < c) { act = i * 64 + select64(m[i], pick); break; }
            pick -= c;
        }
        actions[e] = act;
    }
    if (ctr) {
        __syncthreads();                      // every thread of this CTA has read the step index
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) { ctr[1] = 0; atomicAdd(ctr, 1u); }
        }
    }
}

__global__ void set_u32_kernel(u32* p, u32 v) { p[0] = v; p[1] = 0; }

}  // namespace dq

// ======================================================================== host side / C ABI
using namespace dq;

struct dq_env {
    EnvParams p;
    int device;
    int state_rows;
    size_t smem_bytes;
    // staging for the *_host entry points
    cudaStream_t hstream;
    u32* policy_ctr;             // {step index, finished-CTA count} for dq_policy_random_legal_next
    int32_t* s_actions; uint8_t* s_obs; float* s_reward; uint8_t* s_done; int32_t* s_life; u64* s_legal;
};

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define DQ_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) return fail(DQ_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct DeviceGuard {
    int prev; bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(false) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() { if (ok && prev >= 0) cudaSetDevice(prev); }
};

static u32 threshold_u32(double p) {
    if (!(p > 0.0)) return 0u;
    double t = floor(p * 4294967296.0);
    if (t >= 4294967295.0) return 0xFFFFFFFFu;
    return (u32)t;
}
static void set_thresholds(EnvParams& p, double p_phys, double p_meas) {
    p.T = threshold_u32(p_phys); p.Tm = threshold_u32(p_meas);
    p.T1 = p.T / 3; p.T2 = (u32)((2ull * p.T) / 3);
}

extern "C" const char* dq_last_error(void) { return g_err.c_str(); }
extern "C" int dq_version(void) { return 100; }
extern "C" int64_t dq_launch_count(void) { return g_launches.load(); }

extern "C" int dq_env_create(dq_env** out, int d, int error_model, int use_Y, int volume_depth, double p_phys,
                             double p_meas, int64_t n_envs, uint64_t seed, int64_t env_id_base, int device) {
    if (!out) return fail(DQ_EINVAL, "out is NULL");
    *out = nullptr;
    if (d != 3 && d != 5 && d != 7) return fail(DQ_EINVAL, "d must be 3, 5 or 7 (odd, FL:38; boards are packed into 64 bits)");
    if (error_model != DQ_MODEL_X && error_model != DQ_MODEL_DP) return fail(DQ_EINVAL, "error_model must be DQ_MODEL_X or DQ_MODEL_DP");
    if (volume_depth < 1 || volume_depth > kMaxVd) return fail(DQ_EINVAL, "volume_depth must be in [1,8]");
    if (n_envs < 1 || n_envs > (1ll << 30)) return fail(DQ_EINVAL, "n_envs out of range");
    if (env_id_base < 0 || env_id_base + n_envs > (1ll << 32)) return fail(DQ_EINVAL, "stream ids must fit 32 bits");
    int ndev = 0;
    DQ_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(DQ_EINVAL, "no such CUDA device");
    DeviceGuard g(device);
    if (!g.ok) return fail(DQ_ECUDA, "cudaSetDevice failed");

    dq_env* e = new dq_env();
    memset(e, 0, sizeof(*e));
    EnvParams& p = e->p;
    p.d = d; p.model = error_model; p.use_y = use_Y ? 1 : 0; p.vd = volume_depth;
    p.layers = error_model == DQ_MODEL_X ? 1 : (use_Y ? 3 : 2);
    p.A = p.layers * d * d + 1;
    p.W = (p.A + 63) / 64;
    p.n = (int)n_envs; p.npad = (int)((n_envs + 31) / 32 * 32);
    const int items = volume_depth * (2 * d * d - 1);
    p.rounds = (items + 127) / 128;
    p.obs_bits = (volume_depth + p.layers) * (2 * d + 1) * (2 * d + 1);
    set_thresholds(p, p_phys, p_meas);
    p.k0 = (u32)seed; p.k1 = (u32)(seed >> 32);
    p.env_id_base = (u32)env_id_base;
    p.ref_mode = -1;
    e->device = device;
    e->state_rows = ROW_SYN + volume_depth;
    e->smem_bytes = ((sizeof(Smem) + 15) & ~size_t(15)) + (size_t)(p.obs_bits + 2) * 4;
    cudaError_t err = cudaMalloc(&p.state, (size_t)e->state_rows * p.npad * sizeof(u64));
    if (err != cudaSuccess) { delete e; return fail(DQ_ECUDA, std::string("cudaMalloc(state): ") + cudaGetErrorString(err)); }
    err = cudaMemset(p.state, 0, (size_t)e->state_rows * p.npad * sizeof(u64));
    if (err != cudaSuccess) { cudaFree(p.state); delete e; return fail(DQ_ECUDA, std::string("cudaMemset(state): ") + cudaGetErrorString(err)); }
    err = cudaMalloc(&e->policy_ctr, 2 * sizeof(u32));
    if (err == cudaSuccess) err = cudaMemset(e->policy_ctr, 0, 2 * sizeof(u32));
    if (err != cudaSuccess) { cudaFree(p.state); delete e; return fail(DQ_ECUDA, std::string("cudaMalloc(policy_ctr): ") + cudaGetErrorString(err)); }
    *out = e;
    return DQ_OK;
}

extern "C" int dq_env_destroy(dq_env* e) {
    if (!e) return DQ_OK;
    DeviceGuard g(e->device);
    if (e->hstream) {
        cudaStreamSynchronize(e->hstream);
        cudaFree(e->s_actions); cudaFree(e->s_obs); cudaFree(e->s_reward); cudaFree(e->s_done); cudaFree(e->s_life); cudaFree(e->s_legal);
        cudaStreamDestroy(e->hstream);
    }
    cudaFree(e->p.state);
    cudaFree(e->policy_ctr);
    delete e;
    return DQ_OK;
}

extern "C" int dq_env_info(const dq_env* e, int what, int64_t* out) {
    if (!e || !out) return fail(DQ_EINVAL, "NULL argument");
    const EnvParams& p = e->p;
    switch (what) {
        case DQ_INFO_NUM_ACTIONS: *out = p.A; break;
        case DQ_INFO_OBS_CHANNELS: *out = p.vd + p.layers; break;
        case DQ_INFO_OBS_SIDE: *out = 2 * p.d + 1; break;
        case DQ_INFO_MASK_WORDS: *out = p.W; break;
        case DQ_INFO_STATE_WORDS: *out = e->state_rows; break;
        case DQ_INFO_STATE_STRIDE: *out = p.npad; break;
        case DQ_INFO_NUM_STABS: *out = p.d * p.d - 1; break;
        case DQ_INFO_N_TYPE3: *out = (p.d * p.d - 1) / 2; break;
        case DQ_INFO_N_TYPE1: *out = (p.d * p.d - 1) / 2; break;
        case DQ_INFO_RNG_BLOCKS: *out = 32 * p.rounds; break;
        default: return fail(DQ_EINVAL, "unknown info selector");
    }
    return DQ_OK;
}

extern "C" int dq_env_set_noise(dq_env* e, double p_phys, double p_meas) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    if (!(p_phys >= 0.0) || !(p_meas >= 0.0)) return fail(DQ_EINVAL, "probabilities must be >= 0");
    set_thresholds(e->p, p_phys, p_meas);
    return DQ_OK;
}

extern "C" int dq_env_set_referee_lut(dq_env* e, int mode, const void* lut_a, int64_t bytes_a, const void* lut_b, int64_t bytes_b) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    const EnvParams& p = e->p;
    const int ns = p.d * p.d - 1, nt = ns / 2;
    if (mode == DQ_REFEREE_JOINT) {
        if (ns > 30) return fail(DQ_EINVAL, "joint referee table needs d*d-1 <= 30 stabilizers; use DQ_REFEREE_SPLIT");
        if (!lut_a || bytes_a < ((1ll << ns) + 3) / 4) return fail(DQ_EINVAL, "joint referee table too small");
    } else if (mode == DQ_REFEREE_SPLIT) {
        if (!lut_a || bytes_a < ((1ll << nt) + 3) / 4) return fail(DQ_EINVAL, "X-class referee table too small");
        if (p.model == DQ_MODEL_DP && (!lut_b || bytes_b < ((1ll << nt) + 3) / 4)) return fail(DQ_EINVAL, "Z-class referee table missing or too small");
    } else return fail(DQ_EINVAL, "unknown referee mode");
    e->p.ref_mode = mode;
    e->p.lut_a = (const uint8_t*)lut_a;
    e->p.lut_b = (const uint8_t*)lut_b;
    return DQ_OK;
}

template <bool RESET>
static int launch_env(dq_env* e, const int32_t* actions, uint8_t* obs, float* reward, uint8_t* done, int32_t* lifetime,
                      u64* legal, int auto_reset, cudaStream_t st) {
    const EnvParams& p = e->p;
    if (obs && (reinterpret_cast<uintptr_t>(obs) & 15)) return fail(DQ_EINVAL, "obs must be 16-byte aligned");
    const dim3 grid(p.npad / kEnvsPerCta), block(kThreads);
    switch (p.d) {
        case 3: env_step_kernel<3, RESET><<<grid, block, e->smem_bytes, st>>>(p, actions, obs, reward, done, lifetime, legal, auto_reset); break;
        case 5: env_step_kernel<5, RESET><<<grid, block, e->smem_bytes, st>>>(p, actions, obs, reward, done, lifetime, legal, auto_reset); break;
        case 7: env_step_kernel<7, RESET><<<grid, block, e->smem_bytes, st>>>(p, actions, obs, reward, done, lifetime, legal, auto_reset); break;
    }
    g_launches.fetch_add(1);
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

extern "C" int dq_env_reset(dq_env* e, uint8_t* obs, uint64_t* legal, dq_stream stream) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    DeviceGuard g(e->device);
    return launch_env<true>(e, nullptr, obs, nullptr, nullptr, nullptr, (u64*)legal, 1, (cudaStream_t)stream);
}

extern "C" int dq_env_step(dq_env* e, const int32_t* actions, uint8_t* obs, float* reward, uint8_t* done,
                           int32_t* lifetime, uint64_t* legal, int auto_reset, dq_stream stream) {
    if (!e || !actions) return fail(DQ_EINVAL, "env / actions is NULL");
    if (e->p.ref_mode < 0) return fail(DQ_ESTATE, "no referee set: call dq_env_set_referee_lut first (the reference needs static_decoder too)");
    DeviceGuard g(e->device);
    return launch_env<false>(e, actions, obs, reward, done, lifetime, (u64*)legal, auto_reset, (cudaStream_t)stream);
}

static int ensure_staging(dq_env* e) {
    if (e->hstream) return DQ_OK;
    const EnvParams& p = e->p;
    DQ_CUDA(cudaStreamCreateWithFlags(&e->hstream, cudaStreamNonBlocking));
    DQ_CUDA(cudaMalloc(&e->s_actions, (size_t)p.n * 4));
    DQ_CUDA(cudaMalloc(&e->s_obs, (size_t)p.n * p.obs_bits));
    DQ_CUDA(cudaMalloc(&e->s_reward, (size_t)p.n * 4));
    DQ_CUDA(cudaMalloc(&e->s_done, (size_t)p.n));
    DQ_CUDA(cudaMalloc(&e->s_life, (size_t)p.n * 4));
    DQ_CUDA(cudaMalloc(&e->s_legal, (size_t)p.n * p.W * 8));
    return DQ_OK;
}

static int copy_out(dq_env* e, uint8_t* h_obs, float* h_reward, uint8_t* h_done, int32_t* h_life, uint64_t* h_legal) {
    const EnvParams& p = e->p;
    cudaStream_t s = e->hstream;
    if (h_obs) DQ_CUDA(cudaMemcpyAsync(h_obs, e->s_obs, (size_t)p.n * p.obs_bits, cudaMemcpyDeviceToHost, s));
    if (h_reward) DQ_CUDA(cudaMemcpyAsync(h_reward, e->s_reward, (size_t)p.n * 4, cudaMemcpyDeviceToHost, s));
    if (h_done) DQ_CUDA(cudaMemcpyAsync(h_done, e->s_done, (size_t)p.n, cudaMemcpyDeviceToHost, s));
    if (h_life) DQ_CUDA(cudaMemcpyAsync(h_life, e->s_life, (size_t)p.n * 4, cudaMemcpyDeviceToHost, s));
    if (h_legal) DQ_CUDA(cudaMemcpyAsync(h_legal, e->s_legal, (size_t)p.n * p.W * 8, cudaMemcpyDeviceToHost, s));
    DQ_CUDA(cudaStreamSynchronize(s));
    return DQ_OK;
}

extern "C" int dq_env_reset_host(dq_env* e, uint8_t* h_obs, uint64_t* h_legal) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    DeviceGuard g(e->device);
    int rc = ensure_staging(e);
    if (rc) return rc;
    rc = launch_env<true>(e, nullptr, h_obs ? e->s_obs : nullptr, nullptr, nullptr, nullptr, h_legal ? e->s_legal : nullptr, 1, e->hstream);
    if (rc) return rc;
    return copy_out(e, h_obs, nullptr, nullptr, nullptr, h_legal);
}

extern "C" int dq_env_step_host(dq_env* e, const int32_t* h_actions, uint8_t* h_obs, float* h_reward, uint8_t* h_done,
                                int32_t* h_life, uint64_t* h_legal, int auto_reset) {
    if (!e || !h_actions) return fail(DQ_EINVAL, "env / actions is NULL");
    if (e->p.ref_mode < 0) return fail(DQ_ESTATE, "no referee set: call dq_env_set_referee_lut first");
    DeviceGuard g(e->device);
    int rc = ensure_staging(e);
    if (rc) return rc;
    DQ_CUDA(cudaMemcpyAsync(e->s_actions, h_actions, (size_t)e->p.n * 4, cudaMemcpyHostToDevice, e->hstream));
    rc = launch_env<false>(e, e->s_actions, h_obs ? e->s_obs : nullptr, h_reward ? e->s_reward : nullptr,
                           h_done ? e->s_done : nullptr, h_life ? e->s_life : nullptr, h_legal ? e->s_legal : nullptr,
                           auto_reset, e->hstream);
    if (rc) return rc;
    return copy_out(e, h_obs, h_reward, h_done, h_life, h_legal);
}

extern "C" int dq_env_get_state(dq_env* e, uint64_t* dev_words, dq_stream stream) {
    if (!e || !dev_words) return fail(DQ_EINVAL, "NULL argument");
    DeviceGuard g(e->device);
    DQ_CUDA(cudaMemcpyAsync(dev_words, e->p.state, (size_t)e->state_rows * e->p.npad * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DQ_OK;
}

extern "C" int dq_env_set_state(dq_env* e, const uint64_t* dev_words, dq_stream stream) {
    if (!e || !dev_words) return fail(DQ_EINVAL, "NULL argument");
    DeviceGuard g(e->device);
    DQ_CUDA(cudaMemcpyAsync(e->p.state, dev_words, (size_t)e->state_rows * e->p.npad * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DQ_OK;
}

static int launch_policy(const dq_env* e, const uint64_t* legal, u32 step, u32* ctr, int32_t* actions, cudaStream_t st) {
    const EnvParams& p = e->p;
    policy_random_legal_kernel<<<(p.n + 255) / 256, 256, 0, st>>>((const u64*)legal, p.n, p.W, p.A, p.env_id_base, step, ctr,
                                                                  p.k0, p.k1, actions);
    g_launches.fetch_add(1);
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

extern "C" int dq_policy_random_legal(const dq_env* e, const uint64_t* legal, uint32_t step, int32_t* actions, dq_stream stream) {
    if (!e || !legal || !actions) return fail(DQ_EINVAL, "NULL argument");
    DeviceGuard g(e->device);
    return launch_policy(e, legal, step, nullptr, actions, (cudaStream_t)stream);
}

extern "C" int dq_policy_seek(dq_env* e, uint32_t step, dq_stream stream) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    DeviceGuard g(e->device);
    set_u32_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(e->policy_ctr, step);
    g_launches.fetch_add(1);
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

extern "C" int dq_policy_random_legal_next(dq_env* e, const uint64_t* legal, int32_t* actions, dq_stream stream) {
    if (!e || !legal || !actions) return fail(DQ_EINVAL, "NULL argument");
    DeviceGuard g(e->device);
    return launch_policy(e, legal, 0u, e->policy_ctr, actions, (cudaStream_t)stream);
}
