// dq_env.cu -- sm_100a environment-step kernel + C ABI (include/dq_decoding.h).
//
// One CTA advances a tile of 32 independent lattices with three kinds of warps that never meet at a block barrier
// inside the step loop (a launch runs one step -- dq_env_step* -- or a rollout of many -- dq_env_rollout_random):
//
//   PHYSICS (warp 0, lane = lattice, the lattice's frame / counters / action boards / legal mask live in registers for the
//     whole launch): (built-in random-legal pick,) action -> Pauli frame, true syndrome by shifted XORs, homology label,
//     referee table lookup, reward / done; a step that needs a fresh syndrome volume (identity | repeat | restart of a
//     finished lattice) POPS it from the lattice's volume queue; then lifetime, legal-move mask, the small outputs.
//   GENERATORS (warps 1+W..): keep every lattice's volume queue full.  A volume attempt is drawn from
//     Philox(lattice, attempt index) alone, and the syndrome map is XOR-linear, so everything about attempt t can be
//     computed without the lattice's state:  g_j = syndrome(e_0 ^ .. ^ e_j) ^ m_j  (e_j / m_j: the data-qubit and measurement
//     flips of slice j) and the frame delta E = e_0 ^ .. ^ e_{vd-1}.  Popping it on frame F is then
//     f_j = g_j ^ syndrome(F), F ^= E -- a handful of XORs per lattice -- and the attempt is all-trivial iff every f_j == 0.
//     The attempt index is a monotone per-lattice counter, so a queued attempt is always consumed eventually, by a heavy
//     step or by a restart alike: nothing is speculative.  The queue (DQ_QDEPTH attempts per lattice) persists in
//     device memory between launches.
//   WRITERS (warps 1..W): own the tile's rendered (2d+1)^2-cell layer bitmaps and the tile's observation BIT STREAM in its final
//     layout (lattice-major concatenation of the layer bitmaps, no padding: byte i of the tile's observations is bit i).
//     Per step they take the physics warp's record (one new action cell, or the slices of a fresh volume), re-render
//     what changed, and expand the stream to bytes: thread per 32-bit word, 256-entry table, two 128-bit stores (the tile's 32
//     observations are one contiguous, 32-byte aligned span of HBM).
//
// Hand-offs: physics -> writers through a two-slot record ring guarded by named barriers (bar.arrive / bar.sync, full and
// empty per slot); physics <-> generators through per-lattice head / tail counters in shared memory (release / acquire by
// __threadfence_block + volatile accesses).
// Earlier kernels (barrier-phased 16-lattice tiles, v1-v8) and the ncu evidence that led here: profiles/README.md.
//
// Replaces (reference paths relative to example_notebooks/): Environments.py:99-115 (reset),
// :118-204 (step), :206-235, :238-314 and the Function_Library.py helpers they call.
#include <cuda_runtime.h>
#if defined(__linux__)
#include <sched.h>
#endif
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
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

#include "../../include/dq_decoding.h"
#include "dq_lattice.cuh"

namespace dq {

// Warp roles per CTA: 1 physics warp + writer warps + generator warps.  The physics warp is the one dependent chain that sets the
// step time, and every other resident warp competes with it for issue slots, so the roles are sized to keep up and no larger:
// at d <= 5 one writer warp (847 stream words per tile-step) and three generator warps; at d = 7 the observation is 2.4x larger and a
// volume attempt draws three times the Philox blocks.
#ifndef DQ_WRITERS
#define DQ_WRITERS 1             // observation-writer warps per CTA, d <= 5
#endif
#ifndef DQ_GENS
#define DQ_GENS 3                // volume-generator warps per CTA, d <= 5
#endif
#ifndef DQ_WRITERS_D7
#define DQ_WRITERS_D7 2
#endif
#ifndef DQ_GENS_D7
#define DQ_GENS_D7 5
#endif
#ifndef DQ_QDEPTH
#define DQ_QDEPTH 4              // queued volume attempts per lattice (power of two)
#endif
#ifndef DQ_MIN_BLOCKS
#define DQ_MIN_BLOCKS 4          // resident CTAs per SM the register allocation must allow (16 384 lattices = 512 tiles = 3.5 per SM)
#endif
constexpr int kLpc = 32;                             // lattices per CTA = lanes of the physics warp
constexpr int kQ = DQ_QDEPTH;
template <int D> struct Roles {
    static constexpr int kWriters = D >= 7 ? DQ_WRITERS_D7 : DQ_WRITERS, kGens = D >= 7 ? DQ_GENS_D7 : DQ_GENS;
    static constexpr int kThreads = 32 * (1 + kWriters + kGens);
    static constexpr int kWriterThreads = 32 * kWriters;
    static constexpr int kOwned = (kLpc + kGens - 1) / kGens;   // lattices a generator warp serves: l = i*kGens + g
    static_assert(kWriters >= 1 && kGens >= 1 && kThreads <= 1024, "warp roles");
};
constexpr int kMaxGens = 6;                          // shared-memory arrays and doorbell barriers are sized for this many generator warps
static_assert((kQ & (kQ - 1)) == 0 && kQ >= 2, "queue depth must be a power of two");
constexpr int kMaxVd = 8;
constexpr int kMaxLayers = kMaxVd + 3;
constexpr int kQW = kMaxVd + 2;                      // words of a queued attempt: g_0..g_{vd-1}, EX, EZ
constexpr int kMaxSms = 1024;
constexpr long long kSpinTrapClocks = 4000000000ll;   // ~2 s of SM clocks: far beyond any legitimate wait for a generator warp
#ifndef DQ_ROTATE
#define DQ_ROTATE 0               // 1: rotate the warp roles of the CTAs that share an SM (see env_step_kernel; measured: no gain)
#endif
#ifndef DQ_PHYS_LAST
#define DQ_PHYS_LAST 0            // 1: the physics role on the CTA's last warp instead of its first
#endif
#ifndef DQ_RING
#define DQ_RING 4                // record slots between the physics warp and the writers (power of two)
#endif
constexpr int kRing = DQ_RING;
constexpr int kStreamWords = kLpc * kMaxLayers * 15 * 15 / 32 + 1;
constexpr int BAR_FULL = 1, BAR_EMPTY = BAR_FULL + kRing, BAR_WRITERS = BAR_EMPTY + kRing, BAR_DOOR = BAR_WRITERS + 1;    // named barriers (0 = __syncthreads)
static_assert(BAR_DOOR + kMaxGens <= 16, "the hardware has 16 named barriers per CTA");
static_assert(Roles<3>::kGens <= kMaxGens && Roles<7>::kGens <= kMaxGens, "generator warps");
// The reference loops until a volume is non-trivial, forever if p_phys = p_meas = 0 on a clean frame.
// A kernel must end: after this many attempts on one volume the (trivial) volume is accepted.
constexpr int kMaxAttemptsPerCall = 1 << 20;

// rows of the packed state matrix: frame planes, counters, action boards, OR of the volume's slices,
// then the rendered bitmap of observation layer l in rows ROW_BM + l*PW .. +PW-1 (PW = ceil((2d+1)^2/64))
constexpr int ROW_XB = 0, ROW_ZB = 1, ROW_META = 2, ROW_ACT = 3, ROW_SUM = 6, ROW_BM = 7;
// meta word: [lifetime:32][attempt counter:31][done:1]
DQ_HD u64 meta_pack(u32 life, u32 attempts, u32 done) { return (u64)life | ((u64)(attempts & 0x7FFFFFFFu) << 32) | ((u64)done << 63); }

struct EnvParams {
    int d, model, use_y, vd, layers, A, W;      // W = mask words
    int n, npad;                                // lattices, padded to 32
    int rounds;                                 // Philox rounds per volume attempt (B = 32*rounds)
    int obs_bits;                               // C*H*H
    u32 ob_magic;                               // ceil(2^32 / obs_bits): g / obs_bits == umulhi(g, ob_magic) for g * obs_bits < 2^32
    u32 T, T1, T2, Tm;                          // thresholds (RNG contract)
    u32 Tmx;                                    // max(T, Tm): the one-compare screen of draw_flip_masks
    u64 idw[3];                                 // legal-mask words with only the identity action (A-1) set
    u32 k0, k1;                                 // Philox key
    u32 env_id_base;
    int ref_mode;
    int q_reset;                                // the queued attempts were drawn under other noise rates: discard them
    int max_attempts;                           // all-trivial attempts after which a volume is accepted as it is (default 2^20)
    const uint8_t* lut_a;
    const uint8_t* lut_b;
    u64* state;                                 // [STATE_WORDS][npad]
    u64* queue;                                 // [tile][kQ][kQW][kLpc]: attempt t of a lattice sits in slot t % kQ
    u32* qtail;                                 // [npad]: first attempt index NOT yet queued
    u32* sm_arrivals;                           // [kMaxSms]: CTAs that have started on each SM, ever (role rotation, see env_step_kernel)
};

__device__ __forceinline__ int lut2(const uint8_t* lut, u32 idx) { return (__ldg(lut + (idx >> 2)) >> ((idx & 3) * 2)) & 3; }

template <int D>
__device__ __forceinline__ int referee_class(const EnvParams& p, u64 syn) {
    if (p.ref_mode == DQ_REFEREE_JOINT) return lut2(p.lut_a, stabs_grid_to_joint_index<D>(syn));
    int c = lut2(p.lut_a, stabs_grid_to_type_index<D, 1>(syn)) & 1;
    if (p.model == DQ_MODEL_DP && p.lut_b) c |= (lut2(p.lut_b, stabs_grid_to_type_index<D, 0>(syn)) & 1) << 1;
    return c;
}

// ---- the three synchronisation primitives of the warp-specialised CTA (tests/host/cuda_emu.h serves them on the CPU) ----
__device__ __forceinline__ void bar_sync_named(int id, int count) {          // wait until `count` threads have arrived at barrier `id`
#ifdef DQ_EMU
    dq_emu::named_barrier(id, count);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void bar_arrive_named(int id, int count) {        // arrive without waiting
#ifdef DQ_EMU
    dq_emu::named_barrier_arrive(id, count);
#else
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void spin_pause(unsigned ns = 20) {
#ifdef DQ_EMU
    (void)ns;
    dq_emu::yield();
#else
    __nanosleep(ns);
#endif
}
// Programmatic dependent launch (single-step launches are queued with the attribute, see launch_env): the NEXT launch's CTAs may be
// scheduled as soon as every CTA of this one has said so, and wait in griddepcontrol.wait -- before their first access to global
// memory -- until this grid has completed and its writes are visible.  Both are no-ops in a launch without the attribute.
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
__device__ __forceinline__ u32 sm_id() {
#ifdef DQ_EMU
    return blockIdx.x;
#else
    u32 r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
#endif
}

// A fired draw (rare: p ~ 1e-2 per draw) is folded into the per-slice accumulators of the warp, acc[kind][slice] as two 32-bit
// halves: kind 0 / 1 = the X / Z flips of the data qubits ACCUMULATED up to and including that slice (a flip in slice j is XOR-ed
// into every slice j' >= j: the frame keeps it), kind 2 = the slice's measurement flips.
template <int D>
__device__ __noinline__ void record_event(u32* acc, int item, u32 uv, int nq_items, int n_items, u32 T1, u32 T2, int dp, int vd) {
    typedef Lat<D> L;
    if (item < nq_items) {
        const int j = item / L::NQ, q = item - j * L::NQ, pos = q + q / D;
        const u32 bit = 1u << (pos & 31);
        const bool fx = !dp || uv < T2, fz = dp && uv >= T1;                                  // X or Y / Y or Z
        for (int jj = j; jj < vd; ++jj) {
            if (fx) atomicXor(&acc[((0 * kMaxVd + jj) << 1) + (pos >> 5)], bit);
            if (fz) atomicXor(&acc[((1 * kMaxVd + jj) << 1) + (pos >> 5)], bit);
        }
    } else if (item < n_items) {
        const int mi = item - nq_items, j = mi / L::NS, k = mi - j * L::NS, pos = L::stab_pos(k);
        atomicXor(&acc[((2 * kMaxVd + j) << 1) + (pos >> 5)], 1u << (pos & 31));
    }
}

// The draws of ONE volume attempt of one lattice, executed by a full warp: R rounds of one Philox4x32-10 block per lane
// (draw i = word i/B of block i%B), issued two rounds at a time so two Philox chains overlap.  Every lane thresholds its own
// draws; the few that fire are XOR-ed into the accumulators acc[kind][slice] (record_event).
template <int D>
__device__ __forceinline__ void draw_flip_masks(const EnvParams& p, u32* acc, int lane, u32 env_id, u32 attempt) {
    typedef Lat<D> L;
    const int vd = p.vd, R = p.rounds, nq_items = vd * L::NQ, n_items = vd * (L::NQ + L::NS);
    const int dp = p.model == DQ_MODEL_DP;
    if (lane < 3 * kMaxVd) { acc[2 * lane] = 0; acc[2 * lane + 1] = 0; }
    __syncwarp();
    for (int r = 0; r < R; r += 2) {
        const Philox4 u0 = philox4x32_10(env_id, attempt, (u32)(r * 32 + lane), 0u, p.k0, p.k1);
        Philox4 u1;
        u1.x = u1.y = u1.z = u1.w = 0xffffffffu;                      // an odd R has no second round: words that can never fire
        if (r + 1 < R) u1 = philox4x32_10(env_id, attempt, (u32)(r * 32 + 32 + lane), 0u, p.k0, p.k1);
        const u32 uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        // Screen the lane's eight draws against the larger threshold with one predicate chain; only a lane that may have
        // fired (p ~ 1e-2 per draw) walks its candidates, one per set bit, and applies the exact per-item threshold.
        bool any = false;
#pragma unroll
        for (int w = 0; w < 8; ++w) any |= uu[w] < p.Tmx;
        if (any) {
            u32 hits = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) hits |= (uu[w] < p.Tmx ? 1u : 0u) << w;
            do {
                const int w = __ffs((int)hits) - 1;
                hits &= hits - 1;
                const u32 lo4 = (w & 2) ? ((w & 1) ? uu[3] : uu[2]) : ((w & 1) ? uu[1] : uu[0]);
                const u32 hi4 = (w & 2) ? ((w & 1) ? uu[7] : uu[6]) : ((w & 1) ? uu[5] : uu[4]);
                const u32 uv = (w & 4) ? hi4 : lo4;
                const int item = ((w & 3) * R + r + (w >> 2)) * 32 + lane;      // = word*B + block
                const u32 thr = (item < nq_items) ? p.T : p.Tm;
                if (uv < thr) record_event<D>(acc, item, uv, nq_items, n_items, p.T1, p.T2, dp, vd);
            } while (hits);
        }
    }
    __syncwarp();
}

// One volume attempt (the body of the loop at Environments.py:158-170 / :216-230) in its state-free form, executed by a
// full warp into queue slot `dst` (word w of the entry at dst[w * kLpc]): lane j < vd owns slice j -- the accumulators hold the
// data-qubit flips accumulated up to every slice, one shifted-XOR syndrome per lane gives the slice's g_j.
template <int D>
__device__ __forceinline__ void generate_attempt(const EnvParams& p, u32* acc, int lane, u32 env_id, u32 attempt, volatile u64* dst) {
    const int vd = p.vd;
    draw_flip_masks<D>(p, acc, lane, env_id, attempt);
    if (lane < vd) {
        const u64* a64 = reinterpret_cast<const u64*>(acc);
        const u64 ex = a64[0 * kMaxVd + lane], ez = a64[1 * kMaxVd + lane], m = a64[2 * kMaxVd + lane];
        dst[lane * kLpc] = true_syndrome<D>(ex, ez) ^ m;
        if (lane == vd - 1) { dst[kMaxVd * kLpc] = ex; dst[(kMaxVd + 1) * kLpc] = ez; }
    }
    __syncwarp();                                      // acc is rewritten by the next attempt
}

struct Rollout {            // multi-step launch: see env_step_kernel
    int nsteps, slots, first_slot;
    size_t slot_bytes;      // bytes between observation slots
    size_t out_stride;      // elements between the per-step rows of reward / done / lifetime / actions (legal: x W)
};

struct Smem {
    u32 stream[kStreamWords];         // the tile's observation bit stream: lattice-major concatenation of the layer bitmaps, P bits each,
                                      // no padding = byte i of the tile's observations is bit i; kept in step with the state's bitmap rows
    u64 lut8[256];                    // byte -> its 8 bits as 8 bytes of 0/1
    // physics -> writers, slot = step & 1
    u64 rec_f[kRing][kMaxVd][kLpc];   // the slices of a lattice's fresh volume
    int rec_actbit[kRing][kLpc];      // light step: (action layer << 16) | cell bit to set, else -1
    u32 rec_volmask[kRing];           // lattices (bit = lane) that drew a fresh volume
    // physics <-> generators
    u64 q[kQ][kQW][kLpc];             // rollouts: the tile's volume queues (a single-step launch works on the device-memory copy)
    u32 head[kLpc], tail[kLpc];       // attempt indices: next to pop (= the lattice's attempt counter) / first not yet queued
    u32 quit;                         // the physics warp has popped its last volume of the launch
    u32 gsleep[kMaxGens];                // 1: generator warp g found all its queues full and is about to sleep (or sleeps) at its doorbell
    u32 rot;                          // warp-role rotation of this CTA
    __align__(16) u32 acc[kMaxGens][3 * kMaxVd * 2];   // per-generator-warp flip accumulators (read back as 64-bit words)
};

// Layer bitmap w (P bits in PW u64 words, bits >= P zero) -> its place in the tile's bit stream: bits [off, off + P), at any alignment.
// Whole stream words are stored; the two end words are shared with the neighbouring layers, which other threads may be placing right
// now, so there only this layer's bits are replaced, atomically.
template <int D>
__device__ __forceinline__ void place_layer(u32* stream, int off, const u64 (&w)[Lat<D>::PW]) {
    constexpr int P = Lat<D>::P, PW = Lat<D>::PW, NV = 2 * PW, JMAX = (P + 62) / 32;
    u32 v[NV];
#pragma unroll
    for (int i = 0; i < PW; ++i) { v[2 * i] = (u32)w[i]; v[2 * i + 1] = (u32)(w[i] >> 32); }
    const int k0 = off >> 5, sh = off & 31, end = sh + P;            // the layer occupies bits [sh, end) counted from word k0
#pragma unroll
    for (int j = 0; j < JMAX; ++j) {
        if (j * 32 < end) {
            const u32 lo = (j >= 1 && j - 1 < NV) ? v[j >= 1 ? j - 1 : 0] : 0u, hi = j < NV ? v[j < NV ? j : 0] : 0u;
            const u32 slice = __funnelshift_l(lo, hi, sh);            // (hi << sh) | (lo >> (32 - sh))
            const int b0 = j == 0 ? sh : 0, b1 = min(end - j * 32, 32);
            if (b0 == 0 && b1 == 32) stream[k0 + j] = slice;
            else {
                const u32 mask = (b1 == 32 ? 0xffffffffu : ((1u << b1) - 1u)) & ~((1u << b0) - 1u);
                atomicAnd(&stream[k0 + j], ~mask);
                atomicOr(&stream[k0 + j], slice & mask);
            }
        }
    }
}

__device__ __forceinline__ uint4 expand16(const Smem& sm, u32 h) {     // low 16 bits -> 16 bytes of 0/1 (two table lookups)
    const uint2 a = *reinterpret_cast<const uint2*>(&sm.lut8[h & 0xFFu]);
    const uint2 b = *reinterpret_cast<const uint2*>(&sm.lut8[(h >> 8) & 0xFFu]);
    return make_uint4(a.x, a.y, b.x, b.y);
}

// Observation bytes are never read back by this kernel: evict-first stores (STG.E.EF.128) keep the ring of them that streams
// through L2 from evicting the 4 MB joint referee table.
__device__ __forceinline__ void store_obs16(uint8_t* dst, const uint4 v) { __stcs(reinterpret_cast<uint4*>(dst), v); }

// Legal-move mask words of one lattice (Environments.py:238-271 in closed form): qubits touching the summed faulty syndrome
// or next to an already acted-on qubit, in every action layer, plus the identity.
template <int D>
__device__ __forceinline__ void legal_words(const EnvParams& p, u64 summed, u64 acted, u64 (&mw)[3]) {
    typedef Lat<D> L;
    const u64 lq = qubits_grid_to_compact<D>(qubits_adjacent_to<D>(summed) | qubits_neighbours_of<D>(acted));
    mw[0] = p.idw[0]; mw[1] = p.idw[1]; mw[2] = p.idw[2];      // the identity (a run-time index here would push mw into local memory)
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        if (l < p.layers) {
            const int o = l * L::NQ, i = o >> 6, s = o & 63;
            mw[i] |= lq << s;
            if (s && i + 1 < 3) mw[i + 1] |= lq >> (64 - s);
        }
    }
}

// The tile's observation bytes.  Thread per 32-bit word of the tile's observation bit stream: expanded to 32 bytes of
// 0/1, two 128-bit stores (the tile's 32*C*H*H bytes start 16-byte aligned whenever the caller's buffer is).  Executed by
// `nthr` threads, t = 0..nthr-1.
__device__ __noinline__ void write_observations_unaligned(const Smem& sm, uint8_t* out, int full, int align, int t, int nthr) {
    for (int g = t * 32; g < full; g += nthr * 32) {        // rare: 64-bit stores at 8-byte alignment, else bytes
        const u32 word = sm.stream[g >> 5];
        if (align == 8) {
            const uint4 a = expand16(sm, word), b = expand16(sm, word >> 16);
            *reinterpret_cast<uint2*>(out + g) = make_uint2(a.x, a.y);
            *reinterpret_cast<uint2*>(out + g + 8) = make_uint2(a.z, a.w);
            *reinterpret_cast<uint2*>(out + g + 16) = make_uint2(b.x, b.y);
            *reinterpret_cast<uint2*>(out + g + 24) = make_uint2(b.z, b.w);
        } else {
#pragma unroll 4
            for (int b = 0; b < 32; ++b) out[g + b] = (uint8_t)((word >> b) & 1u);
        }
    }
}
__device__ __forceinline__ void write_observations(const Smem& sm, const EnvParams& p, uint8_t* obs, int env0, int nvalid, int t, int nthr) {
    const int vbytes = nvalid * p.obs_bits;
    uint8_t* out = obs + (size_t)env0 * p.obs_bits;
    const int align = (int)(reinterpret_cast<uintptr_t>(out) & 15);
    const int full = vbytes & ~31;                                                // whole 32-byte groups
    if (align == 0) {
#pragma unroll 2
        for (int g = t * 32; g < full; g += nthr * 32) {
            const u32 word = sm.stream[g >> 5];
            store_obs16(out + g, expand16(sm, word));
            store_obs16(out + g + 16, expand16(sm, word >> 16));
        }
    } else {
        write_observations_unaligned(sm, out, full, align, t, nthr);
    }
    if (t == 0 && full < vbytes) {                                                 // the tile's last, partial group
        const u32 word = sm.stream[full >> 5];
        for (int b = 0; b < vbytes - full; ++b) out[full + b] = (uint8_t)((word >> b) & 1u);
    }
}

// Words [w0, w1) of the tile's stream again (16-byte aligned tiles only): the span of a lattice that drew a fresh volume in a launch
// whose observations were expanded BEFORE the step (see `early` in env_step_kernel).
__device__ __forceinline__ void rewrite_observation_word(const Smem& sm, uint8_t* out, int vbytes, int w) {
    const int g = w * 32, full = vbytes & ~31;
    if (g >= vbytes) return;
    const u32 word = sm.stream[w];
    if (g < full) {
        store_obs16(out + g, expand16(sm, word));
        store_obs16(out + g + 16, expand16(sm, word >> 16));
    } else {
        for (int b = 0; b < vbytes - full; ++b) out[full + b] = (uint8_t)((word >> b) & 1u);
    }
}

template <int D, bool RESET>
__global__ void __launch_bounds__(Roles<D>::kThreads, DQ_MIN_BLOCKS)
env_step_kernel(const EnvParams p, const int32_t* __restrict__ actions, uint8_t* __restrict__ obs0,
                float* __restrict__ reward0, uint8_t* __restrict__ done0, int32_t* __restrict__ lifetime0,
                u64* __restrict__ legal0, int auto_reset, u32* __restrict__ policy_ctr, int32_t* __restrict__ actions_out0,
                const Rollout ro) {
    typedef Lat<D> L;
    constexpr u32 FULL = 0xffffffffu;
    constexpr int PW = L::PW, H = L::H;
    constexpr int kWriters = Roles<D>::kWriters, kGens = Roles<D>::kGens, kThreads = Roles<D>::kThreads;
    constexpr int kWriterThreads = Roles<D>::kWriterThreads, kOwned = Roles<D>::kOwned;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int env0 = blockIdx.x * kLpc;
    const int C = p.vd + p.layers;
    const int nvalid = max(0, min(kLpc, p.n - env0));
    const size_t np = (size_t)p.npad;
    // A rollout keeps the tile's volume queues in shared memory for the launch; a single step pops and refills the device-memory copy in place.
    const bool q_in_smem = ro.nsteps > 1;
    // A single-step launch is one pass of the physics warp's dependent chain, and nothing of that chain needs the expansion table or the
    // bit stream: there the physics warp (warp 0 of the default role placement) leaves the table and the placement to the other warps,
    // ARRIVES at the prologue barrier instead of waiting at it, and starts its chain at once.
    const bool phys_skip = (DQ_ROTATE == 0) && (DQ_PHYS_LAST == 0) && !q_in_smem;
    const int ptid = phys_skip ? tid - 32 : tid, pthreads = phys_skip ? kThreads - 32 : kThreads;
    // ---- prologue, part 1 (no global memory: may run while the previous launch is still finishing): the expansion table
    pdl_launch_dependents();
    if (ptid >= 0)
        for (int i = ptid; i < 256; i += pthreads)
            sm.lut8[i] = (u64)((((u32)i & 0xFu) * 0x00204081u) & 0x01010101u) | ((u64)((((u32)i >> 4) * 0x00204081u) & 0x01010101u) << 32);
    pdl_wait();
    // built-in policy: every CTA reads the step index before its first barrier; the last CTA to finish advances it
    const u32 step0 = policy_ctr ? *reinterpret_cast<volatile u32*>(policy_ctr) : 0u;
    u64* const gq = p.queue + (size_t)blockIdx.x * (kQ * kQW * kLpc);
    volatile u64* const qp = q_in_smem ? &sm.q[0][0][0] : gq;             // (the generators' stores go through this generic pointer)

    // ---- prologue, part 2: layer bitmaps, queue and its counters
    // A single-step launch would end with the whole tile's expansion AFTER the chain.  Most of those bytes do not depend on the step (a
    // light step changes ONE byte of its lattice's observation): the writer warps expand the stream as it stands at launch while the
    // chain runs, patch that byte when the record arrives, and only the lattices that drew a fresh volume are expanded again at the end.
    uint8_t* const obs_tile0 = obs0 ? obs0 + (size_t)ro.first_slot * ro.slot_bytes + (size_t)env0 * p.obs_bits : nullptr;
    const bool early = !RESET && ro.nsteps == 1 && obs0 != nullptr && (reinterpret_cast<uintptr_t>(obs_tile0) & 15) == 0;
    if (q_in_smem) {
        u64* sq = &sm.q[0][0][0];
        for (int i = tid; i < kQ * kQW * kLpc; i += kThreads) sq[i] = gq[i];
    }
    // the bit stream as the state stands at launch: every (lattice, layer) bitmap row into its place, by every thread of the CTA (a
    // single-step launch would otherwise wait for its few writer threads to do this before they can look at the step's record)
    if (obs0 && ptid >= 0) {
        for (int i = ptid; i < kLpc * C; i += pthreads) {
            const int l = i / C, c = i - l * C;
            u64 w[PW];
#pragma unroll
            for (int j = 0; j < PW; ++j) w[j] = p.state[(ROW_BM + c * PW + j) * np + env0 + l];
            place_layer<D>(sm.stream, l * p.obs_bits + c * L::P, w);
        }
    }
    if (tid < kLpc) {
        const u32 head = (u32)(p.state[ROW_META * np + env0 + tid] >> 32) & 0x7FFFFFFFu;
        u32 tail = p.qtail[env0 + tid];
        if (p.q_reset || tail - head > (u32)kQ) tail = head;      // nothing usable queued (first launch, injected state, new noise rates)
        sm.head[tid] = head; sm.tail[tid] = tail;
        if (tid == 0) {
            sm.quit = 0;
            for (int gg = 0; gg < kGens; ++gg) sm.gsleep[gg] = 0;
            // Warp w of a CTA runs on scheduler (w mod 4) of its SM when CTAs are multiples of four warps, so without a rotation
            // every physics warp of an SM -- the one dependent chain that sets the step time -- would share ONE scheduler with
            // its siblings of the co-resident CTAs.  The k-th CTA to start on an SM rotates its roles by k.
#if DQ_ROTATE
            sm.rot = atomicAdd(&p.sm_arrivals[sm_id() % kMaxSms], 1u) % (u32)(kThreads / 32);
#else
            sm.rot = 0;
#endif
        }
    }
    if (phys_skip) {
        if (warp == 0) { __threadfence_block(); __syncwarp(); bar_arrive_named(0, kThreads); }      // counters and flags are published; no wait
        else bar_sync_named(0, kThreads);
    } else {
        __syncthreads();
    }
#if DQ_PHYS_LAST
    const int role = (kThreads / 32) - 1 - warp;                                   // the physics warp is the CTA's last warp
#else
    const int role = (warp + (kThreads / 32) - (int)sm.rot) % (kThreads / 32);     // 0 physics, 1..kWriters writers, then generators
#endif

    if (role == 0) {
        // ================================================================== PHYSICS: lane = lattice
        const int e = env0 + lane;
        const bool live = lane < nvalid;
        const bool policy = !RESET && policy_ctr != nullptr;
        const u32 env_id = p.env_id_base + (u32)e;
        volatile u32* const vtail = sm.tail;
        volatile u32* const vhead = sm.head;
        u64 xb = 0, zb = 0, act[3] = {0, 0, 0};
        const u64 meta0 = p.state[ROW_META * np + e];
        u64 summed = p.state[ROW_SUM * np + e];
        u32 life = (u32)meta0, attempts = (u32)(meta0 >> 32) & 0x7FFFFFFFu, dn = (u32)(meta0 >> 63);
        if (!RESET) {
            xb = p.state[ROW_XB * np + e];
            zb = p.state[ROW_ZB * np + e];
#pragma unroll
            for (int l = 0; l < 3; ++l) if (l < p.layers) act[l] = p.state[(ROW_ACT + l) * np + e];
        }
        u64 mw[3] = {0, 0, 0};              // legal-move mask after the latest step (feeds the next built-in pick)
        if (policy) legal_words<D>(p, summed, act[0] | act[1] | act[2], mw);

        // ---- the first half of a step (Environments.py:131-150): action -> Pauli frame, true syndrome, homology label, and the
        // referee's class -- a table gather whose L2 latency is the longest single wait of the chain.  With the built-in policy the
        // next step's pick only needs this step's legal mask, so this half runs at the END of the previous iteration and the
        // gather is in flight while that iteration stores its outputs and the next one pops its heavy volumes; its result
        // (a_cls) is first looked at when `done` is decided.
        bool a_heavy = false;
        int a_actbit = -1, a_label = 0, a_cls = 0;
        float a_rw = 0.f;
        auto first_half = [&](int a) {
            if (a < 0 || a >= p.A) a = p.A - 1;
            const bool ident = (a == p.A - 1);
            const int layer = ident ? 0 : a / L::NQ, q = ident ? 0 : a % L::NQ;
            const int qr = q / D, qc = q - qr * D;
            const u64 bit = ident ? 0ull : (1ull << (q + qr));
            const u64 cur = layer == 0 ? act[0] : (layer == 1 ? act[1] : act[2]);
            a_heavy = ident || (cur & bit) != 0;
            // Pauli applied by this action layer (Function_Library.py:253-304)
            bool fx, fz;
            if (p.model == DQ_MODEL_X) { fx = true; fz = false; }
            else if (p.use_y) { fx = layer <= 1; fz = layer >= 1; }
            else { fx = layer == 0; fz = layer == 1; }
            if (fx) xb ^= bit;
            if (fz) zb ^= bit;
            const u64 syn = true_syndrome<D>(xb, zb);
            a_label = homology_label<D>(xb, zb);
            a_rw = (a_label == 0 && syn == 0) ? 1.f : 0.f;
            a_cls = a_label;
            if (a_rw == 0.f && live) a_cls = referee_class<D>(p, syn);
            a_actbit = -1;
            if (!a_heavy) {
                if (layer == 0) act[0] |= bit; else if (layer == 1) act[1] |= bit; else act[2] |= bit;
                a_actbit = (layer << 16) | ((2 * qr + 1) * H + 2 * qc + 1);
            }
        };
        // built-in random-legal policy: the pick dq_policy_random_legal would make on this lattice's current legal set
        auto pick_action = [&](u32 word_u) {
            const int c0 = popc64(mw[0]), c1 = popc64(mw[1]);
            const int cnt = c0 + c1 + popc64(mw[2]);
            int pick = (int)mulhi32(word_u, (u32)cnt);              // < cnt (cnt >= 1: the identity is always legal)
            u64 word = mw[0];
            int wbase = 0;
            if (pick >= c0 + c1) { pick -= c0 + c1; word = mw[2]; wbase = 128; }
            else if (pick >= c0) { pick -= c0; word = mw[1]; wbase = 64; }
            return cnt > 0 ? wbase + select64(word, pick) : p.A - 1;
        };
        if (policy) {
            const int a = live ? pick_action(philox4x32_10(env_id, step0, 0u, 1u, p.k0, p.k1).x) : p.A - 1;
            if (actions_out0 && live) actions_out0[e] = a;
            first_half(a);
        }

        // one pass of the volume loop (Environments.py:158-170): every lane that still needs a volume pops one queued attempt
        u32 todo = 0;
        int guard = 0;
        bool restarted = false;
        u64 sum_new = 0;
        int32_t life_out = 0;
        auto pop_pass = [&](int r) {
            if (todo) {
                if (todo == 2u && !restarted) { xb = 0; zb = 0; life = 0; dn = 0; restarted = true; }     // a restart begins on a clean frame
                if (vtail[lane] == attempts) {                        // (rare) the generators have not queued this attempt yet
                    const long long t0 = clock64();
                    while (vtail[lane] == attempts) {
                        spin_pause();
                        if (clock64() - t0 > kSpinTrapClocks) __trap();   // a lost hand-off must end the launch with an error, not hang the GPU
                    }
                }
                __threadfence_block();
                const int slot = (int)(attempts & (u32)(kQ - 1));
                const u64 s0 = true_syndrome<D>(xb, zb);
                u64 nz = 0, ex, ez;
                if (q_in_smem) {                                      // rollout: plain shared-memory loads
#pragma unroll
                    for (int j = 0; j < kMaxVd; ++j) {
                        if (j < p.vd) {
                            const u64 fj = sm.q[slot][j][lane] ^ s0;
                            nz |= fj;
                            sm.rec_f[r][j][lane] = fj;
                        }
                    }
                    ex = sm.q[slot][kMaxVd][lane]; ez = sm.q[slot][kMaxVd + 1][lane];
                } else {                                              // single step: the device-memory copy, refilled in place by this CTA's generators
                    const volatile u64* const ent = gq + (size_t)slot * (kQW * kLpc) + lane;
#pragma unroll
                    for (int j = 0; j < kMaxVd; ++j) {
                        if (j < p.vd) {
                            const u64 fj = ent[j * kLpc] ^ s0;
                            nz |= fj;
                            sm.rec_f[r][j][lane] = fj;
                        }
                    }
                    ex = ent[kMaxVd * kLpc]; ez = ent[(kMaxVd + 1) * kLpc];
                }
                xb ^= ex;
                zb ^= ez;
                life += (u32)p.vd;
                attempts += 1;
                if (nz != 0 || ++guard >= p.max_attempts) {
                    sum_new = nz;
                    guard = 0;
                    if (todo & 1u) { life_out = (int32_t)life; todo &= ~1u; }
                    else todo = 0;
                }
                __threadfence_block();                                // the entry has been read before its slot is offered for refill
                vhead[lane] = attempts;
            }
        };
        // Doorbells: a generator warp that finds all its queues full announces it (gsleep[g] = 1), looks once more, and sleeps in
        // `bar.sync BAR_DOOR+g, 64`.  The physics warp rings after EVERY pass that popped (a lane may need the refill within this very
        // step): whoever clears gsleep[g] decides -- if the physics warp does, it arrives (32 threads) and the generator's sync completes,
        // now or when it gets there; if the generator does (it found work on its second look), nobody arrives and it does not sync.
        // Arrivals and syncs therefore pair one to one, and a pop made after the generator's last look always finds the flag set.
        auto ring_doorbells = [&]() {
            __threadfence_block();
#pragma unroll
            for (int gg = 0; gg < kGens; ++gg) {
                u32 was = 0;
                if (lane == 0) was = atomicCAS(&sm.gsleep[gg], 1u, 0u);
                if (__shfl_sync(FULL, was, 0)) bar_arrive_named(BAR_DOOR + gg, 64);
            }
        };

        size_t oo = 0;
        for (int rs = 0; rs < ro.nsteps; ++rs, oo += ro.out_stride) {
            const int r = rs & (kRing - 1);
            // the next pick's random word depends on (lattice, step index) only: drawn beside this step's dependent chain
            u32 pick_next = 0;
            if (policy && rs + 1 < ro.nsteps) pick_next = philox4x32_10(env_id, step0 + (u32)rs + 1u, 0u, 1u, p.k0, p.k1).x;
            if (!RESET && !policy) first_half((live && actions) ? actions[e] : p.A - 1);

            // the record slot of this step must have been read by the writers (they run at most kRing steps behind)
            if (rs >= kRing) bar_sync_named(BAR_EMPTY + r, 32 + kWriterThreads);

            // ---- fresh volume(s): bit 0 of todo = heavy step (volume on the current frame), bit 1 = restart of a finished lattice
            // (volume on a clean frame), in that order.  The heavy volumes do not depend on the referee, so their first pass runs
            // before `done` is decided.
            life_out = (int32_t)life;
            restarted = false;
            guard = 0;
            u32 dn_out = 0;
            float rw = 0.f;
            int actbit = -1;
            bool heavy = false;
            if (!RESET) {
                heavy = live && a_heavy;
                rw = a_rw; actbit = a_actbit;
                todo = heavy ? 1u : 0u;
                if (__any_sync(FULL, todo != 0)) { pop_pass(r); ring_doorbells(); }
                if (live && a_cls != a_label) dn = 1;                    // Environments.py:150 (a rewarded step cannot end the episode: a_cls == a_label there)
                dn_out = dn;
                if (live && dn && auto_reset) todo |= 2u;
            } else {
                todo = live ? 2u : 0u;    // reset keeps only the attempt counter (the position in the random stream)
            }
            const bool vol = heavy || todo != 0;
            while (__any_sync(FULL, todo != 0)) { pop_pass(r); ring_doorbells(); }
            if (vol) { act[0] = 0; act[1] = 0; act[2] = 0; summed = sum_new; }
            sm.rec_actbit[r][lane] = vol ? -1 : actbit;
            const u32 vm = __ballot_sync(FULL, vol);
            if (lane == 0) sm.rec_volmask[r] = vm;
            __threadfence_block();
            bar_arrive_named(BAR_FULL + r, 32 + kWriterThreads);

            // ---- legal mask; with the built-in policy the next step's pick and first half (its referee gather goes out here)
            if (legal0 || policy) legal_words<D>(p, summed, act[0] | act[1] | act[2], mw);
            const int32_t life_now = life_out;
            u64 lw0 = mw[0], lw1 = mw[1], lw2 = mw[2];
            if (policy && rs + 1 < ro.nsteps) {
                const int a = live ? pick_action(pick_next) : p.A - 1;
                if (actions_out0 && live) actions_out0[oo + ro.out_stride + e] = a;
                first_half(a);
            }
            // ---- the small outputs of this step
            if (live) {
                if (!RESET) {
                    if (reward0) reward0[oo + e] = rw;
                    if (done0) done0[oo + e] = (uint8_t)dn_out;
                    if (lifetime0) lifetime0[oo + e] = life_now;
                }
                if (legal0) {
                    legal0[(oo + e) * p.W] = lw0;
                    if (p.W > 1) legal0[(oo + e) * p.W + 1] = lw1;
                    if (p.W > 2) legal0[(oo + e) * p.W + 2] = lw2;
                }
            }
        }
        // the lattice's state goes back to its rows
        p.state[ROW_XB * np + e] = xb;
        p.state[ROW_ZB * np + e] = zb;
        p.state[ROW_META * np + e] = meta_pack(life, attempts, dn);
        p.state[ROW_SUM * np + e] = summed;
#pragma unroll
        for (int l = 0; l < 3; ++l) if (l < p.layers) p.state[(ROW_ACT + l) * np + e] = act[l];
        __threadfence_block();
        if (lane == 0) *reinterpret_cast<volatile u32*>(&sm.quit) = 1u;
        __syncwarp();
        ring_doorbells();                    // sleeping generators wake up and leave
    } else if (role <= kWriters) {
        // ================================================================== WRITERS
        const int t = (role - 1) * 32 + lane;
        int ring_slot = ro.first_slot;
        if (early) write_observations(sm, p, obs0 + (size_t)ro.first_slot * ro.slot_bytes, env0, nvalid, t, kWriterThreads);
        for (int rs = 0; rs < ro.nsteps; ++rs) {
            const int r = rs & (kRing - 1);
            uint8_t* const obs = obs0 ? obs0 + (size_t)ring_slot * ro.slot_bytes : nullptr;
            bar_sync_named(BAR_FULL + r, 32 + kWriterThreads);
            const u32 vm = sm.rec_volmask[r];
            // (1) apply the step's record to the state's bitmap rows and to the stream: a light step lights one more cell of its action layer ...
            if (t < kLpc) {
                const int ab = sm.rec_actbit[r][t];
                if (ab >= 0) {
                    const int pos = ab & 0xFFFF, row = (p.vd + (ab >> 16)) * PW + (pos >> 6);
                    atomicOr(reinterpret_cast<unsigned long long*>(&p.state[(ROW_BM + row) * np + env0 + t]), 1ull << (pos & 63));
                    if (obs0) {
                        const int sb = t * p.obs_bits + (p.vd + (ab >> 16)) * L::P + pos;
                        atomicOr(&sm.stream[sb >> 5], 1u << (sb & 31));
                        if (early) obs_tile0[sb] = 1;                     // byte i of the tile's observations is bit i of the stream
                    }
                }
            }
            // ... a fresh volume re-renders the lattice: thread = (lattice, layer); action layers are cleared
            {
                const int items = __popc(vm) * C;
                for (int i = t; i < items; i += kWriterThreads) {
                    const int k = i / C, c = i - k * C;
                    const int slot = select64((u64)vm, k);
                    u64 w[PW];
                    syndrome_layer_bitmap<D>(sm.rec_f[r][min(c, kMaxVd - 1)][slot], w);
#pragma unroll
                    for (int j = 0; j < PW; ++j) {
                        if (c >= p.vd) w[j] = 0ull;                          // an action layer has no marker cells
                        p.state[(ROW_BM + c * PW + j) * np + env0 + slot] = w[j];
                    }
                    if (obs0) place_layer<D>(sm.stream, slot * p.obs_bits + c * L::P, w);
                }
            }
            if (rs + kRing < ro.nsteps) bar_arrive_named(BAR_EMPTY + r, 32 + kWriterThreads);     // this thread is done with the record
            bar_sync_named(BAR_WRITERS, kWriterThreads);
            // (2) the tile's observation bytes (the LAST step's are expanded by the whole CTA after the roles have met, see below)
            if (obs && rs + 1 < ro.nsteps) write_observations(sm, p, obs, env0, nvalid, t, kWriterThreads);
            ring_slot = (ring_slot + 1 == ro.slots) ? 0 : ring_slot + 1;
        }
    } else {
        // ================================================================== GENERATORS
        const int g = role - 1 - kWriters;
        u32* const acc = sm.acc[g];
        volatile u32* const vtail = sm.tail;
        volatile u32* const vhead = sm.head;
        const int mine = lane * kGens + g;                        // lanes < kOwned look at one owned lattice each
        bool announced = false;
        for (;;) {
            u32 quit = *reinterpret_cast<volatile u32*>(&sm.quit);
            __threadfence_block();                                 // quit is read BEFORE the heads it was published after
            quit = __shfl_sync(FULL, quit, 0);                     // one value for the warp: the lanes leave the loop together
            // The launch ends with the physics warp: queues that are not full stay that way until the next launch, whose generators
            // fill them beside its physics warp (DQ_QDEPTH attempts of slack per lattice).  Topping them up here instead would make
            // every single-step launch wait for the generator warp with the most pops to replace.
            if (quit) break;
            u32 key = 0;                                           // (emptiness << 8) | (255 - lane): the emptiest queue first, then the lowest lattice
            if (lane < kOwned && mine < nvalid) {
                const u32 fill = vtail[mine] - vhead[mine];
                if (fill < (u32)kQ) key = (((u32)kQ - fill) << 8) | (u32)(255 - lane);
            }
            key = __reduce_max_sync(FULL, key);                    // one REDUX instruction
            if (key == 0) {                                        // every queue of this warp is full
                if (!announced) {                                  // announce, then look once more (a pop may have slipped in between)
                    if (lane == 0) *reinterpret_cast<volatile u32*>(&sm.gsleep[g]) = 1u;
                    __threadfence_block();
                    __syncwarp();
                    announced = true;
                    continue;
                }
                bar_sync_named(BAR_DOOR + g, 64);                  // sleep until the physics warp pops again (or ends); it cleared the flag
                announced = false;
                continue;
            }
            if (announced) {                                       // work turned up on the second look: take the announcement back ...
                u32 was = 0;
                if (lane == 0) was = atomicCAS(&sm.gsleep[g], 1u, 0u);
                if (!__shfl_sync(FULL, was, 0)) bar_sync_named(BAR_DOOR + g, 64);      // ... unless the physics warp already answered it: take its arrival
                announced = false;
            }
            const int l = (255 - (int)(key & 0xFFu)) * kGens + g;
            const u32 t = vtail[l];
            generate_attempt<D>(p, acc, lane, p.env_id_base + (u32)(env0 + l), t,
                                qp + (size_t)(t & (u32)(kQ - 1)) * (kQW * kLpc) + l);
            __threadfence_block();
            __syncwarp();                                          // every lane's words of the entry are written ...
            if (lane == 0) vtail[l] = t + 1u;                      // ... before the entry is offered
            __syncwarp();                                          // and no lane samples the counters of the next pass before that
        }
    }
    __syncthreads();
    // ---- the last step's observation bytes, by every thread: the other roles have nothing left to do, and a single-step launch ends
    //      when its expansion does
    if (early) {
        const u32 vm = sm.rec_volmask[0];
        const int wpl = (p.obs_bits + 31) / 32 + 1, items = __popc(vm) * wpl, vbytes = nvalid * p.obs_bits;
        for (int i = tid; i < items; i += kThreads) {
            const int k = i / wpl, slot = select64((u64)vm, k);
            const int w = ((slot * p.obs_bits) >> 5) + (i - k * wpl);
            if (w < (((slot + 1) * p.obs_bits + 31) >> 5)) rewrite_observation_word(sm, obs_tile0, vbytes, w);
        }
    } else if (obs0) {
        int last_slot = ro.first_slot + (ro.nsteps - 1) % ro.slots;
        if (last_slot >= ro.slots) last_slot -= ro.slots;
        write_observations(sm, p, obs0 + (size_t)last_slot * ro.slot_bytes, env0, nvalid, tid, kThreads);
    }
    // ---- epilogue: the queues and their fill counters persist between launches
    if (q_in_smem) {
        const u64* sq = &sm.q[0][0][0];
        for (int i = tid; i < kQ * kQW * kLpc; i += kThreads) gq[i] = sq[i];
    }
    if (tid < kLpc) p.qtail[env0 + tid] = sm.tail[tid];
    if (policy_ctr && tid == 0 && atomicAdd(policy_ctr + 1, 1u) == gridDim.x - 1) { policy_ctr[1] = 0; atomicAdd(policy_ctr, (u32)ro.nsteps); }
}

// Uniform pick over the sorted legal actions.  `ctr` (optional) = {step index, finished-CTA count} in
// device memory: when given, the step index is read from it and advanced by the last CTA to finish,
// so the launch can sit in a CUDA graph and still draw fresh words on every replay.
__global__ void policy_random_legal_kernel(const u64* __restrict__ legal, int n, int W, int A, u32 env_id_base,
                                           u32 step, u32* __restrict__ ctr, u32 k0, u32 k1,
                                           int32_t* __restrict__ actions) {
    __shared__ u32 s_step;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    u64 m[3] = {0, 0, 0};
    int cnt = 0;
    if (e < n)
        for (int i = 0; i < W; ++i) { m[i] = legal[(size_t)e * W + i]; cnt += popc64(m[i]); }
    if (ctr) {                                // thread 0's read has landed before anyone passes the barrier
        if (threadIdx.x == 0) s_step = *reinterpret_cast<volatile u32*>(ctr);
        __syncthreads();
        step = s_step;
    }
    if (e < n) {
        const Philox4 u = philox4x32_10(env_id_base + (u32)e, step, 0u, 1u, k0, k1);
        int pick = (int)mulhi32(u.x, (u32)cnt), act = A - 1;
        for (int i = 0; i < W; ++i) {
            const int c = popc64(m[i]);
            if (pick < c) { act = i * 64 + select64(m[i], pick); break; }
            pick -= c;
        }
        actions[e] = act;
    }
    // the last CTA to get here advances the step index: by then every CTA has read it
    if (ctr && threadIdx.x == 0 && atomicAdd(ctr + 1, 1u) == gridDim.x - 1) { ctr[1] = 0; atomicAdd(ctr, 1u); }
}

__global__ void set_u32_kernel(u32* p, u32 v) { p[0] = v; p[1] = 0; }

}  // namespace dq

// ======================================================================== host side / C ABI
using namespace dq;

constexpr int kHostChunks = 4;
struct HostCall {                // outputs of a host-buffer call between its begin and its end
    uint8_t* obs; float* reward; uint8_t* done; int32_t* life; uint64_t* legal;
    int chunks, first[kHostChunks], last[kHostChunks];
    bool pending;
    bool direct;                 // the kernel stored the small outputs straight into the caller's (pinned) buffers
};

struct dq_env {
    EnvParams p;
    int device;
    int state_rows;
    size_t smem_bytes;
    // staging for the *_host entry points
    cudaStream_t hstream;
    u32* policy_ctr;             // {step index, finished-CTA count} for dq_policy_random_legal_next
    int32_t* s_actions; uint8_t* s_obs; float* s_reward; uint8_t* s_done; int32_t* s_life; u64* s_legal;      // s_reward .. s_legal point into s_small
    uint8_t* s_small; uint8_t* h_small; size_t small_bytes;
    int32_t* h_actions_pin; bool zero_copy, direct_ok;
    const void* alias_host[8]; void* alias_dev[8]; int alias_n;          // pinned caller buffers seen so far -> their device addresses      // zero-copy form of the host-buffer calls: the kernel reads the actions from, and writes the small outputs to, pinned host memory
    u64* h_packed;               // pinned landing buffer of the bit-packed observation rows (host-side expansion)
    bool q_reset;                // the next launch must discard the queued volume attempts (noise rates changed)
    cudaEvent_t h_ev[kHostChunks];   // "this lattice range of the rows has landed"
    cudaEvent_t h_ev_small;          // "the small outputs have landed"
    HostCall hc;                 // the host-buffer call in flight (dq_env_step_host_begin .. _end)
};

static std::atomic<long long> g_launches{0};
namespace dq {
thread_local std::string g_err_q;                 // last error message of this thread (shared by all translation units)
void count_launch() { g_launches.fetch_add(1); }
}

static int fail(int code, const std::string& msg) { dq::g_err_q = msg; return code; }
static bool host_expand_enabled();
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
    p.Tmx = p.T > p.Tm ? p.T : p.Tm;
}

extern "C" const char* dq_last_error(void) { return dq::g_err_q.c_str(); }
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
    p.idw[(p.A - 1) >> 6] = 1ull << ((p.A - 1) & 63);
    p.n = (int)n_envs; p.npad = (int)((n_envs + 31) / 32 * 32);
    const int items = volume_depth * (2 * d * d - 1);
    p.rounds = (items + 127) / 128;
    p.obs_bits = (volume_depth + p.layers) * (2 * d + 1) * (2 * d + 1);
    p.ob_magic = (u32)((0x100000000ull + (u64)p.obs_bits - 1) / (u64)p.obs_bits);
    set_thresholds(p, p_phys, p_meas);
    p.k0 = (u32)seed; p.k1 = (u32)(seed >> 32);
    p.env_id_base = (u32)env_id_base;
    p.ref_mode = -1;
    e->device = device;
    e->state_rows = ROW_BM + (volume_depth + p.layers) * (((2 * d + 1) * (2 * d + 1) + 63) / 64);
    e->smem_bytes = (sizeof(Smem) + 15) & ~size_t(15);
    cudaError_t err = cudaMalloc(&p.state, (size_t)e->state_rows * p.npad * sizeof(u64));
    if (err != cudaSuccess) { delete e; return fail(DQ_ECUDA, std::string("cudaMalloc(state): ") + cudaGetErrorString(err)); }
    err = cudaMemset(p.state, 0, (size_t)e->state_rows * p.npad * sizeof(u64));
    if (err != cudaSuccess) { cudaFree(p.state); delete e; return fail(DQ_ECUDA, std::string("cudaMemset(state): ") + cudaGetErrorString(err)); }
    err = cudaMalloc(&e->policy_ctr, 2 * sizeof(u32));
    if (err == cudaSuccess) err = cudaMemset(e->policy_ctr, 0, 2 * sizeof(u32));
    if (err != cudaSuccess) { cudaFree(p.state); delete e; return fail(DQ_ECUDA, std::string("cudaMalloc(policy_ctr): ") + cudaGetErrorString(err)); }
    // volume queues: kQ state-free attempts per lattice, tile-major; qtail = 0 with attempt counters at 0 means "empty"
    const size_t qwords = (size_t)(p.npad / kLpc) * kQ * kQW * kLpc;
    err = cudaMalloc(&p.queue, qwords * sizeof(u64));
    if (err == cudaSuccess) err = cudaMemset(p.queue, 0, qwords * sizeof(u64));
    if (err == cudaSuccess) err = cudaMalloc(&p.qtail, (size_t)p.npad * sizeof(u32));
    if (err == cudaSuccess) err = cudaMemset(p.qtail, 0, (size_t)p.npad * sizeof(u32));
    if (err == cudaSuccess) err = cudaMalloc(&p.sm_arrivals, kMaxSms * sizeof(u32));
    if (err == cudaSuccess) err = cudaMemset(p.sm_arrivals, 0, kMaxSms * sizeof(u32));
    if (err != cudaSuccess) {
        cudaFree(p.queue); cudaFree(p.qtail); cudaFree(p.sm_arrivals); cudaFree(e->policy_ctr); cudaFree(p.state); delete e;
        return fail(DQ_ECUDA, std::string("cudaMalloc(volume queues): ") + cudaGetErrorString(err));
    }
    p.max_attempts = kMaxAttemptsPerCall;
    *out = e;
    return DQ_OK;
}

extern "C" int dq_env_destroy(dq_env* e) {
    if (!e) return DQ_OK;
    DeviceGuard g(e->device);
    if (e->hstream) {
        cudaStreamSynchronize(e->hstream);
        cudaFree(e->s_actions); cudaFree(e->s_obs); cudaFree(e->s_small);
        if (e->h_small) cudaFreeHost(e->h_small);
        if (e->h_actions_pin) cudaFreeHost(e->h_actions_pin);
        if (e->h_packed) cudaFreeHost(e->h_packed);
        for (int c = 0; c < kHostChunks; ++c) if (e->h_ev[c]) cudaEventDestroy(e->h_ev[c]);
        if (e->h_ev_small) cudaEventDestroy(e->h_ev_small);
        cudaStreamDestroy(e->hstream);
    }
    cudaFree(e->p.state);
    cudaFree(e->p.queue);
    cudaFree(e->p.qtail);
    cudaFree(e->p.sm_arrivals);
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
        case DQ_INFO_HOST_EXPAND: *out = host_expand_enabled() ? 1 : 0; break;
        default: return fail(DQ_EINVAL, "unknown info selector");
    }
    return DQ_OK;
}

extern "C" int dq_env_set_noise(dq_env* e, double p_phys, double p_meas) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    if (!(p_phys >= 0.0) || !(p_meas >= 0.0)) return fail(DQ_EINVAL, "probabilities must be >= 0");
    set_thresholds(e->p, p_phys, p_meas);
    e->q_reset = true;            // queued attempts were thresholded with the old rates
    return DQ_OK;
}

extern "C" int dq_env_set_max_attempts(dq_env* e, int max_attempts) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    if (max_attempts < 1 || max_attempts > kMaxAttemptsPerCall) return fail(DQ_EINVAL, "max_attempts must be in [1, 2^20]");
    e->p.max_attempts = max_attempts;
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
                      u64* legal, int auto_reset, cudaStream_t st, u32* pctr = nullptr, int32_t* aout = nullptr,
                      Rollout ro = Rollout{1, 1, 0, 0, 0}) {
    EnvParams p = e->p;
    p.q_reset = e->q_reset ? 1 : 0;
    e->q_reset = false;
    const dim3 grid(p.npad / kLpc);
    // single-step launches carry the programmatic-dependent-launch attribute (pdl_wait in the kernel): back-to-back steps overlap one
    // launch's scheduling and table build with the previous launch's tail.  DQ_ENV_PDL=0 turns it off.
    static const bool pdl_on = [] { const char* v = getenv("DQ_ENV_PDL"); return !(v && v[0] == '0'); }();
    const bool pdl = pdl_on && ro.nsteps == 1;
#ifdef DQ_EMU
#define DQ_LAUNCH_ENV(DD) env_step_kernel<DD, RESET><<<grid, dim3(Roles<DD>::kThreads), e->smem_bytes, st>>>(p, actions, obs, reward, done, lifetime, legal, auto_reset, pctr, aout, ro)
#else
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.dynamicSmemBytes = e->smem_bytes; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
#define DQ_LAUNCH_ENV(DD) (cfg.blockDim = dim3(Roles<DD>::kThreads), (void)cudaLaunchKernelEx(&cfg, env_step_kernel<DD, RESET>, p, actions, obs, reward, done, lifetime, legal, auto_reset, pctr, aout, ro))
#endif
    switch (p.d) {
        case 3: DQ_LAUNCH_ENV(3); break;
        case 5: DQ_LAUNCH_ENV(5); break;
        case 7: DQ_LAUNCH_ENV(7); break;
    }
#undef DQ_LAUNCH_ENV
    (void)pdl;
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

extern "C" int dq_env_step_random(dq_env* e, uint8_t* obs, float* reward, uint8_t* done, int32_t* lifetime, uint64_t* legal,
                                  int32_t* actions_out, int auto_reset, dq_stream stream) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    if (e->p.ref_mode < 0) return fail(DQ_ESTATE, "no referee set: call dq_env_set_referee_lut first");
    DeviceGuard g(e->device);
    return launch_env<false>(e, nullptr, obs, reward, done, lifetime, (u64*)legal, auto_reset, (cudaStream_t)stream, e->policy_ctr, actions_out);
}

extern "C" int dq_env_rollout_random(dq_env* e, int n_steps, uint8_t* obs_ring, int ring_slots, int first_slot, float* reward,
                                     uint8_t* done, int32_t* lifetime, uint64_t* legal, int32_t* actions_out, int auto_reset,
                                     dq_stream stream) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    if (e->p.ref_mode < 0) return fail(DQ_ESTATE, "no referee set: call dq_env_set_referee_lut first");
    if (n_steps < 1 || n_steps > 65536) return fail(DQ_EINVAL, "n_steps must be in [1, 65536]");
    if (obs_ring && (ring_slots < 1 || first_slot < 0 || first_slot >= ring_slots)) return fail(DQ_EINVAL, "need ring_slots >= 1 and 0 <= first_slot < ring_slots");
    DeviceGuard g(e->device);
    Rollout ro;
    ro.nsteps = n_steps; ro.slots = obs_ring ? ring_slots : 1; ro.first_slot = obs_ring ? first_slot : 0;
    ro.slot_bytes = (size_t)e->p.n * e->p.obs_bits; ro.out_stride = (size_t)e->p.n;
    return launch_env<false>(e, nullptr, obs_ring, reward, done, lifetime, (u64*)legal, auto_reset, (cudaStream_t)stream, e->policy_ctr, actions_out, ro);
}

static int ensure_staging(dq_env* e) {
    if (e->hstream) return DQ_OK;
    const EnvParams& p = e->p;
    DQ_CUDA(cudaStreamCreateWithFlags(&e->hstream, cudaStreamNonBlocking));
    DQ_CUDA(cudaMalloc(&e->s_actions, (size_t)p.n * 4));
    // the small outputs of a step sit in ONE device block [legal | reward | lifetime | done] so that one copy brings them all back
    const size_t n = (size_t)p.n;
    e->small_bytes = n * p.W * 8 + n * 4 + n * 4 + n;
    DQ_CUDA(cudaMalloc(&e->s_small, e->small_bytes));
    DQ_CUDA(cudaMallocHost(&e->h_small, e->small_bytes));
    // Zero-copy (default; DQ_HOST_ZEROCOPY=0 for the copy form): a step's chain of DMA operations -- actions in, small outputs out --
    // costs more in launch and completion latency than the 64 KB + 270 KB are worth on the bus.  The kernel reads the actions straight
    // from a pinned host buffer (one coalesced PCIe read at the head of the physics chain) and stores reward / done / lifetime / legal
    // masks straight into the pinned block the caller's buffers are filled from (posted writes); only the bitmap rows still travel by copy.
    // When the caller's own buffers are pinned (cudaHostAlloc / cudaHostRegister: the Python binding's are), the kernel reads and writes
    // THEM and nothing is staged at all; DQ_HOST_DIRECT=0 keeps the library's staging block.  (Both switches are read per handle.)
    const char* zv = getenv("DQ_HOST_ZEROCOPY");
    const char* dv = getenv("DQ_HOST_DIRECT");
    const bool zc = !(zv && zv[0] == '0');
    e->zero_copy = zc;
    e->direct_ok = zc && !(dv && dv[0] == '0');
    uint8_t* base = e->s_small;
    if (zc) {
        DQ_CUDA(cudaMallocHost(&e->h_actions_pin, n * 4));
        void* dev_alias = nullptr;
        DQ_CUDA(cudaHostGetDevicePointer(&dev_alias, e->h_small, 0));
        base = static_cast<uint8_t*>(dev_alias);
    }
    e->s_legal = reinterpret_cast<u64*>(base);
    e->s_reward = reinterpret_cast<float*>(base + n * p.W * 8);
    e->s_life = reinterpret_cast<int32_t*>(base + n * p.W * 8 + n * 4);
    e->s_done = reinterpret_cast<uint8_t*>(base + n * p.W * 8 + n * 8);
    for (int c = 0; c < kHostChunks; ++c) DQ_CUDA(cudaEventCreateWithFlags(&e->h_ev[c], cudaEventDisableTiming));
    DQ_CUDA(cudaEventCreateWithFlags(&e->h_ev_small, cudaEventDisableTiming));
    return DQ_OK;
}

// ---- The *_host entry points return byte observations in the caller's host buffer, but what crosses PCIe are the bit-packed
// bitmap rows of the state matrix (7.5x fewer bytes at d = 5; the kernel's byte-expanding writers have nothing to do): they land in
// a pinned buffer chunk by chunk and are expanded to 0/1 bytes by a pool of host threads while the next chunk is still in flight.
// DQ_HOST_EXPAND=0 in the environment selects the plain form instead (the kernel writes bytes, all of them are copied back).
static bool host_expand_enabled() {
    static const bool on = [] { const char* v = getenv("DQ_HOST_EXPAND"); return !(v && v[0] == '0'); }();
    return on;
}

namespace {
// Persistent host workers for the *_host entry points.  A job is `nb` blocks; every participant (the workers and the calling thread)
// owns a contiguous range of them -- the SAME range on every call of the same shape, so the slice of the caller's buffer a thread
// writes stays in that core's cache from step to step -- and steals from the other ranges once its own is done.  A range is one
// 64-bit word [generation : 24 | end : 20 | next : 20] claimed by compare-and-swap: a worker that wakes up late holds a stale
// generation and can never take a block of the job that followed.  The caller waits for BLOCKS (a counter), not for threads: a worker
// that is slow to wake costs nothing.
struct HostPool {
    static constexpr int kMaxParts = 64;
    struct alignas(64) Range { std::atomic<uint64_t> w{0}; };
    std::vector<std::thread> th;
    std::mutex m, run_m;
    std::condition_variable cv_job;
    const std::function<void(int)>* job = nullptr;
    Range range[kMaxParts];
    int parts = 1;
    std::atomic<uint32_t> generation{0};
    alignas(64) std::atomic<int> done{0};
    alignas(64) std::atomic<int> sleepers{0};
    explicit HostPool(int workers) {
        workers = std::max(0, std::min(workers, kMaxParts - 1));
        parts = workers + 1;
        for (int i = 0; i < workers; ++i) th.emplace_back([this, i] { run(i + 1); });
        for (auto& t : th) t.detach();            // the pool lives as long as the process
    }
    static void relax() {
#if defined(__x86_64__)
        _mm_pause();
#endif
    }
    static uint64_t pack(uint32_t gen, int end, int next) { return ((uint64_t)(gen & 0xFFFFFFu) << 40) | ((uint64_t)end << 20) | (uint64_t)next; }
    // claim one block of generation `gen` from range r; -1: none left there (or the range belongs to a later job)
    int claim(int r, uint32_t gen) {
        uint64_t cur = range[r].w.load(std::memory_order_acquire);
        for (;;) {
            if ((uint32_t)(cur >> 40) != (gen & 0xFFFFFFu)) return -1;
            const int end = (int)((cur >> 20) & 0xFFFFFu), next = (int)(cur & 0xFFFFFu);
            if (next >= end) return -1;
            if (range[r].w.compare_exchange_weak(cur, cur + 1, std::memory_order_acq_rel, std::memory_order_acquire)) return next;
        }
    }
    void work(int self, uint32_t gen, const std::function<void(int)>& f) {
        for (int k = 0; k < parts; ++k) {         // own range first, then the neighbours'
            const int r = (self + k) % parts;
            for (int b; (b = claim(r, gen)) >= 0;) {
                f(b);
                done.fetch_add(1, std::memory_order_acq_rel);
            }
        }
    }
    void run(int self) {
        uint32_t seen = 0;
        for (;;) {
            // a step arrives every ~100 us while a host-buffer loop runs: spin briefly for the next job, then sleep
            int spins = 0;
            while (generation.load(std::memory_order_acquire) == seen) {
                if (++spins < 4000) { relax(); continue; }
                std::unique_lock<std::mutex> lk(m);
                sleepers.fetch_add(1, std::memory_order_acq_rel);
                cv_job.wait(lk, [&] { return generation.load(std::memory_order_acquire) != seen; });
                sleepers.fetch_sub(1, std::memory_order_acq_rel);
            }
            seen = generation.load(std::memory_order_acquire);
            const std::function<void(int)>* f = job;          // published before the generation it belongs to (or a later one: then no claim succeeds)
            work(self, seen, *f);
        }
    }
    void parallel_for(int nb, const std::function<void(int)>& f) {      // the caller works too; one job at a time
        if (nb <= 0) return;
        if (nb == 1 || parts == 1 || nb >= (1 << 20)) { for (int b = 0; b < nb; ++b) f(b); return; }
        std::lock_guard<std::mutex> serial(run_m);
        const uint32_t gen = generation.load(std::memory_order_relaxed) + 1;
        job = &f;
        done.store(0, std::memory_order_relaxed);
        for (int r = 0; r < parts; ++r) {
            const int b0 = (int)((long long)nb * r / parts), b1 = (int)((long long)nb * (r + 1) / parts);
            range[r].w.store(pack(gen, b1, b0), std::memory_order_release);
        }
        generation.store(gen, std::memory_order_release);
        if (sleepers.load(std::memory_order_acquire) > 0) {
            { std::lock_guard<std::mutex> lk(m); }
            cv_job.notify_all();
        }
        work(0, gen, f);
        while (done.load(std::memory_order_acquire) != nb) relax();
    }
};

static int usable_cpus() {
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) { const int c = CPU_COUNT(&set); if (c > 0) return c; }
#endif
    return std::max(1, (int)std::thread::hardware_concurrency());
}

// Threads that work on a host-side job, the caller included.  DQ_HOST_THREADS sets it (the Python binding derives it from
// LOCAL_WORLD_SIZE so that the ranks of one box share the cores); default: the CPUs this process may run on, minus one that is left
// to the rest of the process (interpreter threads, the driver's own) when there are eight or more.
HostPool& host_pool() {
    static HostPool* pool = [] {
        int n = usable_cpus();
        if (n >= 8) n -= 1;
        if (const char* v = getenv("DQ_HOST_THREADS")) n = atoi(v);
        return new HostPool(std::max(0, std::min(n, 32) - 1));
    }();
    return *pool;
}

// byte -> its 8 bits as 8 bytes of 0/1 (little-endian: bit 0 first)
const u64* byte_lut() {
    static const u64* t = [] {
        u64* x = new u64[256];
        for (int i = 0; i < 256; ++i) { u64 v = 0; for (int b = 0; b < 8; ++b) v |= (u64)((i >> b) & 1) << (8 * b); x[i] = v; }
        return x;
    }();
    return t;
}
}  // namespace

#if defined(__x86_64__)
// 32 bits -> 32 bytes of 0/1 in one store
__attribute__((target("avx2"))) static inline void expand32_avx2(uint32_t word, uint8_t* out) {
    __m256i v = _mm256_set1_epi32((int)word);
    const __m256i shuf = _mm256_setr_epi64x(0x0000000000000000LL, 0x0101010101010101LL, 0x0202020202020202LL, 0x0303030303030303LL);
    v = _mm256_shuffle_epi8(v, shuf);                                      // byte k of the word over bytes 8k .. 8k+7
    const __m256i bitm = _mm256_set1_epi64x((long long)0x8040201008040201ULL);
    v = _mm256_and_si256(_mm256_cmpeq_epi8(_mm256_and_si256(v, bitm), bitm), _mm256_set1_epi8(1));
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out), v);
}
__attribute__((target("avx2"))) static void expand_lattices_avx2(const u64* packed, size_t npad, int e0, int e1, int C, int PW, int P, uint8_t* obs) {
    // a layer's last cells leave as one more 32-byte store that ends exactly at the layer's end (it overlaps the previous store with
    // the same values): no write past the layer, no partial copy through a temporary
    const int t0 = P - 32, tw = t0 >> 6, ts = t0 & 63;
    for (int e = e0; e < e1; ++e) {
        uint8_t* out = obs + (size_t)e * C * P;
        for (int c = 0; c < C; ++c, out += P) {
            const u64* col = packed + (size_t)(c * PW) * npad + e;
            int i = 0;
            for (; i + 32 <= P; i += 32) expand32_avx2((uint32_t)(col[(size_t)(i >> 6) * npad] >> (i & 63)), out + i);
            if (i < P) {
                u64 lo = col[(size_t)tw * npad] >> ts;
                if (ts > 32) lo |= col[(size_t)(tw + 1) * npad] << (64 - ts);
                expand32_avx2((uint32_t)lo, out + t0);
            }
        }
    }
}
#endif

#if defined(__x86_64__)
// AVX-512BW: a mask register IS the bit row -- 64 bits -> 64 bytes of 0/1 in one instruction; the layer's last word leaves through a
// byte-masked store (nothing is written past the layer).  A quarter of the AVX2 form's instructions; either way the expansion runs at
// the cores' store bandwidth (measured: building the block's bit stream first and storing whole aligned lines, or non-temporal
// stores, are no faster than this).
__attribute__((target("avx512f,avx512bw"))) static void expand_lattices_avx512(const u64* packed, size_t npad, int e0, int e1, int C, int PW, int P, uint8_t* obs) {
    const __m512i one = _mm512_set1_epi8(1);
    const int tail = P - 64 * (PW - 1);                                   // cells in the layer's last word, 1..64
    const __mmask64 tmask = tail >= 64 ? ~(__mmask64)0 : (((__mmask64)1 << tail) - 1);
    for (int e = e0; e < e1; ++e) {
        uint8_t* out = obs + (size_t)e * C * P;
        const u64* col = packed + e;
        for (int c = 0; c < C; ++c, out += P) {
            for (int w = 0; w + 1 < PW; ++w)
                _mm512_storeu_si512(out + 64 * w, _mm512_maskz_mov_epi8((__mmask64)col[(size_t)(c * PW + w) * npad], one));
            const u64 last = col[(size_t)(c * PW + PW - 1) * npad];
            _mm512_mask_storeu_epi8(out + 64 * (PW - 1), tmask, _mm512_maskz_mov_epi8((__mmask64)last, one));
        }
    }
}
#endif

// packed rows [C*PW][npad] (bit i of layer c of lattice e = bit i%64 of row c*PW + i/64, column e) -> obs [n][C][P] bytes of 0/1,
// for the lattices [first, last)
static void expand_packed_host(const u64* packed, size_t npad, int first, int last, int C, int PW, int P, uint8_t* obs) {
    const u64* lut = byte_lut();
    const int per = 64, nb = (last - first + per - 1) / per;
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("DQ_HOST_NO_AVX2");
    static const bool avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && !getenv("DQ_HOST_NO_AVX512") && !getenv("DQ_HOST_NO_AVX2");
#else
    const bool avx2 = false;
#endif
    const std::function<void(int)> block = [&](int b) {
        const int e0 = first + b * per, e1 = std::min(last, e0 + per);
#if defined(__x86_64__)
        if (avx512) { expand_lattices_avx512(packed, npad, e0, e1, C, PW, P, obs); return; }
        if (avx2 && P >= 32) { expand_lattices_avx2(packed, npad, e0, e1, C, PW, P, obs); return; }
#endif
        for (int e = e0; e < e1; ++e) {
            uint8_t* out = obs + (size_t)e * C * P;
            for (int c = 0; c < C; ++c, out += P)
                for (int w = 0; w < PW; ++w) {
                    u64 v = packed[(size_t)(c * PW + w) * npad + e];
                    const int bits = std::min(64, P - 64 * w);
                    uint8_t* o = out + 64 * w;
                    int i = 0;
                    for (; i + 8 <= bits; i += 8, v >>= 8) memcpy(o + i, &lut[v & 0xFF], 8);
                    for (; i < bits; ++i, v >>= 1) o[i] = (uint8_t)(v & 1);
                }
        }
    };
    host_pool().parallel_for(nb, block);
}

static inline int generic_pick(const u64* m, int W, int A, u32 ux) {      // the pick with the portable helpers
    int cnt = 0;
    for (int w = 0; w < W; ++w) cnt += popc64(m[w]);
    int pick = (int)mulhi32(ux, (u32)cnt);
    for (int w = 0; w < W; ++w) {
        const int c = popc64(m[w]);
        if (pick < c) return w * 64 + select64(m[w], pick);
        pick -= c;
    }
    return A - 1;
}
#if defined(__x86_64__)
// The same picks for `count` consecutive lattices with AVX2 + BMI2 + POPCNT: word 0 of Philox4x32-10(counter = (id, step, 0, 1)) for 8
// stream ids at a time, POPCNT for the counts, PDEP + TZCNT for "the k-th set bit" (the portable build spends ~40 ns per lattice in
// libgcc's table popcount and the branchy select; this one ~9 ns).
#define DQ_X86_POLICY __attribute__((target("avx2,bmi2,popcnt")))
DQ_X86_POLICY static inline __attribute__((always_inline)) void mulhilo_x8(__m256i m, __m256i x, __m256i& hi, __m256i& lo) {      // 8 x (32 x 32 -> 64)
    const __m256i pe = _mm256_mul_epu32(m, x);                                   // even lanes: 64-bit products
    const __m256i po = _mm256_mul_epu32(m, _mm256_srli_epi64(x, 32));            // odd lanes
    lo = _mm256_blend_epi32(pe, _mm256_slli_epi64(po, 32), 0xAA);
    hi = _mm256_blend_epi32(_mm256_srli_epi64(pe, 32), po, 0xAA);
}
DQ_X86_POLICY static void policy_block_x86(const u64* legal, int W, int A, u32 id0, int count, u32 step, u32 key0, u32 key1, int32_t* actions) {
    const __m256i M0 = _mm256_set1_epi32((int)0xD2511F53u), M1 = _mm256_set1_epi32((int)0xCD9E8D57u);
    for (int base = 0; base < count; base += 8) {
        __m256i a = _mm256_add_epi32(_mm256_set1_epi32((int)(id0 + (u32)base)), _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7));
        __m256i b = _mm256_set1_epi32((int)step), c = _mm256_setzero_si256(), d = _mm256_set1_epi32(1);
        u32 k0 = key0, k1 = key1;
        for (int r = 0; r < 10; ++r) {
            __m256i h0, l0, h1, l1;
            mulhilo_x8(M0, a, h0, l0);
            mulhilo_x8(M1, c, h1, l1);
            const __m256i n0 = _mm256_xor_si256(_mm256_xor_si256(h1, b), _mm256_set1_epi32((int)k0));
            const __m256i n2 = _mm256_xor_si256(_mm256_xor_si256(h0, d), _mm256_set1_epi32((int)k1));
            a = n0; b = l1; c = n2; d = l0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        alignas(32) u32 u8[8];
        _mm256_store_si256(reinterpret_cast<__m256i*>(u8), a);
        const int m8 = count - base < 8 ? count - base : 8;
        for (int k = 0; k < m8; ++k) {
            const u64* m = legal + (size_t)(base + k) * W;
            int cnt = 0;
            for (int w = 0; w < W; ++w) cnt += (int)_mm_popcnt_u64(m[w]);
            int pick = (int)(((u64)u8[k] * (u32)cnt) >> 32), act = A - 1;
            for (int w = 0; w < W; ++w) {
                const int cw = (int)_mm_popcnt_u64(m[w]);
                if (pick < cw) { act = w * 64 + __builtin_ctzll(_pdep_u64(1ull << pick, m[w])); break; }
                pick -= cw;
            }
            actions[base + k] = act;
        }
    }
}
#endif

// Uniform pick over the sorted legal actions on the HOST (same draw as dq_policy_random_legal): for callers that drive the
// *_host entry points and hold the legal masks in host memory.
extern "C" int dq_policy_random_legal_host(const dq_env* e, const uint64_t* h_legal, uint32_t step_index, int32_t* h_actions) {
    if (!e || !h_legal || !h_actions) return fail(DQ_EINVAL, "NULL argument");
    const EnvParams& p = e->p;
    const int per = 512, nb = (p.n + per - 1) / per;
#if defined(__x86_64__)
    static const bool x86_fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt") && !getenv("DQ_HOST_NO_AVX2");
#else
    const bool x86_fast = false;
#endif
    const std::function<void(int)> block = [&](int b) {
        const int i0 = b * per, cnt = std::min(p.n, i0 + per) - i0;
#if defined(__x86_64__)
        if (x86_fast) { policy_block_x86(h_legal + (size_t)i0 * p.W, p.W, p.A, p.env_id_base + (u32)i0, cnt, step_index, p.k0, p.k1, h_actions + i0); return; }
#endif
        for (int i = i0; i < i0 + cnt; ++i)
            h_actions[i] = generic_pick(h_legal + (size_t)i * p.W, p.W, p.A, philox4x32_10(p.env_id_base + (u32)i, step_index, 0u, 1u, p.k0, p.k1).x);
    };
    host_pool().parallel_for(nb, block);
    return DQ_OK;
}

// Host-side expansion on its own: the bit-packed rows a *_host_packed call returned -> byte observations [n][C][H][H] of 0/1
// (what Environments.py hands to the agent).  Same code, same host threads, as the expansion inside dq_env_step_host.
extern "C" int dq_unpack_observations_host(const uint64_t* h_packed, int64_t stride, int64_t n, int d, int channels, uint8_t* h_obs) {
    if (!h_packed || !h_obs) return fail(DQ_EINVAL, "NULL argument");
    if (d < 3 || d > 7 || !(d & 1) || channels < 1 || n < 0 || stride < n) return fail(DQ_EINVAL, "need odd d in [3, 7], channels >= 1 and stride >= n >= 0");
    const int side = 2 * d + 1, P = side * side;
    expand_packed_host(h_packed, (size_t)stride, 0, (int)n, channels, (P + 63) / 64, P, h_obs);
    return DQ_OK;
}

// ---- host-buffer calls in two halves: *_begin queues the copy-in, the launch and every copy-out on the handle's own stream and
// returns; *_end waits for the results and (host-side expansion) turns the landed bitmap rows into bytes.  Two handles driven
// begin(A) begin(B) end(A) begin(A) end(B) ... overlap one handle's kernel and copies with the other's expansion on the host.
static int host_chunks() {          // lattice ranges the bitmap rows cross PCIe in (each one costs two driver calls; DQ_HOST_CHUNKS, default 2)
    static const int n = [] { const char* v = getenv("DQ_HOST_CHUNKS"); const int c = v ? atoi(v) : 2; return std::max(1, std::min(c, kHostChunks)); }();
    return n;
}

static int queue_outputs(dq_env* e, bool bytes_on_device, uint64_t* h_packed_user) {
    const EnvParams& p = e->p;
    cudaStream_t s = e->hstream;
    HostCall& hc = e->hc;
    if (e->zero_copy) { /* the kernel has stored them in h_small (or in the caller's pinned buffers) itself */ }
    else if (hc.reward || hc.done || hc.life) DQ_CUDA(cudaMemcpyAsync(e->h_small, e->s_small, e->small_bytes, cudaMemcpyDeviceToHost, s));
    else if (hc.legal) DQ_CUDA(cudaMemcpyAsync(e->h_small, e->s_small, (size_t)p.n * p.W * 8, cudaMemcpyDeviceToHost, s));
    DQ_CUDA(cudaEventRecord(e->h_ev_small, s));
    const size_t rows = (size_t)(e->state_rows - ROW_BM), words = rows * p.npad;
    const u64* src = p.state + (size_t)ROW_BM * p.npad;
    hc.chunks = 0;
    if (h_packed_user) {
        DQ_CUDA(cudaMemcpyAsync(h_packed_user, src, words * sizeof(u64), cudaMemcpyDeviceToHost, s));
    } else if (hc.obs && bytes_on_device) {
        DQ_CUDA(cudaMemcpyAsync(hc.obs, e->s_obs, (size_t)p.n * p.obs_bits, cudaMemcpyDeviceToHost, s));
    } else if (hc.obs) {
        if (!e->h_packed) DQ_CUDA(cudaMallocHost(&e->h_packed, words * sizeof(u64)));
        // lattice ranges of the rows, one event each: the first range is being expanded while the others are still on the bus
        const int nch = p.n >= 4096 ? host_chunks() : 1;
        const int per = ((p.npad / nch) + 31) / 32 * 32;
        for (int c = 0, first = 0; c < nch && first < p.n; ++c, first += per) {
            const int cnt = std::min(per, p.npad - first);
            DQ_CUDA(cudaMemcpy2DAsync(e->h_packed + first, (size_t)p.npad * 8, src + first, (size_t)p.npad * 8, (size_t)cnt * 8, rows,
                                      cudaMemcpyDeviceToHost, s));
            DQ_CUDA(cudaEventRecord(e->h_ev[c], s));
            hc.first[c] = first; hc.last[c] = std::min(p.n, first + cnt);
            hc.chunks = c + 1;
        }
    }
    hc.pending = true;
    return DQ_OK;
}

// device address of a caller buffer if it is pinned host memory, else NULL (looked up once per buffer)
static void* pinned_alias(dq_env* e, const void* h) {
    if (!h || !e->direct_ok) return nullptr;
    for (int i = 0; i < e->alias_n; ++i) if (e->alias_host[i] == h) return e->alias_dev[i];
    void* dev = nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type == cudaMemoryTypeHost) dev = at.devicePointer;
    else cudaGetLastError();                                   // (older drivers report unregistered memory as an error)
    if (e->alias_n < 8) { e->alias_host[e->alias_n] = h; e->alias_dev[e->alias_n] = dev; ++e->alias_n; }
    return dev;
}

static int host_begin(dq_env* e, bool reset, const int32_t* h_actions, uint8_t* h_obs, uint64_t* h_packed_user, float* h_reward, uint8_t* h_done,
                      int32_t* h_life, uint64_t* h_legal, int auto_reset) {
    if (e->hc.pending) return fail(DQ_ESTATE, "a host-buffer call is already in flight on this handle: call dq_env_step_host_end first");
    DeviceGuard g(e->device);
    int rc = ensure_staging(e);
    if (rc) return rc;
    const EnvParams& p = e->p;
    const bool bytes_on_device = h_obs && !host_expand_enabled();
    if (bytes_on_device && !e->s_obs) DQ_CUDA(cudaMalloc(&e->s_obs, (size_t)p.n * p.obs_bits));
    e->hc = HostCall{};
    e->hc.obs = h_obs; e->hc.reward = h_reward; e->hc.done = h_done; e->hc.life = h_life; e->hc.legal = h_legal;
    // small outputs: straight into the caller's buffers when every one of them is pinned, else through the staging block
    float* d_reward = e->s_reward; uint8_t* d_done = e->s_done; int32_t* d_life = e->s_life; u64* d_legal = e->s_legal;
    {
        void* ar = pinned_alias(e, h_reward); void* ad = pinned_alias(e, h_done); void* al = pinned_alias(e, h_life); void* ag = pinned_alias(e, h_legal);
        if ((!h_reward || ar) && (!h_done || ad) && (!h_life || al) && (!h_legal || ag) && (h_reward || h_done || h_life || h_legal)) {
            d_reward = static_cast<float*>(ar); d_done = static_cast<uint8_t*>(ad); d_life = static_cast<int32_t*>(al); d_legal = static_cast<u64*>(ag);
            e->hc.direct = true;
        }
    }
    if (reset) {
        rc = launch_env<true>(e, nullptr, bytes_on_device ? e->s_obs : nullptr, nullptr, nullptr, nullptr, h_legal ? d_legal : nullptr, 1, e->hstream);
    } else {
        const int32_t* d_actions = e->s_actions;
        if (void* aa = pinned_alias(e, h_actions)) {
            d_actions = static_cast<const int32_t*>(aa);                   // the caller must leave it alone until the call's end, as with any async copy
        } else if (e->zero_copy) {
            memcpy(e->h_actions_pin, h_actions, (size_t)p.n * 4);          // (no launch of this handle is in flight: hc.pending was false)
            void* dev_alias = nullptr;
            DQ_CUDA(cudaHostGetDevicePointer(&dev_alias, e->h_actions_pin, 0));
            d_actions = static_cast<const int32_t*>(dev_alias);
        } else {
            DQ_CUDA(cudaMemcpyAsync(e->s_actions, h_actions, (size_t)p.n * 4, cudaMemcpyHostToDevice, e->hstream));
        }
        rc = launch_env<false>(e, d_actions, bytes_on_device ? e->s_obs : nullptr, h_reward ? d_reward : nullptr,
                               h_done ? d_done : nullptr, h_life ? d_life : nullptr, h_legal ? d_legal : nullptr, auto_reset, e->hstream);
    }
    if (rc) return rc;
    return queue_outputs(e, bytes_on_device, h_packed_user);
}

static int host_end(dq_env* e) {
    if (!e->hc.pending) return fail(DQ_ESTATE, "no host-buffer call in flight on this handle");
    DeviceGuard g(e->device);
    const EnvParams& p = e->p;
    HostCall& hc = e->hc;
    hc.pending = false;
    const int side = 2 * p.d + 1, P = side * side;
    DQ_CUDA(cudaEventSynchronize(e->h_ev_small));
    if (!hc.direct) {
        const size_t n = (size_t)p.n;
        if (hc.legal) memcpy(hc.legal, e->h_small, n * p.W * 8);
        if (hc.reward) memcpy(hc.reward, e->h_small + n * p.W * 8, n * 4);
        if (hc.life) memcpy(hc.life, e->h_small + n * p.W * 8 + n * 4, n * 4);
        if (hc.done) memcpy(hc.done, e->h_small + n * p.W * 8 + n * 8, n);
    }
    for (int c = 0; c < hc.chunks; ++c) {
        DQ_CUDA(cudaEventSynchronize(e->h_ev[c]));
        expand_packed_host(e->h_packed, (size_t)p.npad, hc.first[c], hc.last[c], p.vd + p.layers, (P + 63) / 64, P, hc.obs);
    }
    DQ_CUDA(cudaStreamSynchronize(e->hstream));
    return DQ_OK;
}

extern "C" int dq_env_step_host_begin(dq_env* e, const int32_t* h_actions, uint8_t* h_obs, float* h_reward, uint8_t* h_done,
                                      int32_t* h_life, uint64_t* h_legal, int auto_reset) {
    if (!e || !h_actions) return fail(DQ_EINVAL, "env / actions is NULL");
    if (e->p.ref_mode < 0) return fail(DQ_ESTATE, "no referee set: call dq_env_set_referee_lut first");
    return host_begin(e, false, h_actions, h_obs, nullptr, h_reward, h_done, h_life, h_legal, auto_reset);
}

extern "C" int dq_env_step_host_end(dq_env* e) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    return host_end(e);
}

extern "C" int dq_env_reset_host(dq_env* e, uint8_t* h_obs, uint64_t* h_legal) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    const int rc = host_begin(e, true, nullptr, h_obs, nullptr, nullptr, nullptr, nullptr, h_legal, 1);
    return rc ? rc : host_end(e);
}

extern "C" int dq_env_step_host(dq_env* e, const int32_t* h_actions, uint8_t* h_obs, float* h_reward, uint8_t* h_done,
                                int32_t* h_life, uint64_t* h_legal, int auto_reset) {
    const int rc = dq_env_step_host_begin(e, h_actions, h_obs, h_reward, h_done, h_life, h_legal, auto_reset);
    return rc ? rc : host_end(e);
}

// Host-buffer calls that return the observations PACKED (the rows the Q-network and the replay ring consume: one bit per
// cell, uint64 [C*PW][STATE_STRIDE]) instead of one byte per cell: nothing to expand on either side.
extern "C" int dq_env_reset_host_packed(dq_env* e, uint64_t* h_packed, uint64_t* h_legal) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    const int rc = host_begin(e, true, nullptr, nullptr, h_packed, nullptr, nullptr, nullptr, h_legal, 1);
    return rc ? rc : host_end(e);
}

extern "C" int dq_env_step_host_packed(dq_env* e, const int32_t* h_actions, uint64_t* h_packed, float* h_reward, uint8_t* h_done,
                                       int32_t* h_life, uint64_t* h_legal, int auto_reset) {
    if (!e || !h_actions) return fail(DQ_EINVAL, "env / actions is NULL");
    if (e->p.ref_mode < 0) return fail(DQ_ESTATE, "no referee set: call dq_env_set_referee_lut first");
    const int rc = host_begin(e, false, h_actions, nullptr, h_packed, h_reward, h_done, h_life, h_legal, auto_reset);
    return rc ? rc : host_end(e);
}

extern "C" int dq_env_packed_obs(dq_env* e, uint64_t** dev_rows, int64_t* n_rows, int64_t* stride) {
    if (!e || !dev_rows) return fail(DQ_EINVAL, "NULL argument");
    *dev_rows = e->p.state + (size_t)ROW_BM * e->p.npad;
    if (n_rows) *n_rows = e->state_rows - ROW_BM;
    if (stride) *stride = e->p.npad;
    return DQ_OK;
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
    policy_random_legal_kernel<<<(p.n + 127) / 128, 128, 0, st>>>((const u64*)legal, p.n, p.W, p.A, p.env_id_base, step, ctr,
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
