// dq_env.cu -- sm_100a environment-step kernel + C ABI (include/dq_decoding.h).
//
// One CTA (4 warps) advances a tile of 16 independent lattices.  Per-lattice state is bit-packed into
// uint64 rows of a [row][lattice] matrix that stays L2-resident between steps:
// Pauli-frame planes, action boards, counters, AND the already-rendered (2d+1)^2-cell bitmap of every
// observation layer, so that a step only re-renders what changed.  A launch runs one step (dq_env_step*) or a
// rollout of many (dq_env_rollout_random).  For the whole launch the tile's layer bitmaps are mirrored in shared
// memory, together with the tile's observation BIT STREAM in its final layout (lattice-major concatenation of the
// layer bitmaps, no padding: byte i of the tile's observations is bit i of the stream).  Per step (2 block barriers):
//   A  warp 0, lane = lattice: (built-in random-legal pick,) apply the action to the Pauli frame, true syndrome
//      by shifted XORs, homology label, referee table lookup, reward / done, heavy (identity | repeat) flag.
//      Meanwhile warps 1.. write the PREVIOUS step's observations (D) and, once per group of steps, draw the
//      policy's random words of the next group (they depend on (lattice, step index) only)
//   -- barrier; a light step sets its one new cell in the action-layer bitmap and in the stream
//   B  warp per flagged lattice: draw a fresh syndrome volume (generate_volume) -- Philox4x32-10, one
//      block per lane, draws screened by one min-reduce, fired draws folded into per-slice flip masks, warp
//      prefix-XOR over slices -- then the lane that owns a slice builds its layer bitmap in registers and the
//      lattice's span of the stream is re-gathered
//   -- barrier
//   C  warp 0, lane = lattice: lifetime and legal-move mask (kept in registers for the next step's pick)
//   D  thread per 32-bit word of the stream: expand to 32 bytes of 0/1 through a 256-entry table, two aligned
//      128-bit stores (the tile's 16 observations are one contiguous, 16-byte aligned span of HBM)
// Earlier variants (32-lattice tiles staged by TMA bulk copies; warp-autonomous 4/8-lattice groups) and
// the ncu evidence that led here are summarised in profiles/README.md.
//
// Replaces (reference paths relative to example_notebooks/): Environments.py:99-115 (reset),
// :118-204 (step), :206-235, :238-314 and the Function_Library.py helpers they call.
#include <cuda_runtime.h>
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

#ifndef DQ_THREADS
#define DQ_THREADS 128
#endif
#ifndef DQ_MIN_BLOCKS
#define DQ_MIN_BLOCKS 7          // resident CTAs per SM the register allocation must allow
#endif
#ifndef DQ_EPC
#define DQ_EPC 16
#endif
constexpr int kEpc = DQ_EPC;                       // lattices per CTA: 16*L observation bytes are a multiple of 16 for every L
constexpr int kThreads = DQ_THREADS;
constexpr int kWarps = kThreads / 32;
#ifndef DQ_PREFETCH
#define DQ_PREFETCH 0            // rollouts: warps 1.. draw the flip masks of every lattice's next volume attempt ahead of time
#endif
#ifndef DQ_REFILL
#define DQ_REFILL 2              // ... at most this many lattices per warp and step
#endif
#ifndef DQ_BATCHB
#define DQ_BATCHB 0              // 1: phase B in rounds of two stages: masks drawn warp per lattice, then applied / rendered lane per (lattice, slice)
#endif                           // 2: volumes warp per lattice as in the default build, only their finalisation (state, bitmaps, stream) batched
#ifndef DQ_DEFER
#define DQ_DEFER 0               // 1: work that nothing on the step's dependent chain waits for leaves the chain: the frame / counters / action
#endif                           //    boards of the tile are mirrored in shared memory (phase A reads no global state), and a re-rendered
                                 //    lattice's span of the bit stream is refreshed by the threads that expand it (phase D), not by phase B
                                 // 2: ... and so do the layer bitmaps of a fresh volume: phase B leaves the slices in shared memory, the threads
                                 //    of the next window render them (thread = (lattice, layer)) before they expand the stream
#ifndef DQ_STREAM_OBS
#define DQ_STREAM_OBS 0          // 1: observation bytes leave with evict-first stores (STG.E.EF.128): they are never read back by this
#endif                           //    kernel, and a ring of them streaming through L2 otherwise evicts lines of the 4 MB joint referee
                                 //    table (the v8 capture reads 150 KB of DRAM per step, and warp 0 waits ~1000 cycles per step on that lookup)
constexpr bool kPrefetch = DQ_PREFETCH != 0;
constexpr bool kBatchB = DQ_BATCHB != 0;
#ifndef DQ_MIRROR
#define DQ_MIRROR (DQ_DEFER != 0)   // the shared-memory mirror of the frame / counters / action boards alone (part of DQ_DEFER; also combines with DQ_BATCHB)
#endif
#if DQ_DEFER && !DQ_MIRROR
#error "DQ_DEFER needs DQ_MIRROR"
#endif
constexpr bool kDefer = DQ_DEFER != 0;
static_assert(!(kPrefetch && kBatchB), "DQ_PREFETCH and DQ_BATCHB both use the prepared-mask buffers");
static_assert(!(kDefer && kBatchB), "DQ_DEFER restructures the default phase B");
constexpr int kRefill = DQ_REFILL;
constexpr u32 kNoAttempt = 0xffffffffu;              // attempt indices have 31 bits
constexpr int kPickGroup = (kThreads - 32) / kEpc;   // rollout steps whose policy words warps 1.. draw in one go (thread = (step, lattice))
static_assert(kPickGroup >= 1, "warps 1.. must cover at least one step of the tile");
constexpr int kMaxVd = 8;
constexpr int kMaxLayers = kMaxVd + 3;
constexpr int kStreamWords = (kEpc * kMaxLayers * 15 * 15 + 31) / 32 + 1;   // d = 7 at the deepest volume
// The reference loops until a volume is non-trivial, forever if p_phys = p_meas = 0 on a clean frame.
// A kernel must end: after this many attempts in one call the (trivial) volume is accepted.
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
    u32 ob_magic;                               // ceil(2^32 / obs_bits): g / obs_bits == umulhi(g, ob_magic) for g < 2^16
    u32 T, T1, T2, Tm;                          // thresholds (RNG contract)
    u32 Tmx;                                    // max(T, Tm): the one-compare screen of generate_volume
    u64 idw[3];                                 // legal-mask words with only the identity action (A-1) set
    u32 k0, k1;                                 // Philox key
    u32 env_id_base;
    int ref_mode;
    const uint8_t* lut_a;
    const uint8_t* lut_b;
    u64* state;                                 // [STATE_WORDS][npad]
};

#ifndef DQ_LUT_KEEP
#define DQ_LUT_KEEP 0            // 1: referee-table loads carry an L2 evict-last policy (the table competes with the observation stream for L2)
#endif
#if DQ_LUT_KEEP && !defined(DQ_EMU_DYNAMIC_SMEM)
__device__ __forceinline__ int lut2(const uint8_t* lut, u32 idx) {
    u64 pol;
    u32 v;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(lut + (idx >> 2)), "l"(pol));
    return (int)((v >> ((idx & 3) * 2)) & 3u);
}
#else
__device__ __forceinline__ int lut2(const uint8_t* lut, u32 idx) { return (__ldg(lut + (idx >> 2)) >> ((idx & 3) * 2)) & 3; }
#endif

template <int D>
__device__ __forceinline__ int referee_class(const EnvParams& p, u64 syn) {
    if (p.ref_mode == DQ_REFEREE_JOINT) return lut2(p.lut_a, stabs_grid_to_joint_index<D>(syn));
    int c = lut2(p.lut_a, stabs_grid_to_type_index<D, 1>(syn)) & 1;
    if (p.model == DQ_MODEL_DP && p.lut_b) c |= (lut2(p.lut_b, stabs_grid_to_type_index<D, 0>(syn)) & 1) << 1;
    return c;
}

// A fired draw (rare: p ~ 1e-2 per draw) is folded into the per-slice flip accumulators of the warp,
// acc[kind][slice] as two 32-bit halves, kind 0 = data-qubit X flips, 1 = Z flips, 2 = measurement flips.
template <int D>
__device__ __noinline__ void record_event(u32* acc, int item, u32 uv, int nq_items, int n_items, u32 T1, u32 T2, int dp) {
    typedef Lat<D> L;
    if (item < nq_items) {
        const int j = item / L::NQ, q = item - j * L::NQ, pos = q + q / D;
        const u32 bit = 1u << (pos & 31);
        if (!dp || uv < T2) atomicXor(&acc[((0 * kMaxVd + j) << 1) + (pos >> 5)], bit);     // X or Y
        if (dp && uv >= T1) atomicXor(&acc[((1 * kMaxVd + j) << 1) + (pos >> 5)], bit);     // Y or Z
    } else if (item < n_items) {
        const int mi = item - nq_items, j = mi / L::NS, k = mi - j * L::NS, pos = L::stab_pos(k);
        atomicXor(&acc[((2 * kMaxVd + j) << 1) + (pos >> 5)], 1u << (pos & 31));
    }
}

// The draws of ONE volume attempt of one lattice, executed by a full warp: R rounds of one Philox4x32-10 block per lane
// (draw i = word i/B of block i%B), issued two rounds at a time so two Philox chains overlap.  Every lane thresholds its own
// draws; the few that fire are XOR-ed into the accumulators acc[kind][slice] (record_event).  They depend on (lattice, attempt
// index) only -- not on the lattice's state -- which is what lets warps 1.. draw them ahead of time (DQ_PREFETCH).
// Returns (warp-uniform) whether a data-qubit draw fired.
template <int D>
__device__ __forceinline__ bool draw_flip_masks(const EnvParams& p, u32* acc, int lane, u32 env_id, u32 attempt) {
    typedef Lat<D> L;
    constexpr u32 FULL = 0xffffffffu;
    const int vd = p.vd, R = p.rounds, nq_items = vd * L::NQ, n_items = vd * (L::NQ + L::NS);
    const int dp = p.model == DQ_MODEL_DP;
    if (lane < 3 * kMaxVd) { acc[2 * lane] = 0; acc[2 * lane + 1] = 0; }
    __syncwarp();
    bool evq = false;
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
                if (uv < thr) {
                    record_event<D>(acc, item, uv, nq_items, n_items, p.T1, p.T2, dp);
                    evq |= item < nq_items;
                }
            } while (hits);
        }
    }
    const bool anyq = __any_sync(FULL, evq);
    __syncwarp();
    return anyq;
}

// One volume (Environments.py:158-176 / :216-235) for one lattice, executed by a full warp: attempts until one is
// non-trivial.  An attempt's flip masks come from `pre` when they were drawn ahead for exactly this attempt index (pre_att),
// else they are drawn now into `acc`.  Lane j < vd then owns slice j: a warp prefix-XOR gives the frame after every slice, one
// shifted-XOR syndrome per lane the faulty slices.  Updates xb, zb (frame), life, attempts (all warp-uniform); returns this
// lane's slice.
template <int D>
__device__ __forceinline__ u64 generate_volume(const EnvParams& p, u32* acc, const u32* pre, u32 pre_att, bool pre_anyq,
                                               int lane, u32 env_id, u64& xb, u64& zb, u32& life, u32& attempts) {
    constexpr u32 FULL = 0xffffffffu;
    const int vd = p.vd;
    bool nontrivial;
    u64 f = 0;
    int guard = 0;
    do {
        const u32* masks = acc;
        bool anyq;
        if (pre_att == attempts) { masks = pre; anyq = pre_anyq; }
        else anyq = draw_flip_masks<D>(p, acc, lane, env_id, attempts);
        u64 ex = 0, ez = 0, m = 0;
        if (lane < vd) {
            const u64* a64 = reinterpret_cast<const u64*>(masks);
            ex = a64[0 * kMaxVd + lane]; ez = a64[1 * kMaxVd + lane]; m = a64[2 * kMaxVd + lane];
        }
        __syncwarp();
        if (anyq) {                           // inclusive prefix XOR over slices: frame delta after slice `lane`
#pragma unroll
            for (int off = 1; off < kMaxVd; off <<= 1) {
                const u64 tx = __shfl_up_sync(FULL, ex, off), tz = __shfl_up_sync(FULL, ez, off);
                if (lane >= off) { ex ^= tx; ez ^= tz; }
            }
        }
        const u64 fx = xb ^ ex, fz = zb ^ ez;
        f = (lane < vd) ? (true_syndrome<D>(fx, fz) ^ m) : 0ull;
        nontrivial = __any_sync(FULL, f != 0);
        if (anyq) {
            xb = __shfl_sync(FULL, fx, vd - 1);
            zb = __shfl_sync(FULL, fz, vd - 1);
        }
        life += (u32)vd;
        attempts += 1;
    } while (!nontrivial && ++guard < kMaxAttemptsPerCall);
    return f;
}

struct Rollout {            // multi-step launch: see env_step_kernel
    int nsteps, slots, first_slot;
    size_t slot_bytes;      // bytes between observation slots
    size_t out_stride;      // elements between the per-step rows of reward / done / lifetime / actions (legal: x W)
};

struct Smem {
    u64 bm[kEpc][kMaxLayers * 4];     // [lattice][layer*PW + word]: rendered bitmap of every observation layer (mirror of the state rows)
    u32 stream[kStreamWords];         // the tile's observation bit stream: lattice-major concatenation of the layer bitmaps, P bits each,
                                      // no padding = byte i of the tile's observations is bit i; kept in step with bm, so phase D is a
                                      // straight expansion of consecutive words
    u64 lut8[256];                    // byte -> its 8 bits as 8 bytes of 0/1 (phase D)
    u64 fx[kEpc], fz[kEpc], fmeta[kEpc];   // phase A -> B hand-off: frame planes, counters
    u64 sum[kEpc], acted[kEpc];       // OR of the volume's slices; OR of the action boards
    u32 acc[kWarps][3 * kMaxVd * 2];  // per-warp flip accumulators of generate_volume
    u32 pre[(kPrefetch || kBatchB) ? kEpc : 1][3 * kMaxVd * 2];   // DQ_PREFETCH: flip masks of attempt pre_att[lattice] of each lattice, drawn ahead
    u32 pre_att[kEpc], att[kEpc];     // ... the attempt index they belong to; the lattice's current attempt counter
    uint8_t pre_anyq[kEpc];
    u32 pick_u[2][kPickGroup][kEpc];  // built-in policy: word 0 of the (lattice, step) policy block, drawn a group of steps ahead, double-buffered
    int32_t life_out[kEpc];
    int actbit[kEpc];                 // light step: (action layer << 16) | cell bit to set, else -1
    uint8_t task[kEpc], task_flags[kEpc];
    int ntask;
    int npending[2];                  // DQ_BATCHB: lattices that still need a volume attempt after a round (by round parity)
    u64 fsl[(DQ_BATCHB == 2 || DQ_DEFER == 2) ? kEpc : 1][kMaxVd];   // DQ_BATCHB=2: the slices of a finished volume, handed from its warp to the batched finalisation
#if DQ_MIRROR
    u64 cx[kEpc], cz[kEpc], cmeta[kEpc], cact[3][kEpc];   // the tile's frame planes, counters and action boards (mirror of the state rows)
#endif
#if DQ_DEFER
    u32 dirty[2];                     // by step parity: lattices (bit = slot) whose span of `stream` is older than their bitmaps
#endif
};

// 32 bits of the tile's observation bit stream starting at bit `o` of (lattice, layer): the stream is the concatenation of the
// layer bitmaps (P bits each), lattice-major, so a 32-bit window touches at most two of them (P >= 49).  A lattice's bitmap words
// are contiguous in shared memory, read here as 32-bit words.
template <int D>
__device__ __forceinline__ u32 gather32(const Smem& sm, int lat, int layer, int o, int C) {
    typedef Lat<D> L;
    constexpr int PW = L::PW, P = L::P;
    const u32* w = reinterpret_cast<const u32*>(&sm.bm[lat][layer * PW]) + (o >> 5);
    u32 v = __funnelshift_r(w[0], w[1], o & 31);     // w[1] may belong to the next layer: those bits are masked off below
    const int n1 = P - o;                             // bits left in this layer (bits >= P of a bitmap are zero)
    if (n1 < 32) {
        v &= (1u << n1) - 1u;
        int l2 = layer + 1, lat2 = lat;
        if (l2 == C) { l2 = 0; lat2 = lat + 1; }
        if (lat2 < kEpc) v |= reinterpret_cast<const u32*>(&sm.bm[lat2][l2 * PW])[0] << n1;
    }
    return v;
}

// word `wi` of the tile's observation bit stream, gathered from the layer bitmaps
template <int D>
__device__ __forceinline__ u32 stream_word(const Smem& sm, const EnvParams& p, int wi, int C) {
    const int g = wi * 32;
    const int lat = (int)__umulhi((u32)g, p.ob_magic), r = g - lat * p.obs_bits;
    const int layer = r / Lat<D>::P, o = r - layer * Lat<D>::P;
    return lat < kEpc ? gather32<D>(sm, lat, layer, o, C) : 0u;
}

__device__ __forceinline__ uint4 expand16(const Smem& sm, u32 h) {     // low 16 bits -> 16 bytes of 0/1 (two table lookups)
    const uint2 a = *reinterpret_cast<const uint2*>(&sm.lut8[h & 0xFFu]);
    const uint2 b = *reinterpret_cast<const uint2*>(&sm.lut8[(h >> 8) & 0xFFu]);
    return make_uint4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ void store_obs16(uint8_t* dst, const uint4 v) {
#if DQ_STREAM_OBS
    __stcs(reinterpret_cast<uint4*>(dst), v);
#else
    *reinterpret_cast<uint4*>(dst) = v;
#endif
}

#if DQ_DEFER == 2
// Barrier over the `count` threads of the CTA that name barrier `id` (1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_barrier(int id, int count) {
#ifdef DQ_EMU_DYNAMIC_SMEM
    dq_emu::named_barrier(id, count);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

// DQ_DEFER=2: the layer bitmaps of the lattices in dirty mask `dm` (fresh volumes of the step being written), one (lattice, layer)
// pair per thread: syndrome layers from the slices phase B left in fsl, action layers cleared.  Executed by threads
// [0, nthr) of the caller's group; the caller synchronises the group before anyone reads the bitmaps.
template <int D>
__device__ __forceinline__ void render_dirty(Smem& sm, const EnvParams& p, u32 dm, int t, int nthr, int env0, int C) {
    typedef Lat<D> L;
    constexpr int PW = L::PW;
    const size_t np = (size_t)p.npad;
    const int items = __popc(dm) * C;
    for (int i = t; i < items; i += nthr) {
        const int k = i / C, c = i - k * C;
        const int slot = select64((u64)dm, k);
        const int e = env0 + slot;
        u64 w[PW];
        syndrome_layer_bitmap<D>(sm.fsl[slot][min(c, kMaxVd - 1)], w);
#pragma unroll
        for (int j = 0; j < PW; ++j) {
            const u64 v = c < p.vd ? w[j] : 0ull;               // an action layer has no marker cells
            sm.bm[slot][c * PW + j] = v;
            p.state[(ROW_BM + c * PW + j) * np + e] = v;
        }
    }
}
#endif

#if DQ_DEFER
// DQ_DEFER: word `wi` of the bit stream as phase D needs it.  Phase B no longer refreshes the span of a lattice it re-rendered; it
// sets the lattice's bit in the step's dirty mask `dm`, and the thread that expands a word overlapping such a lattice gathers
// it from the bitmaps here and puts it back (every word has one owner per pass; the bitmaps are not written during phase D).
template <int D>
__device__ __forceinline__ u32 fresh_word(Smem& sm, const EnvParams& p, int wi, u32 dm, int C) {
    if (dm) {
        const u32 la = __umulhi((u32)(wi * 32), p.ob_magic), lb = min(__umulhi((u32)(wi * 32 + 31), p.ob_magic), (u32)(kEpc - 1));
        if (((dm >> la) | (dm >> lb)) & 1u) {       // la <= lb < kEpc: the word's bits past the tile's last lattice are zero either way
            const u32 v = stream_word<D>(sm, p, wi, C);
            sm.stream[wi] = v;
            return v;
        }
    }
    return sm.stream[wi];
}
#endif

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

// Phase D: the tile's observation bytes.  Thread per 32-bit word of the tile's observation bit stream: expanded to 32 bytes of
// 0/1, two 128-bit stores (the tile's 16*C*H*H bytes start 16-byte aligned whenever the caller's buffer is).  Executed by
// threads [first, first + nthr) of the CTA.
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
#if !DQ_DEFER
__device__ __forceinline__ void write_observations(const Smem& sm, const EnvParams& p, uint8_t* obs, int env0, int nvalid,
                                                   int t, int nthr) {
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
#else
// DQ_DEFER form: the same expansion, every word read through fresh_word (dm = the dirty mask of the step being written)
template <int D>
__device__ __forceinline__ void write_observations(Smem& sm, const EnvParams& p, uint8_t* obs, int env0, int nvalid,
                                                   int t, int nthr, u32 dm, int C) {
    const int vbytes = nvalid * p.obs_bits;
    uint8_t* out = obs + (size_t)env0 * p.obs_bits;
    const int align = (int)(reinterpret_cast<uintptr_t>(out) & 15);
    const int full = vbytes & ~31;                                                // whole 32-byte groups
    if (align == 0) {
#pragma unroll 2
        for (int g = t * 32; g < full; g += nthr * 32) {
            const u32 word = fresh_word<D>(sm, p, g >> 5, dm, C);
            store_obs16(out + g, expand16(sm, word));
            store_obs16(out + g + 16, expand16(sm, word >> 16));
        }
    } else {
        if (dm)                                                                   // same word ownership as the copy loop below
            for (int g = t * 32; g < full; g += nthr * 32) fresh_word<D>(sm, p, g >> 5, dm, C);
        write_observations_unaligned(sm, out, full, align, t, nthr);
    }
    if (t == 0 && full < vbytes) {                                                 // the tile's last, partial group
        const u32 word = fresh_word<D>(sm, p, full >> 5, dm, C);
        for (int b = 0; b < vbytes - full; ++b) out[full + b] = (uint8_t)((word >> b) & 1u);
    }
}
#endif

template <int D, bool RESET>
__global__ void __launch_bounds__(kThreads, DQ_MIN_BLOCKS)
env_step_kernel(const EnvParams p, const int32_t* __restrict__ actions, uint8_t* __restrict__ obs0,
                float* __restrict__ reward0, uint8_t* __restrict__ done0, int32_t* __restrict__ lifetime0,
                u64* __restrict__ legal0, int auto_reset, u32* __restrict__ policy_ctr, int32_t* __restrict__ actions_out0,
                const Rollout ro) {
    typedef Lat<D> L;
    constexpr u32 FULL = 0xffffffffu;
    constexpr int PW = L::PW, H = L::H;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int env0 = blockIdx.x * kEpc;
    const int C = p.vd + p.layers;
    const int nvalid = max(0, min(kEpc, p.n - env0));      // a tile past the last lattice (n padded to 32) has none
    const size_t np = (size_t)p.npad;
    // built-in policy: every CTA reads the step index before its first barrier; the last CTA to finish advances it
    const u32 step0 = policy_ctr ? *reinterpret_cast<volatile u32*>(policy_ctr) : 0u;

    // The rendered layer bitmaps and the summed syndrome of the tile's lattices live in shared memory for the whole launch
    // (and in the state rows, kept in step): a step only re-renders what it changes.
    for (int i0 = 0; i0 < C * PW * kEpc; i0 += 6 * kThreads) {        // the loads of a batch are in flight together
        u64 w6[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int i = min(i0 + j * kThreads + tid, C * PW * kEpc - 1), row = i / kEpc;
            w6[j] = p.state[(ROW_BM + row) * np + env0 + (i - row * kEpc)];
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int i = i0 + j * kThreads + tid, row = i / kEpc;
            if (i < C * PW * kEpc) sm.bm[i - row * kEpc][row] = w6[j];
        }
    }
    if (tid < kEpc) {
        sm.sum[tid] = p.state[ROW_SUM * np + env0 + tid];
        sm.pre_att[tid] = kNoAttempt;
        if (kPrefetch) sm.att[tid] = (u32)(p.state[ROW_META * np + env0 + tid] >> 32) & 0x7FFFFFFFu;
#if DQ_MIRROR
        if (!RESET) {
            sm.cx[tid] = p.state[ROW_XB * np + env0 + tid];
            sm.cz[tid] = p.state[ROW_ZB * np + env0 + tid];
            sm.cmeta[tid] = p.state[ROW_META * np + env0 + tid];
#pragma unroll
            for (int l = 0; l < 3; ++l) sm.cact[l][tid] = l < p.layers ? p.state[(ROW_ACT + l) * np + env0 + tid] : 0ull;
        }
#endif
#if DQ_DEFER
        if (tid < 2) sm.dirty[tid] = 0;
#endif
    }
    // built-in policy: the random word of a pick depends only on (lattice, step index), so it never has to sit on the
    // step's dependent chain: the words of the first kPickGroup steps are drawn here, those of every later group by warps 1..
    // one group ahead (below), into the buffer warp 0 is not reading
    if (!RESET && policy_ctr && tid < kPickGroup * kEpc) {
        const int ahead = tid / kEpc, lat = tid - ahead * kEpc;
        if (ahead < ro.nsteps)
            sm.pick_u[0][ahead][lat] = philox4x32_10(p.env_id_base + (u32)(env0 + lat), step0 + (u32)ahead, 0u, 1u, p.k0, p.k1).x;
    }
    for (int i = tid; i < 256; i += kThreads)
        sm.lut8[i] = (u64)((((u32)i & 0xFu) * 0x00204081u) & 0x01010101u) | ((u64)((((u32)i >> 4) * 0x00204081u) & 0x01010101u) << 32);
    __syncthreads();
    for (int wi = tid; wi < (kEpc * p.obs_bits + 31) >> 5; wi += kThreads) sm.stream[wi] = stream_word<D>(sm, p, wi, C);
    __syncthreads();

    // A rollout (dq_env_rollout_random) runs ro.nsteps steps of this tile's lattices in one launch: lattices are independent,
    // so a tile never waits for the slowest tile of the previous step.  Step s writes observation slot (first_slot+s) % slots
    // and row s of the per-step outputs.  A single step is the nsteps == 1 case.
    int ring_slot = ro.first_slot;
    size_t oo = 0;
    u64 mw[3] = {0, 0, 0};              // warp 0, lane = lattice: legal-move mask after the latest step (phase C -> next phase A)
    uint8_t* obs_prev = nullptr;        // the previous step's observation slot: written while warp 0 runs this step's phase A
    int gphase = 0, gbuf = 0;           // rs % kPickGroup, (rs / kPickGroup) & 1
    for (int rs = 0; rs < ro.nsteps; ++rs, oo += ro.out_stride, ring_slot = (ring_slot + 1 == ro.slots) ? 0 : ring_slot + 1) {
    uint8_t* const obs = obs0 ? obs0 + (size_t)ring_slot * ro.slot_bytes : nullptr;
    float* const reward = reward0 ? reward0 + oo : nullptr;
    uint8_t* const done_out = done0 ? done0 + oo : nullptr;
    int32_t* const lifetime = lifetime0 ? lifetime0 + oo : nullptr;
    int32_t* const actions_out = actions_out0 ? actions_out0 + oo : nullptr;
    u64* const legal = legal0 ? legal0 + oo * p.W : nullptr;

    // ---- phase A: warp 0, lane = lattice.  Warps 1.. meanwhile write the PREVIOUS step's observations (phase D): phase A does
    //      not touch the bitmaps -- the cell a light step adds is applied after the barrier.
    if (warp != 0) {
        if (!RESET && policy_ctr && gphase == 0 && rs + kPickGroup < ro.nsteps && tid - 32 < kPickGroup * kEpc) {
            // policy words of the NEXT group of steps -> the other buffer (last read two barriers ago, next read after this step's barriers)
            const int ahead = (tid - 32) / kEpc, lat = (tid - 32) - ahead * kEpc, s = rs + kPickGroup + ahead;
            if (s < ro.nsteps)
                sm.pick_u[gbuf ^ 1][ahead][lat] = philox4x32_10(p.env_id_base + (u32)(env0 + lat), step0 + (u32)s, 0u, 1u, p.k0, p.k1).x;
        }
#if !DQ_DEFER
        if (obs_prev) write_observations(sm, p, obs_prev, env0, nvalid, tid - 32, kThreads - 32);
#else
        {
            const u32 dm = sm.dirty[(rs & 1) ^ 1];           // the previous step's fresh volumes (0 at the first step)
#if DQ_DEFER == 2
            if (dm) {               // block-uniform
                render_dirty<D>(sm, p, dm, tid - 32, kThreads - 32, env0, C);
                if (obs_prev) named_barrier(1, kThreads - 32);
            }
#endif
            if (obs_prev) write_observations<D>(sm, p, obs_prev, env0, nvalid, tid - 32, kThreads - 32, dm, C);
        }
#endif
        if (kPrefetch && !RESET && ro.nsteps > 1) {
            // draw ahead: lattices whose prepared masks are not those of their next attempt (consumed, or never drawn).  This
            // window (warp 0 runs phase C and A) is separated from phase B, which reads them, by the block barriers.
            int done = 0;
            for (int lat = warp - 1; lat < nvalid && done < kRefill; lat += kWarps - 1) {
                const u32 want = sm.att[lat];
                if (sm.pre_att[lat] == want) continue;                     // warp-uniform
                const bool anyq = draw_flip_masks<D>(p, sm.pre[lat], lane, p.env_id_base + (u32)(env0 + lat), want);
                if (lane == 0) { sm.pre_att[lat] = want; sm.pre_anyq[lat] = anyq ? 1 : 0; }
                ++done;
            }
        }
    } else {
        const int e = env0 + lane;
        const bool mine = lane < kEpc, live = lane < nvalid;
        u64 xb = 0, zb = 0, meta = 0, act[3] = {0, 0, 0};
        u32 flags = 0;
        if (mine) {
#if !DQ_MIRROR
            meta = p.state[ROW_META * np + e];
#else
            meta = RESET ? p.state[ROW_META * np + e] : sm.cmeta[lane];
#endif
            int actbit = -1;
            if (!RESET) {
                int a = (live && actions) ? actions[e] : p.A - 1;
#if !DQ_MIRROR
                xb = p.state[ROW_XB * np + e];
                zb = p.state[ROW_ZB * np + e];
#pragma unroll
                for (int l = 0; l < 3; ++l) if (l < p.layers) act[l] = p.state[(ROW_ACT + l) * np + e];
#else
                xb = sm.cx[lane];
                zb = sm.cz[lane];
#pragma unroll
                for (int l = 0; l < 3; ++l) act[l] = sm.cact[l][lane];
#endif
                if (policy_ctr && live) {
                    // built-in random-legal policy (dq_env_step_random): the pick dq_policy_random_legal would make
                    // on this lattice's current legal set, with the step index read from device memory
                    if (rs == 0) legal_words<D>(p, sm.sum[lane], act[0] | act[1] | act[2], mw);   // later steps: phase C of the previous step left it
                    const int ib = p.A - 1;
                    const int cnt = popc64(mw[0]) + popc64(mw[1]) + popc64(mw[2]);
                    int pick = (int)mulhi32(sm.pick_u[gbuf][gphase][lane], (u32)cnt);              // < cnt (cnt >= 1: the identity is always legal)
                    const int c0 = popc64(mw[0]), c1 = popc64(mw[1]);
                    u64 word = mw[0];
                    int wbase = 0;
                    if (pick >= c0 + c1) { pick -= c0 + c1; word = mw[2]; wbase = 128; }
                    else if (pick >= c0) { pick -= c0; word = mw[1]; wbase = 64; }
                    a = cnt > 0 ? wbase + select64(word, pick) : ib;
                    if (actions_out) actions_out[e] = a;
                }
                if (a < 0 || a >= p.A) a = p.A - 1;
                const bool ident = (a == p.A - 1);
                const int layer = ident ? 0 : a / L::NQ, q = ident ? 0 : a % L::NQ;
                const int qr = q / D, qc = q - qr * D;
                const u64 bit = ident ? 0ull : (1ull << (q + qr));
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
                else if (live && referee_class<D>(p, syn) != label) dn = 1;
                if (!heavy) {
                    if (layer == 0) act[0] |= bit; else if (layer == 1) act[1] |= bit; else act[2] |= bit;
                    actbit = (layer << 16) | ((2 * qr + 1) * H + 2 * qc + 1);
                }
                meta = (meta & ~(1ull << 63)) | ((u64)dn << 63);
                if (live) {
                    if (reward) reward[e] = rw;
                    if (done_out) done_out[e] = (uint8_t)dn;
                    flags = (heavy ? 1u : 0u) | ((dn && auto_reset) ? 2u : 0u);   // padding lattices never draw volumes
                }
                sm.life_out[lane] = (int32_t)(u32)meta;
                if (!flags) {                  // light step: the frame and the touched action board go back now
                    p.state[ROW_XB * np + e] = xb;
                    p.state[ROW_ZB * np + e] = zb;
                    p.state[ROW_META * np + e] = meta;
                    if (!ident) p.state[(ROW_ACT + layer) * np + e] = layer == 0 ? act[0] : (layer == 1 ? act[1] : act[2]);
#if DQ_MIRROR
                    sm.cx[lane] = xb; sm.cz[lane] = zb; sm.cmeta[lane] = meta;
                    sm.cact[0][lane] = act[0]; sm.cact[1][lane] = act[1]; sm.cact[2][lane] = act[2];
#endif
                }
            } else if (live) {
                flags = 2u;               // reset keeps only the attempt counter (the position in the random stream)
            }
            sm.fx[lane] = xb; sm.fz[lane] = zb; sm.fmeta[lane] = meta;
            sm.acted[lane] = act[0] | act[1] | act[2];
            sm.actbit[lane] = flags ? -1 : actbit;
        }
        const u32 tmask = __ballot_sync(FULL, flags != 0);
        if (flags) {
            const int pos = __popc(tmask & ((1u << lane) - 1));
            sm.task[pos] = (uint8_t)lane; sm.task_flags[pos] = (uint8_t)flags;
        }
        if (lane == 0) sm.ntask = __popc(tmask);
    }
    __syncthreads();

#if DQ_DEFER
    if (tid == 0) sm.dirty[(rs & 1) ^ 1] = 0;  // the previous step's mask: phase D consumed it before the barrier above
#endif
    if (tid < kEpc) {                          // a light step's action lights one more cell of its action layer
        const int ab = sm.actbit[tid];
        if (ab >= 0) {
            const int pos = ab & 0xFFFF, row = (p.vd + (ab >> 16)) * PW + (pos >> 6);
            const u64 wv = sm.bm[tid][row] | (1ull << (pos & 63));
            sm.bm[tid][row] = wv;
            p.state[(ROW_BM + row) * np + env0 + tid] = wv;
            const int sb = tid * p.obs_bits + (p.vd + (ab >> 16)) * L::P + pos;
            atomicOr(&sm.stream[sb >> 5], 1u << (sb & 31));
        }
    }
    // ---- phase B: fresh volume(s) for the flagged lattices, then their layer bitmaps and stream spans
    if constexpr (!kBatchB) {
    // warp per flagged lattice
        for (int t = warp; t < sm.ntask; t += kWarps) {
            const int slot = sm.task[t], fl = sm.task_flags[t];
            const u32 env_id = p.env_id_base + (u32)(env0 + slot);
            const int e = env0 + slot;
            u64 bx = sm.fx[slot], bz = sm.fz[slot];
            const u64 bm0 = sm.fmeta[slot];
            u32 life = (u32)bm0, attempts = (u32)(bm0 >> 32) & 0x7FFFFFFFu, dn = (u32)(bm0 >> 63);
            u64 f = 0;
            int32_t lo = (int32_t)life;
            // bit 0 of fl: heavy step (volume on the current frame), bit 1: restart of a finished lattice (volume on a clean frame), in
            // that order; one copy of the generator serves both
#pragma unroll 1
            for (int todo = fl; todo; ) {
                const bool restart = !(todo & 1);
                if (restart) { bx = 0; bz = 0; life = 0; dn = 0; }
                f = generate_volume<D>(p, sm.acc[warp], sm.pre[kPrefetch ? slot : 0], sm.pre_att[slot], sm.pre_anyq[slot] != 0,
                                       lane, env_id, bx, bz, life, attempts);
                if (!restart) lo = (int32_t)life;
                todo = restart ? 0 : (todo & 2);
            }
            u64 summed = f;                                 // lanes >= vd hold 0
            summed |= __shfl_xor_sync(FULL, summed, 1);
            summed |= __shfl_xor_sync(FULL, summed, 2);
            summed |= __shfl_xor_sync(FULL, summed, 4);
            if (lane == 0) {
                p.state[ROW_XB * np + e] = bx;
                p.state[ROW_ZB * np + e] = bz;
                p.state[ROW_META * np + e] = meta_pack(life, attempts, dn);
                p.state[ROW_SUM * np + e] = summed;
                sm.sum[slot] = summed; sm.acted[slot] = 0;
                if (kPrefetch) sm.att[slot] = attempts;
                if (!RESET) sm.life_out[slot] = lo;
#if DQ_MIRROR
                if (!RESET) {
                    sm.cx[slot] = bx; sm.cz[slot] = bz; sm.cmeta[slot] = meta_pack(life, attempts, dn);
                    sm.cact[0][slot] = 0; sm.cact[1][slot] = 0; sm.cact[2][slot] = 0;
                }
#endif
#if DQ_DEFER
                atomicOr(&sm.dirty[rs & 1], 1u << slot);
#endif
            }
#if DQ_DEFER == 2
            if (lane < kMaxVd) sm.fsl[slot][lane] = f;          // lanes >= vd hold 0; rendered in the next window (render_dirty)
#endif
            if (lane < p.layers) p.state[(ROW_ACT + lane) * np + e] = 0;     // not deferred: phase A of the next step may write this row
            // render: lane j < vd holds slice j and builds that layer's bitmap in registers; lanes vd..C-1 clear the action layers
            if (DQ_DEFER != 2 && lane < C) {
                u64 w[PW];
                syndrome_layer_bitmap<D>(f, w);
#pragma unroll
                for (int i = 0; i < PW; ++i) {
                    const u64 v = lane < p.vd ? w[i] : 0ull;
                    sm.bm[slot][lane * PW + i] = v;
                    p.state[(ROW_BM + lane * PW + i) * np + e] = v;
                }
            }
            __syncwarp();
            // this lattice's span of the tile's bit stream; its first and last word are shared with the neighbouring lattices, which
            // other warps may be re-rendering right now: only this lattice's bits of those are replaced, atomically
            // (DQ_DEFER: left to phase D, see fresh_word)
            if constexpr (!kDefer) {
                const int b0 = slot * p.obs_bits, b1 = b0 + p.obs_bits;
                for (int wi = (b0 >> 5) + lane; wi <= ((b1 - 1) >> 5); wi += 32) {
                    const int lo = max(b0 - wi * 32, 0), hi = min(b1 - wi * 32, 32);          // bits [lo, hi) of the word are this lattice's
                    const u32 mask = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                    const u32 v = stream_word<D>(sm, p, wi, C);
                    if (mask == 0xffffffffu) sm.stream[wi] = v;
                    else { atomicAnd(&sm.stream[wi], ~mask); atomicOr(&sm.stream[wi], v & mask); }
                }
            }
        }
    } else if constexpr (DQ_BATCHB == 2) {
    // DQ_BATCHB=2.  Stage 1, warp per flagged lattice: the volume(s) exactly as in the default build (attempts retried inside the
    // warp, no block barrier), but nothing of the few-lane work that follows: the finished slices, frame and counters are left in
    // shared memory.  Stage 2, after one block barrier, lane = (lattice of a group of four, slice): state rows, layer bitmaps,
    // action boards and stream spans of four lattices in one pass.
    for (int t = warp; t < sm.ntask; t += kWarps) {
        const int slot = sm.task[t], fl = sm.task_flags[t];
        const u32 env_id = p.env_id_base + (u32)(env0 + slot);
        u64 bx = sm.fx[slot], bz = sm.fz[slot];
        const u64 bm0 = sm.fmeta[slot];
        u32 life = (u32)bm0, attempts = (u32)(bm0 >> 32) & 0x7FFFFFFFu, dn = (u32)(bm0 >> 63);
        u64 f = 0;
        int32_t lo = (int32_t)life;
#pragma unroll 1
        for (int todo = fl; todo; ) {
            const bool restart = !(todo & 1);
            if (restart) { bx = 0; bz = 0; life = 0; dn = 0; }
            f = generate_volume<D>(p, sm.acc[warp], sm.pre[0], kNoAttempt, false, lane, env_id, bx, bz, life, attempts);
            if (!restart) lo = (int32_t)life;
            todo = restart ? 0 : (todo & 2);
        }
        if (lane < kMaxVd) sm.fsl[slot][lane] = f;               // lanes >= vd hold 0
        if (lane == 0) {
            sm.fx[slot] = bx; sm.fz[slot] = bz; sm.fmeta[slot] = meta_pack(life, attempts, dn);
            if (!RESET) sm.life_out[slot] = lo;
        }
    }
    __syncthreads();
    {
        const int ntask = sm.ntask;
        for (int c = warp; c * 4 < ntask; c += kWarps) {
            const int j = lane >> 3, sl = lane & 7, t = c * 4 + j;
            const bool complete = t < ntask;
            const int slot = complete ? sm.task[t] : 0;
            const int e = env0 + slot;
            const u64 f = complete ? sm.fsl[slot][sl] : 0ull;
            u64 summed = f;
            summed |= __shfl_xor_sync(FULL, summed, 1, 8);
            summed |= __shfl_xor_sync(FULL, summed, 2, 8);
            summed |= __shfl_xor_sync(FULL, summed, 4, 8);
            if (complete) {
                if (sl == 0) {
                    p.state[ROW_XB * np + e] = sm.fx[slot];
                    p.state[ROW_ZB * np + e] = sm.fz[slot];
                    p.state[ROW_META * np + e] = sm.fmeta[slot];
                    p.state[ROW_SUM * np + e] = summed;
                    sm.sum[slot] = summed; sm.acted[slot] = 0;
#if DQ_MIRROR
                    if (!RESET) {
                        sm.cx[slot] = sm.fx[slot]; sm.cz[slot] = sm.fz[slot]; sm.cmeta[slot] = sm.fmeta[slot];
                        sm.cact[0][slot] = 0; sm.cact[1][slot] = 0; sm.cact[2][slot] = 0;
                    }
#endif
                }
                u64 w[PW];
                syndrome_layer_bitmap<D>(f, w);
                if (sl < p.vd) {
#pragma unroll
                    for (int i = 0; i < PW; ++i) {
                        sm.bm[slot][sl * PW + i] = w[i];
                        p.state[(ROW_BM + sl * PW + i) * np + e] = w[i];
                    }
                }
                if (sl < p.layers) {                           // the action boards and their layers are cleared
                    p.state[(ROW_ACT + sl) * np + e] = 0;
#pragma unroll
                    for (int i = 0; i < PW; ++i) {
                        sm.bm[slot][(p.vd + sl) * PW + i] = 0ull;
                        p.state[(ROW_BM + (p.vd + sl) * PW + i) * np + e] = 0ull;
                    }
                }
            }
            __syncwarp();
            // stream spans of the group's lattices as one list of (lattice, word) pairs over the lanes
            const u32 cmask = __ballot_sync(FULL, complete && sl == 0);
            const int nws = ((p.obs_bits + 31) >> 5) + 1;
            for (int base = 0; base < 4 * nws; base += 32) {
                const int idx = base + lane;
                const int jj = (idx >= nws) + (idx >= 2 * nws) + (idx >= 3 * nws), k = idx - jj * nws;
                const int sj = __shfl_sync(FULL, slot, (jj & 3) * 8);
                const int b0 = sj * p.obs_bits, b1 = b0 + p.obs_bits, wi = (b0 >> 5) + k;
                if (idx < 4 * nws && ((cmask >> (jj * 8)) & 1u) && wi <= ((b1 - 1) >> 5)) {
                    const int lo = max(b0 - wi * 32, 0), hi = min(b1 - wi * 32, 32);
                    const u32 mask = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                    const u32 v = stream_word<D>(sm, p, wi, C);
                    if (mask == 0xffffffffu) sm.stream[wi] = v;
                    else { atomicAnd(&sm.stream[wi], ~mask); atomicOr(&sm.stream[wi], v & mask); }
                }
            }
        }
    }
    __syncthreads();
    } else {
    // DQ_BATCHB=1.  Everything after the draws of a volume attempt is a few lanes' work per lattice, so a warp per lattice runs it at
    // 5-8 active lanes.  Here a round is two stages: (1) the flip masks of ONE attempt of every lattice that still needs a volume,
    // warp per lattice (all lanes busy: Philox + thresholds); (2) after a block barrier, lane = (lattice of a group of four,
    // slice): prefix-XOR, syndromes, triviality test, and for the lattices whose volume is complete the state, the layer bitmaps
    // and the stream spans -- one pass for up to four lattices.  Task flags: bit 0 heavy volume pending, bit 1 restart pending,
    // bit 2 restart in progress (its frame was reset).  Rounds repeat while a lattice drew an all-trivial volume or still has
    // its restart to do (block-uniform: counted in npending[round parity]).
    for (int round = 0;; ++round) {
        for (int t = warp; t < sm.ntask; t += kWarps) {
            if (!sm.task_flags[t]) continue;
            const int slot = sm.task[t];
            const u32 att = (u32)(sm.fmeta[slot] >> 32) & 0x7FFFFFFFu;
            const bool anyq = draw_flip_masks<D>(p, sm.pre[slot], lane, p.env_id_base + (u32)(env0 + slot), att);
            if (lane == 0) sm.pre_anyq[slot] = anyq ? 1 : 0;
        }
        if (tid == 0) sm.npending[round & 1] = 0;
        __syncthreads();
        const int ntask = sm.ntask;
        for (int c = warp; c * 4 < ntask; c += kWarps) {
            const int j = lane >> 3, sl = lane & 7, t = c * 4 + j;
            int fl = t < ntask ? sm.task_flags[t] : 0;
            const bool act = fl != 0;
            const int slot = t < ntask ? sm.task[t] : 0;
            const int e = env0 + slot;
            u64 xb = sm.fx[slot], zb = sm.fz[slot];
            const u64 bm0 = sm.fmeta[slot];
            u32 life = (u32)bm0, attempts = (u32)(bm0 >> 32) & 0x7FFFFFFFu, dn = (u32)(bm0 >> 63);
            const bool heavy = (fl & 1) != 0;
            if (act && !heavy && (fl & 2)) { xb = 0; zb = 0; life = 0; dn = 0; fl = 4; }      // the restart begins on a clean frame
            u64 ex = 0, ez = 0, m = 0;
            if (act && sl < p.vd) {
                const u64* a64 = reinterpret_cast<const u64*>(sm.pre[slot]);
                ex = a64[0 * kMaxVd + sl]; ez = a64[1 * kMaxVd + sl]; m = a64[2 * kMaxVd + sl];
            }
#pragma unroll
            for (int off = 1; off < kMaxVd; off <<= 1) {          // inclusive prefix XOR over the slices of each 8-lane group
                const u64 tx = __shfl_up_sync(FULL, ex, off, 8), tz = __shfl_up_sync(FULL, ez, off, 8);
                if (sl >= off) { ex ^= tx; ez ^= tz; }
            }
            const u64 fx = xb ^ ex, fz = zb ^ ez;
            const u64 f = (act && sl < p.vd) ? (true_syndrome<D>(fx, fz) ^ m) : 0ull;
            const u32 bal = __ballot_sync(FULL, f != 0);
            const bool nontrivial = ((bal >> (j * 8)) & 0xFFu) != 0 || round >= kMaxAttemptsPerCall - 1;
            xb = __shfl_sync(FULL, fx, p.vd - 1, 8);
            zb = __shfl_sync(FULL, fz, p.vd - 1, 8);
            life += (u32)p.vd;
            attempts += 1;
            bool complete = false, heavy_done = false;
            if (act && nontrivial) {
                if (heavy) { heavy_done = true; fl &= ~1; complete = fl == 0; }
                else { fl = 0; complete = true; }
            }
            u64 summed = f;
            summed |= __shfl_xor_sync(FULL, summed, 1, 8);
            summed |= __shfl_xor_sync(FULL, summed, 2, 8);
            summed |= __shfl_xor_sync(FULL, summed, 4, 8);
            if (act && sl == 0) {
                sm.fx[slot] = xb; sm.fz[slot] = zb; sm.fmeta[slot] = meta_pack(life, attempts, dn);
                sm.task_flags[t] = (uint8_t)fl;
                if (fl) atomicAdd(&sm.npending[round & 1], 1);
                if (heavy_done && !RESET) sm.life_out[slot] = (int32_t)life;
                if (complete) {
                    p.state[ROW_XB * np + e] = xb;
                    p.state[ROW_ZB * np + e] = zb;
                    p.state[ROW_META * np + e] = meta_pack(life, attempts, dn);
                    p.state[ROW_SUM * np + e] = summed;
                    sm.sum[slot] = summed; sm.acted[slot] = 0;
#if DQ_MIRROR
                    if (!RESET) {
                        sm.cx[slot] = xb; sm.cz[slot] = zb; sm.cmeta[slot] = meta_pack(life, attempts, dn);
                        sm.cact[0][slot] = 0; sm.cact[1][slot] = 0; sm.cact[2][slot] = 0;
                    }
#endif
                }
            }
            if (complete) {
                u64 w[PW];
                syndrome_layer_bitmap<D>(f, w);
                if (sl < p.vd) {
#pragma unroll
                    for (int i = 0; i < PW; ++i) {
                        sm.bm[slot][sl * PW + i] = w[i];
                        p.state[(ROW_BM + sl * PW + i) * np + e] = w[i];
                    }
                }
                if (sl < p.layers) {                           // the action boards and their layers are cleared
                    p.state[(ROW_ACT + sl) * np + e] = 0;
#pragma unroll
                    for (int i = 0; i < PW; ++i) {
                        sm.bm[slot][(p.vd + sl) * PW + i] = 0ull;
                        p.state[(ROW_BM + (p.vd + sl) * PW + i) * np + e] = 0ull;
                    }
                }
            }
            __syncwarp();
            // stream spans of the completed lattices of this group of four, as one list of (lattice, word) pairs over the lanes
            // (a span's first / last word is shared with the neighbouring lattices: atomics)
            {
                const u32 cmask = __ballot_sync(FULL, complete && sl == 0);
                const int nws = ((p.obs_bits + 31) >> 5) + 1;             // words a lattice's span can touch
                for (int base = 0; base < 4 * nws; base += 32) {
                    const int idx = base + lane;
                    const int jj = (idx >= nws) + (idx >= 2 * nws) + (idx >= 3 * nws), k = idx - jj * nws;
                    const int sj = __shfl_sync(FULL, slot, (jj & 3) * 8);
                    const int b0 = sj * p.obs_bits, b1 = b0 + p.obs_bits, wi = (b0 >> 5) + k;
                    if (idx < 4 * nws && ((cmask >> (jj * 8)) & 1u) && wi <= ((b1 - 1) >> 5)) {
                        const int lo = max(b0 - wi * 32, 0), hi = min(b1 - wi * 32, 32);
                        const u32 mask = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                        const u32 v = stream_word<D>(sm, p, wi, C);
                        if (mask == 0xffffffffu) sm.stream[wi] = v;
                        else { atomicAnd(&sm.stream[wi], ~mask); atomicOr(&sm.stream[wi], v & mask); }
                    }
                }
            }
        }
        __syncthreads();
        if (sm.npending[round & 1] == 0) break;
    }
    }
    if constexpr (!kBatchB) __syncthreads();

    // ---- phase C: lifetime and legal mask (warp 0, lane = lattice; the mask also feeds the next step's built-in pick)
    if (tid < kEpc) {
        const int slot = tid;
        if (lifetime && !RESET && slot < nvalid) lifetime[env0 + slot] = sm.life_out[slot];
        if (legal || policy_ctr) {
            legal_words<D>(p, sm.sum[slot], sm.acted[slot], mw);
            if (legal && slot < nvalid) {
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (i < p.W) legal[(size_t)(env0 + slot) * p.W + i] = mw[i];
            }
        }
    }
    obs_prev = obs;
    if (++gphase == kPickGroup) { gphase = 0; gbuf ^= 1; }
    }   // rollout step
#if !DQ_DEFER
    if (obs_prev) write_observations(sm, p, obs_prev, env0, nvalid, tid, kThreads);           // the last step's observations
#else
    {
        const u32 dm = sm.dirty[(ro.nsteps - 1) & 1];
#if DQ_DEFER == 2
        render_dirty<D>(sm, p, dm, tid, kThreads, env0, C);        // the last phase B lies behind barrier 2
        if (obs_prev) __syncthreads();                             // block-uniform
#endif
        if (obs_prev) write_observations<D>(sm, p, obs_prev, env0, nvalid, tid, kThreads, dm, C);
    }
#endif
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

struct dq_env {
    EnvParams p;
    int device;
    int state_rows;
    size_t smem_bytes;
    // staging for the *_host entry points
    cudaStream_t hstream;
    u32* policy_ctr;             // {step index, finished-CTA count} for dq_policy_random_legal_next
    int32_t* s_actions; uint8_t* s_obs; float* s_reward; uint8_t* s_done; int32_t* s_life; u64* s_legal;
    u64* h_packed;               // DQ_HOST_EXPAND: pinned landing buffer of the bit-packed observation rows
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
    *out = e;
    return DQ_OK;
}

extern "C" int dq_env_destroy(dq_env* e) {
    if (!e) return DQ_OK;
    DeviceGuard g(e->device);
    if (e->hstream) {
        cudaStreamSynchronize(e->hstream);
        cudaFree(e->s_actions); cudaFree(e->s_obs); cudaFree(e->s_reward); cudaFree(e->s_done); cudaFree(e->s_life); cudaFree(e->s_legal);
        if (e->h_packed) cudaFreeHost(e->h_packed);
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
        case DQ_INFO_HOST_EXPAND: *out = host_expand_enabled() ? 1 : 0; break;
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
                      u64* legal, int auto_reset, cudaStream_t st, u32* pctr = nullptr, int32_t* aout = nullptr,
                      Rollout ro = Rollout{1, 1, 0, 0, 0}) {
    const EnvParams& p = e->p;
    const dim3 grid(p.npad / kEpc), block(kThreads);
    switch (p.d) {
        case 3: env_step_kernel<3, RESET><<<grid, block, e->smem_bytes, st>>>(p, actions, obs, reward, done, lifetime, legal, auto_reset, pctr, aout, ro); break;
        case 5: env_step_kernel<5, RESET><<<grid, block, e->smem_bytes, st>>>(p, actions, obs, reward, done, lifetime, legal, auto_reset, pctr, aout, ro); break;
        case 7: env_step_kernel<7, RESET><<<grid, block, e->smem_bytes, st>>>(p, actions, obs, reward, done, lifetime, legal, auto_reset, pctr, aout, ro); break;
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
    DQ_CUDA(cudaMalloc(&e->s_obs, (size_t)p.n * p.obs_bits));
    DQ_CUDA(cudaMalloc(&e->s_reward, (size_t)p.n * 4));
    DQ_CUDA(cudaMalloc(&e->s_done, (size_t)p.n));
    DQ_CUDA(cudaMalloc(&e->s_life, (size_t)p.n * 4));
    DQ_CUDA(cudaMalloc(&e->s_legal, (size_t)p.n * p.W * 8));
    return DQ_OK;
}

// ---- DQ_HOST_EXPAND=1 (opt-in, read once per process): the *_host entry points keep their interface (byte observations in the
// caller's host buffer) but move the observations over PCIe as the bit-packed rows of the state matrix (7.5x fewer bytes, the
// kernel's byte-expanding phase skipped) and expand them here, on a small pool of host threads.
static bool host_expand_enabled() {
    static const bool on = [] { const char* v = getenv("DQ_HOST_EXPAND"); return v && v[0] == '1'; }();
    return on;
}

namespace {
struct HostPool {                 // persistent workers: blocks of a job are claimed from an atomic counter
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    const std::function<void(int)>* job = nullptr;
    int nblocks = 0, generation = 0, active = 0;
    std::atomic<int> next{0};
    explicit HostPool(int n) {
        for (int i = 0; i < n; ++i) th.emplace_back([this] { run(); });
        for (auto& t : th) t.detach();            // the pool lives as long as the process
    }
    void run() {
        int seen = 0;
        for (;;) {
            const std::function<void(int)>* f;
            int nb;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return generation != seen; });
                seen = generation; f = job; nb = nblocks;
            }
            for (int b; (b = next.fetch_add(1)) < nb;) (*f)(b);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--active == 0) cv_done.notify_one();
            }
        }
    }
    void parallel_for(int nb, const std::function<void(int)>& f) {      // the caller works too
        {
            std::lock_guard<std::mutex> lk(m);
            job = &f; nblocks = nb; next.store(0); active = (int)th.size(); ++generation;
        }
        cv_job.notify_all();
        for (int b; (b = next.fetch_add(1)) < nb;) f(b);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return active == 0; });
    }
};

HostPool& host_pool() {
    static HostPool* pool = [] {
        int n = (int)std::thread::hardware_concurrency();
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

// packed rows [C*PW][npad] (bit i of layer c of lattice e = bit i%64 of row c*PW + i/64, column e) -> obs [n][C][P] bytes of 0/1
static void expand_packed_host(const u64* packed, size_t npad, int n, int C, int PW, int P, uint8_t* obs) {
    const u64* lut = byte_lut();
    const int per = 256, nb = (n + per - 1) / per;
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("DQ_HOST_NO_AVX2");
#else
    const bool avx2 = false;
#endif
    const std::function<void(int)> block = [&](int b) {
        const int e1 = std::min(n, (b + 1) * per);
#if defined(__x86_64__)
        if (avx2 && P >= 32) { expand_lattices_avx2(packed, npad, b * per, e1, C, PW, P, obs); return; }
#endif
        for (int e = b * per; e < e1; ++e) {
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

static int copy_packed(dq_env* e, uint64_t* h_packed);
// the tail of a *_host call under DQ_HOST_EXPAND: packed rows to the pinned landing buffer, the small outputs, one synchronise, expand
static int copy_out_expanding(dq_env* e, uint8_t* h_obs, float* h_reward, uint8_t* h_done, int32_t* h_life, uint64_t* h_legal);

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
    const bool expand = h_obs && host_expand_enabled();
    rc = launch_env<true>(e, nullptr, (h_obs && !expand) ? e->s_obs : nullptr, nullptr, nullptr, nullptr, h_legal ? e->s_legal : nullptr, 1, e->hstream);
    if (rc) return rc;
    if (expand) return copy_out_expanding(e, h_obs, nullptr, nullptr, nullptr, h_legal);
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
    const bool expand = h_obs && host_expand_enabled();
    rc = launch_env<false>(e, e->s_actions, (h_obs && !expand) ? e->s_obs : nullptr, h_reward ? e->s_reward : nullptr,
                           h_done ? e->s_done : nullptr, h_life ? e->s_life : nullptr, h_legal ? e->s_legal : nullptr,
                           auto_reset, e->hstream);
    if (rc) return rc;
    if (expand) return copy_out_expanding(e, h_obs, h_reward, h_done, h_life, h_legal);
    return copy_out(e, h_obs, h_reward, h_done, h_life, h_legal);
}

// Host-buffer calls that return the observations PACKED (the rows the Q-network and the replay ring consume: one bit per
// cell, uint64 [C*PW][STATE_STRIDE]) instead of one byte per cell: 7.5x fewer bytes over PCIe, and the byte-expanding phase of
// the kernel is skipped altogether.
static int copy_packed(dq_env* e, uint64_t* h_packed) {
    if (!h_packed) return DQ_OK;
    const size_t words = (size_t)(e->state_rows - ROW_BM) * e->p.npad;
    DQ_CUDA(cudaMemcpyAsync(h_packed, e->p.state + (size_t)ROW_BM * e->p.npad, words * sizeof(u64), cudaMemcpyDeviceToHost, e->hstream));
    return DQ_OK;
}

static int copy_out_expanding(dq_env* e, uint8_t* h_obs, float* h_reward, uint8_t* h_done, int32_t* h_life, uint64_t* h_legal) {
    const EnvParams& p = e->p;
    const size_t words = (size_t)(e->state_rows - ROW_BM) * p.npad;
    if (!e->h_packed) DQ_CUDA(cudaMallocHost(&e->h_packed, words * sizeof(u64)));
    int rc = copy_packed(e, e->h_packed);
    if (rc) return rc;
    rc = copy_out(e, nullptr, h_reward, h_done, h_life, h_legal);          // synchronises the stream
    if (rc) return rc;
    const int side = 2 * p.d + 1, P = side * side;
    expand_packed_host(e->h_packed, (size_t)p.npad, p.n, p.vd + p.layers, (P + 63) / 64, P, h_obs);
    return DQ_OK;
}

extern "C" int dq_env_reset_host_packed(dq_env* e, uint64_t* h_packed, uint64_t* h_legal) {
    if (!e) return fail(DQ_EINVAL, "env is NULL");
    DeviceGuard g(e->device);
    int rc = ensure_staging(e);
    if (rc) return rc;
    rc = launch_env<true>(e, nullptr, nullptr, nullptr, nullptr, nullptr, h_legal ? e->s_legal : nullptr, 1, e->hstream);
    if (rc) return rc;
    rc = copy_packed(e, h_packed);
    if (rc) return rc;
    return copy_out(e, nullptr, nullptr, nullptr, nullptr, h_legal);
}

extern "C" int dq_env_step_host_packed(dq_env* e, const int32_t* h_actions, uint64_t* h_packed, float* h_reward, uint8_t* h_done,
                                       int32_t* h_life, uint64_t* h_legal, int auto_reset) {
    if (!e || !h_actions) return fail(DQ_EINVAL, "env / actions is NULL");
    if (e->p.ref_mode < 0) return fail(DQ_ESTATE, "no referee set: call dq_env_set_referee_lut first");
    DeviceGuard g(e->device);
    int rc = ensure_staging(e);
    if (rc) return rc;
    DQ_CUDA(cudaMemcpyAsync(e->s_actions, h_actions, (size_t)e->p.n * 4, cudaMemcpyHostToDevice, e->hstream));
    rc = launch_env<false>(e, e->s_actions, nullptr, h_reward ? e->s_reward : nullptr, h_done ? e->s_done : nullptr,
                           h_life ? e->s_life : nullptr, h_legal ? e->s_legal : nullptr, auto_reset, e->hstream);
    if (rc) return rc;
    rc = copy_packed(e, h_packed);
    if (rc) return rc;
    return copy_out(e, nullptr, h_reward, h_done, h_life, h_legal);
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
