// dq_lattice.cuh -- bit-board geometry of the distance-D rotated surface code.
//
// Everything the environment kernel needs about the lattice is a handful of 64-bit
// masks and shift patterns.  All boards (Pauli-frame planes, action layers, syndromes)
// live on ONE (D+1) x (D+1) grid packed row-major into a uint64 (D <= 7  =>  <= 64 bits):
//     qubit (r,c), 0 <= r,c < D        -> bit r*(D+1) + c        (column D / row D stay empty)
//     plaquette (a,b), 0 <= a,b <= D   -> bit a*(D+1) + b
// so that plaquette (a,b) = parity of qubits (a-1,b-1),(a-1,b),(a,b-1),(a,b) becomes four
// shifted XORs of a plane, with no wrap-around between rows (the empty column absorbs it).
//
// Behavioural spec (reference paths relative to /root/reference/example_notebooks):
//   plaquette presence / type ... Function_Library.py:23-61  (generateSurfaceCodeLattice)
//   syndrome .................... Function_Library.py:162-184
//   homology label .............. Function_Library.py:306-336
//   stabilizer draw order ....... Function_Library.py:186-233 (generate_faulty_syndrome)
//   neighbours / legal moves .... Environments.py:238-271, 349-372
//   observation embedding ....... Environments.py:273-314
//
// Host-compilable (g++) so tests/host_bits_check.cpp can check every helper against the oracle.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DQ_HD __host__ __device__ __forceinline__
#else
#define DQ_HD inline
#endif

namespace dq {

typedef uint64_t u64;
typedef uint32_t u32;

DQ_HD int popc64(u64 x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

template <int D>
struct Lat {
    static constexpr int G = D + 1;          // grid stride
    static constexpr int NQ = D * D;         // data qubits
    static constexpr int NS = D * D - 1;     // stabilizers
    static constexpr int NT = (D * D - 1) / 2;   // stabilizers per type
    static constexpr int H = 2 * D + 1;      // observation side
    static constexpr int P = H * H;          // cells per observation layer
    static constexpr int PW = (P + 63) / 64; // uint64 words per layer bitmap

    static constexpr bool present(int a, int b) {
        return !((a == 0 && b % 2 == 0) || (a == D && b % 2 == 1) ||
                 (b == 0 && a % 2 == 1) || (b == D && a % 2 == 0));
    }
    static constexpr u64 qmask() {
        u64 m = 0;
        for (int r = 0; r < D; ++r) for (int c = 0; c < D; ++c) m |= 1ull << (r * G + c);
        return m;
    }
    // present plaquettes of type 1 ((a+b) even: parity of the Z plane) / type 3 (odd: X plane)
    static constexpr u64 tmask(int odd) {
        u64 m = 0;
        for (int a = 0; a <= D; ++a) for (int b = 0; b <= D; ++b)
            if (present(a, b) && ((a + b) % 2) == odd) m |= 1ull << (a * G + b);
        return m;
    }
    static constexpr u64 col0() { u64 m = 0; for (int r = 0; r < D; ++r) m |= 1ull << (r * G); return m; }
    static constexpr u64 row0() { return (1ull << D) - 1; }

    static constexpr u64 QMASK = qmask();
    static constexpr u64 T1 = tmask(0);
    static constexpr u64 T3 = tmask(1);
    static constexpr u64 COL0 = col0();
    static constexpr u64 ROW0 = row0();

    // grid position of the k-th stabilizer in draw order
    static constexpr int stab_pos(int k) {
        constexpr int nb = (D - 1) / 2;
        if (k < (D - 1) * (D - 1)) return (1 + k / (D - 1)) * G + 1 + k % (D - 1);
        k -= (D - 1) * (D - 1);
        if (k < nb) return 0 * G + 2 * k + 1;            // top
        k -= nb;
        if (k < nb) return D * G + 2 * k + 2;            // bottom
        k -= nb;
        if (k < nb) return (2 * k + 2) * G + 0;          // left
        k -= nb;
        return (2 * k + 1) * G + D;                      // right
    }
    // grid position of the k-th stabilizer of the given parity (odd=1: type 3) in draw order
    static constexpr int type_pos(int odd, int k) {
        int seen = 0;
        for (int i = 0; i < NS; ++i) {
            int p = stab_pos(i);
            if (((p / G + p % G) % 2) == odd) { if (seen == k) return p; ++seen; }
        }
        return 0;
    }
    // constant marker cells of a syndrome layer (Environments.py:280-298), bit x*H+y
    static constexpr u64 marker_word(int w) {
        u64 m = 0;
        for (int x = 0; x < H; ++x) for (int y = 0; y < H; ++y) {
            bool v = ((x == 0 || x == 2 * D) && y % 2 == 1) || ((y == 0 || y == 2 * D) && x % 2 == 1) ||
                     (x % 2 == 1 && y % 2 == 1 && (x + y) % 4 == 0);
            int bit = x * H + y;
            if (v && bit / 64 == w) m |= 1ull << (bit % 64);
        }
        return m;
    }
};

// ---- plane -> plaquette parities --------------------------------------------------------
template <int D> DQ_HD u64 plaquette_parity(u64 plane) {
    constexpr int G = Lat<D>::G;
    u64 t = plane ^ (plane << 1);
    return t ^ (t << G);
}
template <int D> DQ_HD u64 true_syndrome(u64 xb, u64 zb) {
    return (plaquette_parity<D>(zb) & Lat<D>::T1) | (plaquette_parity<D>(xb) & Lat<D>::T3);
}
template <int D> DQ_HD int homology_label(u64 xb, u64 zb) {
    return (popc64(xb & Lat<D>::COL0) & 1) + 2 * (popc64(zb & Lat<D>::ROW0) & 1);
}

// ---- compact (row-major D*D) <-> grid qubit boards -------------------------------------
template <int D> DQ_HD u64 qubits_compact_to_grid(u64 c) {
    u64 g = 0;
#pragma unroll
    for (int r = 0; r < D; ++r) g |= ((c >> (r * D)) & ((1ull << D) - 1)) << (r * (D + 1));
    return g;
}
template <int D> DQ_HD u64 qubits_grid_to_compact(u64 g) {
    u64 c = 0;
#pragma unroll
    for (int r = 0; r < D; ++r) c |= ((g >> (r * (D + 1))) & ((1ull << D) - 1)) << (r * D);
    return c;
}

// ---- draw-order stabilizer index <-> grid ------------------------------------------------
template <int D> DQ_HD u64 stabs_grid_to_compact(u64 s) {
    constexpr int G = D + 1, M = D - 1, nb = (D - 1) / 2;
    u64 c = 0;
#pragma unroll
    for (int a = 1; a < D; ++a) c |= ((s >> (a * G + 1)) & ((1ull << M) - 1)) << ((a - 1) * M);
#pragma unroll
    for (int k = 0; k < 4 * nb; ++k) c |= ((s >> Lat<D>::stab_pos(M * M + k)) & 1ull) << (M * M + k);
    return c;
}
template <int D> DQ_HD u64 stabs_compact_to_grid(u64 c) {
    constexpr int G = D + 1, M = D - 1, nb = (D - 1) / 2;
    u64 s = 0;
#pragma unroll
    for (int a = 1; a < D; ++a) s |= ((c >> ((a - 1) * M)) & ((1ull << M) - 1)) << (a * G + 1);
#pragma unroll
    for (int k = 0; k < 4 * nb; ++k) s |= ((c >> (M * M + k)) & 1ull) << Lat<D>::stab_pos(M * M + k);
    return s;
}
// JOINT referee-table index ("joint order", include/dq_decoding.h): grid rows 1..D-1 are D contiguous
// present plaquettes each (columns 1..D on odd rows, 0..D-1 on even rows); the top and bottom boundary
// rows interleave into columns 1..D-1.  Only present plaquettes are ever set in a true syndrome.
template <int D> DQ_HD u32 stabs_grid_to_joint_index(u64 s) {
    constexpr int G = D + 1;
    static_assert(D * D - 1 <= 32 || D == 7, "joint index must fit 32 bits");
    if (D * D - 1 > 32) return 0;             // d = 7 has no joint table (dq_env_set_referee_lut rejects it)
    u32 idx = 0;
#pragma unroll
    for (int a = 1; a < D; ++a) idx |= ((u32)(s >> (a * G + (a & 1))) & ((1u << D) - 1)) << (((a - 1) * D) & 31);
    const u32 tb = ((u32)s | (u32)(s >> (D * G))) & ((1u << D) - 1);
    return idx | ((tb >> 1) << ((D * (D - 1)) & 31));
}
// index over the stabilizers of one type (ODD=1: type 3 / X-sensitive), draw order restricted
template <int D, int ODD> DQ_HD u32 stabs_grid_to_type_index(u64 s) {
    u32 c = 0;
#pragma unroll
    for (int k = 0; k < Lat<D>::NT; ++k) c |= (u32)((s >> Lat<D>::type_pos(ODD, k)) & 1ull) << k;
    return c;
}

// ---- legal moves (closed form of Environments.py:238-271 + :187-196) ---------------------
// qubits touching a plaquette of `summed`
template <int D> DQ_HD u64 qubits_adjacent_to(u64 summed) {
    constexpr int G = D + 1;
    return (summed | (summed >> 1) | (summed >> G) | (summed >> (G + 1))) & Lat<D>::QMASK;
}
// union of the in-lattice 8-neighbourhoods (self excluded) of the qubits in `q`
template <int D> DQ_HD u64 qubits_neighbours_of(u64 q) {
    constexpr int G = D + 1;
    return ((q << 1) | (q >> 1) | (q << G) | (q >> G) | (q << (G + 1)) | (q >> (G + 1)) |
            (q << (G - 1)) | (q >> (G - 1))) & Lat<D>::QMASK;
}

// ---- observation layer bitmaps (bit x*H+y of a P-bit little-endian multiword) -------------
DQ_HD u32 spread2_8(u32 x) {        // bit i (i<8) -> bit 2i
    x = (x | (x << 4)) & 0x0F0Fu;
    x = (x | (x << 2)) & 0x3333u;
    x = (x | (x << 1)) & 0x5555u;
    return x;
}
template <int NW> DQ_HD void or_bits_const(u64 (&w)[NW], int off, u64 v) {   // off is a compile-time constant after unrolling
    int i = off >> 6, s = off & 63;
    w[i] |= v << s;
    if (s != 0 && i + 1 < NW) w[i + 1] |= v >> (64 - s);
}
// syndrome layer: markers | f[a][b] at cell (2a,2b)
template <int D> DQ_HD void syndrome_layer_bitmap(u64 f, u64 (&w)[Lat<D>::PW]) {
    constexpr int G = D + 1, H = Lat<D>::H;
#pragma unroll
    for (int i = 0; i < Lat<D>::PW; ++i) w[i] = Lat<D>::marker_word(i);
#pragma unroll
    for (int a = 0; a <= D; ++a) {
        u32 row = (u32)(f >> (a * G)) & ((1u << G) - 1);
        or_bits_const<Lat<D>::PW>(w, 2 * a * H, (u64)spread2_8(row));
    }
}
// action layer: completed action on qubit (r,c) at cell (2r+1,2c+1)
template <int D> DQ_HD void action_layer_bitmap(u64 act, u64 (&w)[Lat<D>::PW]) {
    constexpr int G = D + 1, H = Lat<D>::H;
#pragma unroll
    for (int i = 0; i < Lat<D>::PW; ++i) w[i] = 0;
#pragma unroll
    for (int r = 0; r < D; ++r) {
        u32 row = (u32)(act >> (r * G)) & ((1u << D) - 1);
        or_bits_const<Lat<D>::PW>(w, (2 * r + 1) * H + 1, (u64)spread2_8(row));
    }
}

// ---- Philox4x32-10 ---------------------------------------------------------------------
DQ_HD u32 mulhi32(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}
struct Philox4 { u32 x, y, z, w; };
DQ_HD Philox4 philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        u32 h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        u32 h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        u32 n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// n-bit field (n <= 64) at bit offset `off` of a little-endian u32 stream; the stream must be
// readable two words past the field.
DQ_HD u64 extract_bits(const u32* s, int off, int n) {
    int w = off >> 5, sh = off & 31;
    u64 lo = ((u64)s[w + 1] << 32) | s[w];
    u64 v = lo >> sh;
    if (sh) v |= (u64)s[w + 2] << (64 - sh);
    return n >= 64 ? v : (v & ((1ull << n) - 1));
}

// position of the k-th (0-based) set bit, k < popcount(x): binary search over popcounts, branch-free (clearing the
// lowest set bit k times made a warp's dependent chain as long as its largest k)
DQ_HD int popc32(u32 x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
DQ_HD int select64(u64 x, int k) {
    u32 w = (u32)x;
    int base = 0;
    const int c0 = popc32(w);
    if (k >= c0) { k -= c0; w = (u32)(x >> 32); base = 32; }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const int c = popc32(w & ((1u << s) - 1u));
        if (k >= c) { k -= c; w >>= s; base += s; }
    }
    return base;
}

}  // namespace dq
