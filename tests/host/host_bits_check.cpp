// Host build (g++) of the device bit-board helpers in deepq_decoding_b200/csrc/dq_lattice.cuh so the
// CPU test-suite can check each of them against the oracle without a GPU.  Test infrastructure only.
#include "../../deepq_decoding_b200/csrc/dq_lattice.cuh"
using namespace dq;

#define DISPATCH(expr3, expr5, expr7) (d == 3 ? (expr3) : d == 5 ? (expr5) : (expr7))

extern "C" {
u64 hb_true_syndrome(int d, u64 xb, u64 zb) { return DISPATCH(true_syndrome<3>(xb, zb), true_syndrome<5>(xb, zb), true_syndrome<7>(xb, zb)); }
int hb_label(int d, u64 xb, u64 zb) { return DISPATCH(homology_label<3>(xb, zb), homology_label<5>(xb, zb), homology_label<7>(xb, zb)); }
u64 hb_q_c2g(int d, u64 c) { return DISPATCH(qubits_compact_to_grid<3>(c), qubits_compact_to_grid<5>(c), qubits_compact_to_grid<7>(c)); }
u64 hb_q_g2c(int d, u64 g) { return DISPATCH(qubits_grid_to_compact<3>(g), qubits_grid_to_compact<5>(g), qubits_grid_to_compact<7>(g)); }
u64 hb_s_c2g(int d, u64 c) { return DISPATCH(stabs_compact_to_grid<3>(c), stabs_compact_to_grid<5>(c), stabs_compact_to_grid<7>(c)); }
u64 hb_s_g2c(int d, u64 g) { return DISPATCH(stabs_grid_to_compact<3>(g), stabs_grid_to_compact<5>(g), stabs_grid_to_compact<7>(g)); }
u32 hb_joint_index(int d, u64 s) { return d == 3 ? stabs_grid_to_joint_index<3>(s) : stabs_grid_to_joint_index<5>(s); }
u32 hb_type_index(int d, int odd, u64 s) {
    if (odd) return DISPATCH((stabs_grid_to_type_index<3, 1>(s)), (stabs_grid_to_type_index<5, 1>(s)), (stabs_grid_to_type_index<7, 1>(s)));
    return DISPATCH((stabs_grid_to_type_index<3, 0>(s)), (stabs_grid_to_type_index<5, 0>(s)), (stabs_grid_to_type_index<7, 0>(s)));
}
u64 hb_adjacent(int d, u64 s) { return DISPATCH(qubits_adjacent_to<3>(s), qubits_adjacent_to<5>(s), qubits_adjacent_to<7>(s)); }
u64 hb_neighbours(int d, u64 q) { return DISPATCH(qubits_neighbours_of<3>(q), qubits_neighbours_of<5>(q), qubits_neighbours_of<7>(q)); }
void hb_syn_layer(int d, u64 f, u64* out) {
    if (d == 3) { u64 w[Lat<3>::PW]; syndrome_layer_bitmap<3>(f, w); for (int i = 0; i < Lat<3>::PW; ++i) out[i] = w[i]; }
    if (d == 5) { u64 w[Lat<5>::PW]; syndrome_layer_bitmap<5>(f, w); for (int i = 0; i < Lat<5>::PW; ++i) out[i] = w[i]; }
    if (d == 7) { u64 w[Lat<7>::PW]; syndrome_layer_bitmap<7>(f, w); for (int i = 0; i < Lat<7>::PW; ++i) out[i] = w[i]; }
}
void hb_act_layer(int d, u64 a, u64* out) {
    if (d == 3) { u64 w[Lat<3>::PW]; action_layer_bitmap<3>(a, w); for (int i = 0; i < Lat<3>::PW; ++i) out[i] = w[i]; }
    if (d == 5) { u64 w[Lat<5>::PW]; action_layer_bitmap<5>(a, w); for (int i = 0; i < Lat<5>::PW; ++i) out[i] = w[i]; }
    if (d == 7) { u64 w[Lat<7>::PW]; action_layer_bitmap<7>(a, w); for (int i = 0; i < Lat<7>::PW; ++i) out[i] = w[i]; }
}
void hb_philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32* out) {
    Philox4 p = philox4x32_10(c0, c1, c2, c3, k0, k1); out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = p.w;
}
u64 hb_extract_bits(const u32* s, int off, int n) { return extract_bits(s, off, n); }
int hb_select64(u64 x, int k) { return select64(x, k); }
u64 hb_masks(int d, int which) {
    switch (which) {
        case 0: return DISPATCH(Lat<3>::QMASK, Lat<5>::QMASK, Lat<7>::QMASK);
        case 1: return DISPATCH(Lat<3>::T1, Lat<5>::T1, Lat<7>::T1);
        case 2: return DISPATCH(Lat<3>::T3, Lat<5>::T3, Lat<7>::T3);
    }
    return 0;
}
}
