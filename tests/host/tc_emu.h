// tc_emu.h -- TEST INFRASTRUCTURE: what the bf16 half of csrc/dq_qnet.cu needs on top of cuda_emu.h when it is compiled for the
// CPU by tests/emu_qnet.py (tc variant): a bfloat16 storage type with round-to-nearest-even conversion, and float4.
// The tcgen05 / TMEM kernels themselves cannot run on a CPU: emu_qnet.py swaps the regions the product source marks
// "[tcgen05 kernels: begin/end]" for the plain loops of tests/host/tc_ref_gemm.inc / tc_ref_dw.inc (same launch arguments,
// bf16 operands, fp32 accumulation), so that everything AROUND them -- the layer geometry and padding, the transposes, the
// col2im gather, masks, bias sums, the host-side sequence of dq_qnet_forward_tc_train / dq_qnet_backward_tc -- executes as written.
#pragma once
#include <stdint.h>
#include <string.h>
struct __nv_bfloat16 { uint16_t x; };
static inline __nv_bfloat16 __float2bfloat16(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    __nv_bfloat16 r;
    if ((u & 0x7fffffffu) > 0x7f800000u) { r.x = 0x7fff; return r; }
    u += 0x7fffu + ((u >> 16) & 1u);
    r.x = (uint16_t)(u >> 16);
    return r;
}
static inline float __bfloat162float(__nv_bfloat16 b) { uint32_t u = (uint32_t)b.x << 16; float f; memcpy(&f, &u, 4); return f; }
static inline unsigned short __bfloat16_as_ushort(__nv_bfloat16 b) { return b.x; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
