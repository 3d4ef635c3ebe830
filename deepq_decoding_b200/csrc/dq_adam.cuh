// Keras-2 Adam update of one parameter, shared by the stand-alone optimizer kernel (dq_qnet.cu) and the fused
// all-reduce + Adam kernel (dq_comm.cu).  Every operation is an explicitly rounded intrinsic so that the compiler's
// FMA contraction cannot differ between the two kernels: a sharded run with either collective produces the same bits.
//   m <- b1 m + (1-b1) g ;  v <- b2 v + (1-b2) g^2 ;  p <- p - lr_t m / (sqrt(v) + eps)      (keras/optimizers.py, Adam.get_updates)
#pragma once
#include <cuda_runtime.h>

namespace dq {
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float lr_t, float b1, float b2, float eps) {
    m = __fmaf_rn(b1, m, __fmul_rn(__fsub_rn(1.f, b1), g));
    v = __fmaf_rn(b2, v, __fmul_rn(__fmul_rn(__fsub_rn(1.f, b2), g), g));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(lr_t, m), __fadd_rn(__fsqrt_rn(v), eps)));
}
}  // namespace dq
