// empty on purpose (see cuda_bf16.h)
