// empty on purpose: tests/host/cuda_emu.h (force-included) provides what the emulated sources use
