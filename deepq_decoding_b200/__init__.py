"""B200-native hot path of DeepQ-Decoding: the fault-tolerant surface-code decoding environment and the DQN inner loop.

    envs        VecSurfaceCodeEnv (N lattices per launch) and the reference-named N=1 class
    agents      keras-rl surface: DQNAgent.fit / test, policies, SequentialMemory, FileLogger, Adam
    qnet        QNetwork (flat Keras-layout parameters, fp32 SIMT and bf16 tcgen05 forward, backward)
    referee     referee decoders as lookup tables (shipped d=5 MLPs tabulated, minimum-weight tables for d=3/7)
    h5lite      Keras .h5f weight files, read and write
    evaluate    error-rate sweeps;  curriculum: iterative training over error rates;  parallel: one process per GPU

Everything computes through `libdq_decoding.so` (include/dq_decoding.h); there is no CPU path.
"""
__version__ = "0.1.0"
