mkdir -p gpurun_out
DQ_DECODING_LIB=$PWD/deepq_decoding_b200/libdq_e32_t256_b4.so timeout 300 python -m pytest tests/test_env_gpu.py -m gpu -q --timeout 200 -x -k "rollout or trajectory or traj" 2>&1 | tail -3
for v in e32_t256_b4 e32_t128_b7 e32_t128_b4; do
  lib=deepq_decoding_b200/libdq_$v.so
  echo "variant=$v"
  DQ_DECODING_LIB=$PWD/$lib DQ_ONLY_ROLLOUT=64 timeout 100 python tools/prof_rollout.py 2>&1 | tail -1
  DQ_DECODING_LIB=$PWD/$lib DQ_ONLY_ROLLOUT=64 DQ_N=8192 DQ_D=7 timeout 100 python tools/prof_rollout.py 2>&1 | tail -1
done
