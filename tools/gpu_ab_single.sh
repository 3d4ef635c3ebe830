#!/bin/bash
# same-box A/B of the env kernel's per-launch cost: default build vs control builds (DQ_DECODING_LIB), tools/prof_rollout.py
TAG=${1:-absingle}
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_env_gpu.py -x -q > gpurun_out/${TAG}_pytest_env.out 2>&1; echo "pytest env rc=$?"; tail -1 gpurun_out/${TAG}_pytest_env.out
for rep in 1 2; do
for v in new old qglobal; do
  if [ $v = new ]; then unset DQ_DECODING_LIB; else export DQ_DECODING_LIB=build/variants/libdq_$v.so; fi
  [ $v != new ] && [ ! -f "$DQ_DECODING_LIB" ] && continue
  timeout 40 python tools/prof_rollout.py > gpurun_out/${TAG}_${v}_${rep}.out 2> gpurun_out/${TAG}_${v}_${rep}.err; echo "$v $rep: $(cut -c1-400 gpurun_out/${TAG}_${v}_${rep}.out)"
done
done
unset DQ_DECODING_LIB
