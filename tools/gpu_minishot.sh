#!/bin/bash
# The shortest useful GPU call: env parity tests, rollout timing of the default build and of one control build, a short bench line.
TAG=${1:-mini}
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_env_gpu.py -x -q > gpurun_out/${TAG}_pytest_env.out 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/${TAG}_pytest_env.out
timeout 20 python tools/prof_rollout.py > gpurun_out/${TAG}_ab_new.out 2> gpurun_out/${TAG}_ab_new.err; cut -c1-600 gpurun_out/${TAG}_ab_new.out
[ -n "$2" ] && { DQ_ONLY_ROLLOUT=256 DQ_DECODING_LIB=$2 timeout 15 python tools/prof_rollout.py > gpurun_out/${TAG}_ab_ctl.out 2>&1; cat gpurun_out/${TAG}_ab_ctl.out; }
timeout 30 python bench.py --cpu-seconds 1 --no-dqn > gpurun_out/${TAG}_bench.out 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench.out
