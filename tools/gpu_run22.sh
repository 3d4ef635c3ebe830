mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_env.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_env.log
tail -15 gpurun_out/pytest_env.log
timeout 200 python tools/prof_rollout.py > gpurun_out/rollout_c3.json 2> gpurun_out/rollout.err; cat gpurun_out/rollout_c3.json; tail -3 gpurun_out/rollout.err
DQ_N=8192 DQ_D=7 timeout 200 python tools/prof_rollout.py > gpurun_out/rollout_c5.json 2>> gpurun_out/rollout.err; cat gpurun_out/rollout_c5.json
