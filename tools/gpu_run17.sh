mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qnet_gpu.py -m gpu -q --timeout 600 -k "reference_script" > gpurun_out/pytest_q7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q7.log
tail -30 gpurun_out/pytest_q7.log
