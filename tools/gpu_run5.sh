mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qnet_gpu.py -m gpu -q --timeout 300 -k tensor_core > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log
