mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_qnet_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q.log
tail -40 gpurun_out/pytest_q.log
