#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_qnet_gpu.py -x -q -k "fit or learns or training_run or reference_named" > gpurun_out/fitcheck_pytest.out 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/fitcheck_pytest.out
timeout 200 python tools/train_demo.py --model X --p 0.007 --envs 4096 --steps 4e7 --target bf16 --train bf16 --out gpurun_out/train_x_p007_bf16_v2.json 2>&1 | tail -1 | cut -c1-400
timeout 200 python tools/train_demo.py --model X --p 0.007 --envs 4096 --steps 4e7 --target bf16 --train fp32 --out gpurun_out/train_x_p007_fp32_v2.json 2>&1 | tail -1 | cut -c1-400
