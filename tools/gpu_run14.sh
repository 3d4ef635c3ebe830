mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qnet_gpu.py -m gpu -q --timeout 600 -k "curriculum or update or fit" > gpurun_out/pytest_q6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q6.log
tail -12 gpurun_out/pytest_q6.log
python tools/prof_train.py 1024 fp32; python tools/prof_train.py 1024 bf16; python tools/prof_train.py 4096 bf16
timeout 600 python tools/train_demo.py --model X --p 0.007 --steps 4e7 --eps-steps 1e7 --target bf16 --out gpurun_out/train_x_p007_bf16t.json 2>&1 | tail -2
