#!/bin/bash
# from-scratch training, same box: fp32 updates vs bf16 tensor-core updates (everything else equal), then the depolarising curriculum with bf16 updates
mkdir -p gpurun_out
timeout 200 python tools/train_demo.py --model X --p 0.007 --envs 4096 --steps 4e7 --target bf16 --train fp32 --out gpurun_out/train_x_p007_fp32.json 2>&1 | tail -2
timeout 200 python tools/train_demo.py --model X --p 0.007 --envs 4096 --steps 4e7 --target bf16 --train bf16 --out gpurun_out/train_x_p007_bf16.json 2>&1 | tail -2
timeout 400 python tools/train_demo.py --model DP --p 0.007 --envs 4096 --eps-steps 8e6 --target bf16 --train bf16 --curriculum 0.001:3e7,0.003:4e7,0.005:7e7,0.007:1.6e8 --out gpurun_out/train_dp_curriculum_bf16.json 2>&1 | tail -2
python - <<PY
import json
for f in ("train_x_p007_fp32", "train_x_p007_bf16", "train_dp_curriculum_bf16"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "train_seconds %.1f env-steps/s %.3g lifetime %.1f +- %.1f" % (d["train_seconds"], d["train_env_steps_per_s"], d["test_mean_lifetime"], d["test_se"]))
    except Exception as ex:
        print(f, "missing:", ex)
PY
