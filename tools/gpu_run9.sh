mkdir -p gpurun_out
timeout 900 python tools/train_demo.py --model DP --p 0.007 --steps 6e7 --eps-steps 1.5e7 --out gpurun_out/train_dp_p007.json 2>&1 | tail -8
timeout 600 python tools/train_demo.py --model X --p 0.007 --steps 4e7 --eps-steps 1e7 --out gpurun_out/train_x_p007.json 2>&1 | tail -4
