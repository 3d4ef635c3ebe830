mkdir -p gpurun_out
timeout 600 python tools/train_demo.py --model X --p 0.007 --steps 4e7 --eps-steps 1e7 --out gpurun_out/train_x_p007.json 2>&1 | tail -3
timeout 1200 python tools/train_demo.py --model DP --curriculum "0.001:2e7,0.003:2e7,0.005:3e7,0.007:6e7" --eps-steps 6e6 --out gpurun_out/train_dp_curriculum.json 2>&1 | grep -v "^step" | tail -6
