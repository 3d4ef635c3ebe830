mkdir -p gpurun_out
timeout 1500 python tools/train_demo.py --model DP --curriculum "0.001:3e7,0.003:4e7,0.005:7e7,0.007:1.6e8" --eps-steps 8e6 --target bf16 --lr 1e-4 --updates 2 --out gpurun_out/train_dp_curriculum.json 2>&1 | grep -E "^tested|^\{|step (2|4|6|10|14)[0-9]{7} " | tail -12
