mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for W in 8 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2971$W tools/bench_allreduce.py > gpurun_out/allreduce_n$W.json 2> gpurun_out/allreduce_n$W.err; echo "rc=$?"
cat gpurun_out/allreduce_n$W.json
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench rc=$?"
cat gpurun_out/bench_n8.json | cut -c1-1500
