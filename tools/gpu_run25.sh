mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 500 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 2 --steps 2000 --warmup 10 --no-dqn --cpu-seconds 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 2>/dev/null | cut -c1-400
