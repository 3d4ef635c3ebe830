mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_qnet_gpu.py -m gpu -q --timeout 500 > gpurun_out/pytest_mg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mg.log
tail -25 gpurun_out/pytest_mg.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/bench_allreduce.py > gpurun_out/allreduce_n2.json 2> gpurun_out/allreduce_n2.err; echo "rc=$?"
cat gpurun_out/allreduce_n2.json; tail -5 gpurun_out/allreduce_n2.err
