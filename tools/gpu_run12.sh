mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/train_multi_gpu.py 2> gpurun_out/mgpu.err | tail -2 | tee gpurun_out/train_2gpu.json
tail -3 gpurun_out/mgpu.err
