#!/bin/bash
# 2-GPU regression (gpurun --gpus 2): multi-GPU tests, both bench arms at N = 2 with the driver's command line
TAG=${1:-n2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -q --timeout 500 > gpurun_out/${TAG}_pytest_multi.txt 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-160 gpurun_out/${TAG}_bench_ref.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4g e2e %.4g n_gpus %d" % (d["value"], d["e2e"]["value"], d["n_gpus"]))
        print(json.dumps(d.get("dqn_dp"))[:1200])
PY
