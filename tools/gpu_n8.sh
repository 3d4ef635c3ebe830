#!/bin/bash
# N-GPU bench with the driver's command line (gpurun --gpus N): both arms
N=${1:-8}; TAG=${2:-n8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4g e2e %.4g packed %.4g n_gpus %d" % (d["value"], d["e2e"]["value"], d["e2e_packed"]["value"], d["n_gpus"]))
        print(json.dumps(d.get("dqn_dp"))[:900])
PY
