#!/bin/bash
# Q-network chains with programmatic dependent launch: parity first, then update / forward timings with and without (DQ_QNET_PDL=0)
TAG=${1:-pdl}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_qnet_gpu.py -x -q > gpurun_out/${TAG}_pytest_qnet.out 2>&1; echo "pytest qnet rc=$?"; tail -3 gpurun_out/${TAG}_pytest_qnet.out
for rep in 1 2; do
for P in 1 0; do
  echo "DQ_QNET_PDL=$P"
  DQ_QNET_PDL=$P timeout 60 python tools/prof_train.py 4096 bf16 bf16
  DQ_QNET_PDL=$P timeout 60 python tools/prof_train.py 1024 bf16 bf16
  DQ_QNET_PDL=$P timeout 60 python tools/prof_train.py 4096 fp32 fp32
done
done 2>&1 | tee gpurun_out/${TAG}_update_times.txt
for P in 1 0; do DQ_QNET_PDL=$P timeout 200 python bench.py --cpu-seconds 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_pdl$P.out 2> gpurun_out/${TAG}_bench_pdl$P.err; echo "bench rc=$?"; done
python - <<PY
import json
for P in (1, 0):
    for l in open("gpurun_out/${TAG}_bench_pdl%d.out" % P):
        if l.startswith("{"):
            q = json.loads(l)["dqn"]
            print("PDL", P, "act %.4g/s (%.1f us) fwd bf16 %.1f us train_bf16 %.4g/s upd %s" % (q["act_env_steps_per_s"], q["act_ms_per_iteration"] * 1e3, q["qnet_forward_bf16_ms"] * 1e3, q["update_ms"].get("train_bf16_env_steps_per_s", 0), json.dumps({k: round(v, 3) for k, v in q["update_ms"].items() if k.endswith("_ms")})))
PY
