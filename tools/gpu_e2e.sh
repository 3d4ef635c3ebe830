#!/bin/bash
TAG=${1:-e2e}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_env_gpu.py tests/test_zz_packed_host_gpu.py -x -q > gpurun_out/${TAG}_pytest.out 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.out
for rep in 1 2 3; do
  timeout 200 python bench.py --cpu-seconds 1 --no-dqn --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_${rep}.out 2> gpurun_out/${TAG}_bench_${rep}.err
  python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench_${rep}.out"):
    if l.startswith("{"):
        d = json.loads(l); print("rep ${rep}: e2e %.4g packed %.4g two_handles %.4g value %.4g" % (d["e2e"]["value"], d["e2e_packed"]["value"], d["e2e"]["two_handles"]["value"], d["value"]))
PY
done
