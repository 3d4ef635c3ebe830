#!/bin/bash
# host-buffer calls: zero-copy small outputs / actions vs the copy form (DQ_HOST_ZEROCOPY=0), parity first
TAG=${1:-zc}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_env_gpu.py tests/test_zz_packed_host_gpu.py -x -q > gpurun_out/${TAG}_pytest.out 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.out
DQ_HOST_ZEROCOPY=0 timeout 200 python -m pytest tests/test_env_gpu.py -x -q -k host > gpurun_out/${TAG}_pytest_copy.out 2>&1; echo "pytest (copy form) rc=$?"; tail -1 gpurun_out/${TAG}_pytest_copy.out
for rep in 1 2; do
for Z in 1 0; do
  DQ_HOST_ZEROCOPY=$Z timeout 200 python bench.py --cpu-seconds 1 --no-dqn --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_zc${Z}_${rep}.out 2> gpurun_out/${TAG}_bench_zc${Z}_${rep}.err
  python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench_zc${Z}_${rep}.out"):
    if l.startswith("{"):
        d = json.loads(l); print("zero_copy=${Z} rep ${rep}: e2e %.4g packed %.4g two_handles %.4g" % (d["e2e"]["value"], d["e2e_packed"]["value"], d["e2e"]["two_handles"]["value"]))
PY
done
done
timeout 100 python tools/prof_host_path.py > gpurun_out/${TAG}_host_path.json 2>&1; head -c 900 gpurun_out/${TAG}_host_path.json
