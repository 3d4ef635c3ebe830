#!/bin/bash
# GPU check of the bf16 training path: parity tests first (short timeouts: a lost mbarrier arrival traps after ~2 s per kernel), then the bench's dqn leg.
TAG=${1:-tctrain}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_qnet_gpu.py -x -q -k "bf16_training_path" > gpurun_out/${TAG}_pytest_tc.out 2>&1; echo "pytest tc rc=$?"; tail -15 gpurun_out/${TAG}_pytest_tc.out
timeout 300 python -m pytest tests/test_qnet_gpu.py -x -q -k "bf16_training_learns" > gpurun_out/${TAG}_pytest_learn.out 2>&1; echo "pytest learn rc=$?"; tail -5 gpurun_out/${TAG}_pytest_learn.out
timeout 300 python bench.py --cpu-seconds 2 > gpurun_out/${TAG}_bench.out 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/%s_bench.out" % "${TAG}"):
    if l.startswith("{"):
        d = json.loads(l)
        print(json.dumps(d.get("dqn"), indent=0)[:3000])
        print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
PY
