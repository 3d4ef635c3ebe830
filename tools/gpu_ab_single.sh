#!/bin/bash
# same-box A/B of the env kernel's per-launch cost: default build vs control builds (DQ_DECODING_LIB), tools/prof_rollout.py
TAG=${1:-absingle}
VARIANTS=${2:-"new old"}
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_env_gpu.py -x -q > gpurun_out/${TAG}_pytest_env.out 2>&1; echo "pytest env rc=$?"; tail -1 gpurun_out/${TAG}_pytest_env.out
for rep in 1 2; do
for v in $VARIANTS; do
  unset DQ_DECODING_LIB DQ_ENV_PDL
  case $v in
    new) ;;
    nopdl) export DQ_ENV_PDL=0 ;;
    *) export DQ_DECODING_LIB=build/variants/libdq_$v.so; [ -f "$DQ_DECODING_LIB" ] || continue ;;
  esac
  timeout 40 python tools/prof_rollout.py > gpurun_out/${TAG}_${v}_${rep}.out 2> gpurun_out/${TAG}_${v}_${rep}.err; echo "$v $rep: $(cut -c1-200 gpurun_out/${TAG}_${v}_${rep}.out)"
done
done
unset DQ_DECODING_LIB DQ_ENV_PDL
timeout 120 python bench.py --cpu-seconds 1 --no-dqn --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.out 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench.out"):
    if l.startswith("{"):
        d = json.loads(l); print("value %.4g single %s e2e %.4g" % (d["value"], json.dumps(d["single_step_launches"]), d["e2e"]["value"]))
PY
