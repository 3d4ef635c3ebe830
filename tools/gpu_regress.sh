#!/bin/bash
# Regression on one B200 (run under gpurun): the GPU test suite, smoke, the bench with the driver's command line and with its defaults.
TAG=${1:-reg}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_steps20.json 2> gpurun_out/${TAG}_bench_steps20.err; echo "bench steps20 rc=$?"
timeout 600 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench default rc=$?"
python - <<PY
import json
for name in ("steps20", "default"):
    for l in open("gpurun_out/${TAG}_bench_%s.json" % name):
        if l.startswith("{"):
            d = json.loads(l)
            print(name, "value %.3g e2e %.3g packed %.3g frac %.3f single %.2f us" % (d["value"], d["e2e"]["value"], d["e2e_packed"]["value"], d["roofline"]["frac"], d["single_step_launches"]["ms_per_step"] * 1e3))
            q = d.get("dqn") or {}
            print(" act %.3g train %.3g upd %s" % (q.get("act_env_steps_per_s", 0), q.get("train_env_steps_per_s", 0), json.dumps(q.get("update_ms"))))
            print(" fwd bf16 ms", q.get("qnet_forward_bf16_ms"), "dqn_dp", json.dumps({k: v for k, v in (d.get("dqn_dp") or {}).items() if k in ("fused_us", "identical", "fit_finite", "error")}))
PY
