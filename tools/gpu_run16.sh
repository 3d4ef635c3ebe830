mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_env_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_e.log 2>&1; tail -3 gpurun_out/pytest_e.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_final.json"))
print("c3 value %.4e ms/step %.5f e2e %.3e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]))
for r in j["roofline_scaling"]: print({k: round(v, 3) for k, v in r.items()})
PY
timeout 600 python bench.py --workload c1 --steps 8192 --cpu-seconds 2 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1 rc=$?"; tail -2 gpurun_out/bench_c1.err; cut -c1-200 gpurun_out/bench_c1.json
