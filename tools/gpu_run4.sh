mkdir -p gpurun_out
echo skip-tests

timeout 600 python bench.py --cpu-seconds 3 > gpurun_out/bench_dqn.json 2> gpurun_out/bench_dqn.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_dqn.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_dqn.json"))
print("value %.3e  e2e %.3e  kernel_us %.2f" % (j["value"], j["e2e"]["value"], j["roofline"]["kernel_us_mean"]))
print(json.dumps(j["dqn"], indent=1))
PY
