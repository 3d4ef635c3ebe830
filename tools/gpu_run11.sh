mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_all2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all2.log
tail -25 gpurun_out/pytest_all2.log
timeout 600 python bench.py --cpu-seconds 2 --no-dqn > gpurun_out/bench_v5f.json 2> gpurun_out/bench_v5f.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_v5f.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_v5f.json"))
print("value %.3e ms/step %.4f kernel_us %.2f frac %.3f e2e %.3e" % (j["value"], j["ms_per_step"], j["roofline"]["kernel_us_mean"], j["roofline"]["frac"], j["e2e"]["value"]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 40 -c 3 -o gpurun_out/r1g_env_step python bench.py --steps 64 --warmup 16 --cpu-seconds 0.2 --no-dqn > gpurun_out/ncu2.log 2>&1
