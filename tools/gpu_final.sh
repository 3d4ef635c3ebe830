mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 32768 --warmup 64 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_final.json"))
print("value %.4e ms/step %.5f e2e %.3e clocks %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["clocks"]))
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in j["roofline"].items()})
print("cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"])
print("dqn", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in j["dqn"].items() if "steps" in k or "_ms" in k or "tflops" in k})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 64 --warmup 16 --cpu-seconds 0.2 --no-dqn > gpurun_out/ncu1.log 2>&1
timeout 600 python tools/train_demo.py --model X --p 0.007 --steps 4e7 --eps-steps 1e7 --out gpurun_out/train_x_p007.json 2>&1 | tail -2
