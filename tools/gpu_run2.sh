mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
for v in default; do
  if [ $v = default ]; then unset DQ_DECODING_LIB; else export DQ_DECODING_LIB=$PWD/deepq_decoding_b200/libdq_variant_$v.so; fi
  timeout 300 python bench.py --steps 8192 --warmup 64 --cpu-seconds 0.5 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_$v.json"))
    print("$v", "value=%.3e" % j["value"], "ms/step=%.4f" % j["ms_per_step"], "kernel_us=%.2f" % j["roofline"]["kernel_us_mean"], "frac=%.3f" % j["roofline"]["frac"], "e2e=%.3e" % j["e2e"]["value"])
except Exception as e:
    print("$v failed", e)
PY
done
unset DQ_DECODING_LIB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1e_launches.csv python bench.py --steps 64 --warmup 16 --cpu-seconds 0.2 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 40 -c 3 -o gpurun_out/r1e_env_step python bench.py --steps 64 --warmup 16 --cpu-seconds 0.2 > gpurun_out/ncu2.log 2>&1
ls gpurun_out
