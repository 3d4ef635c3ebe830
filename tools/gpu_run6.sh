mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qnet_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_q2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q2.log
tail -15 gpurun_out/pytest_q2.log
timeout 600 python bench.py --cpu-seconds 2 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_tc.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_tc.json"))
print(json.dumps(j["dqn"], indent=1))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|conv1_bits|gemm_fwd|prep_wt|dueling|eps_greedy|env_step" -c 120 --csv --log-file gpurun_out/r1_qnet_launches.csv python bench.py --steps 16 --warmup 3 --cpu-seconds 0.2 > gpurun_out/ncu_q1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c 6 -o gpurun_out/r1_tc_gemm python bench.py --steps 16 --warmup 3 --cpu-seconds 0.2 > gpurun_out/ncu_q2.log 2>&1
ls gpurun_out | head -40
