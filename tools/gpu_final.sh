# Round-end regression on one B200 (run under gpurun): GPU tests, smoke, both bench arms, ncu launch list + full capture.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1_launches_rollout.csv python bench.py --steps 1024 --warmup 16 --cpu-seconds 0.2 --no-dqn > gpurun_out/ncu1.log 2>&1
DQ_ONLY_ROLLOUT=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 4 -c 1 -o gpurun_out/r1_rollout64 python tools/prof_rollout.py > gpurun_out/ncu_ro.log 2>&1
