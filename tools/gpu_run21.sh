mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_ro.json 2> gpurun_out/bench_ro.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/bench_ro.json; tail -3 gpurun_out/bench_ro.err
DQ_ONLY_ROLLOUT=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 4 -c 1 -o gpurun_out/r1_rollout python tools/prof_rollout.py > gpurun_out/ncu_ro.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_ro.log
