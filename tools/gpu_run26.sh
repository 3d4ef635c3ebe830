mkdir -p gpurun_out
DQ_ONLY_ROLLOUT=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 4 -c 1 -o gpurun_out/r1_rollout64b python tools/prof_rollout.py > gpurun_out/ncu_ro.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ro.log
