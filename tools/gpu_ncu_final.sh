#!/bin/bash
# ncu evidence for the final build: launch list of the bench command, --set full capture of one 64-step rollout launch, and the bf16 update's kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_rollout.csv python bench.py --steps 1024 --warmup 16 --cpu-seconds 0.2 --no-dqn > gpurun_out/r2f_ncu1.log 2>&1; echo "launch list rc=$?"
DQ_ONLY_ROLLOUT=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 4 -c 1 -f -o gpurun_out/r2f_rollout64 python tools/prof_rollout.py > gpurun_out/r2f_ncu_ro.log 2>&1; echo "full capture rc=$?"
ncu -i gpurun_out/r2f_rollout64.ncu-rep --page raw --csv > gpurun_out/r2f_rollout64_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_dw_tma -s 10 -c 1 -f -o gpurun_out/r2f_tc_dw python tools/prof_train.py 4096 bf16 bf16 > gpurun_out/r2f_ncu_dw.log 2>&1; echo "dw capture rc=$?"
ncu -i gpurun_out/r2f_tc_dw.ncu-rep --page raw --csv > gpurun_out/r2f_tc_dw_raw.csv 2>/dev/null
python - <<'PY'
import csv
for name in ("r2f_rollout64_raw", "r2f_tc_dw_raw"):
    rows = list(csv.reader(open("gpurun_out/%s.csv" % name)))
    hdr, vals = rows[0], rows[-1]
    want = ("Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "dram__bytes_write.sum", "dram__bytes_read.sum", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active.avg.pct", "smsp__issue_active.avg.pct", "sm__throughput.avg.pct", "dram__throughput.avg.pct")
    print(name)
    for h, v in zip(hdr, vals):
        if any(w in h for w in want): print("  ", h, v[:90])
PY
