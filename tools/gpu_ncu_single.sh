#!/bin/bash
# ncu --set full of ONE single-step launch (dq_env_rollout_random with 1 step) and of one 4-step launch: where a launch's fixed cost goes
mkdir -p gpurun_out
DQ_ONLY_ROLLOUT=1 DQ_CALLS=6 timeout 300 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 6 -c 1 -f -o gpurun_out/r2_single_step python tools/prof_rollout.py > gpurun_out/ncu_single.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2_single_step.ncu-rep --page raw --csv > gpurun_out/r2_single_step_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_single_step.ncu-rep --page source --csv > gpurun_out/r2_single_step_source.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_single_step_raw.csv")))
hdr, vals = rows[0], rows[-1]
want = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_write.sum", "dram__bytes_read.sum", "launch__occupancy_limit")
for h, v in zip(hdr, vals):
    if any(w in h for w in want): print(h, v)
PY
