#!/bin/bash
# per-kernel launch list of one bf16 update (cold-cache, serialised ncu times) + warm update time at several batches
TAG=${1:-tcprof}
mkdir -p gpurun_out
for B in 1024 4096; do timeout 60 python tools/prof_train.py $B bf16 bf16; timeout 60 python tools/prof_train.py $B bf16 fp32; done 2>&1 | tee gpurun_out/${TAG}_update_times.txt
timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none --launch-skip 0 -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_train.py 4096 bf16 bf16 > gpurun_out/${TAG}_ncu.out 2>&1; echo "ncu rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = r[vi]
ids = list(per)
# the last update = the launches after the last replay... print the final 60 launches
for i in ids[-75:]:
    d = per[i]
    print(i, d["k"][:60], d.get("gpu__time_duration.sum"), d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
PY
