#!/bin/bash
# the bf16 forward kernel by kernel at 16 384 observations: duration + tensor-pipe activity (ncu, cold-cache serialised), and --set full of the layer-1 kernel
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"tc_gemm|head_dueling" -c 40 --csv --log-file gpurun_out/r2f_qnet_fwd_launches.csv python tools/prof_qnet.py > gpurun_out/r2f_qnet_fwd.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2f_qnet_fwd_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID"); gi = hdr.index("Grid Size")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {"k": r[ki].split("(")[0], "grid": r[gi]})[r[mi]] = r[vi]
for i in list(per)[-5:]:
    d = per[i]; print(d["k"][:40], d["grid"], {k.split(".")[0]: v for k, v in d.items() if k not in ("k", "grid")})
PY
tail -2 gpurun_out/r2f_qnet_fwd.log
