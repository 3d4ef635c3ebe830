#!/bin/bash
TAG=${1:-tcprof2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_qnet_gpu.py -x -q > gpurun_out/${TAG}_pytest_qnet.out 2>&1; echo "pytest qnet rc=$?"; tail -3 gpurun_out/${TAG}_pytest_qnet.out
for MC in 4 8 16 32 64; do echo "DQ_TC_DW_MINCHUNKS=$MC"; DQ_TC_DW_MINCHUNKS=$MC timeout 60 python tools/prof_train.py 4096 bf16 bf16; done 2>&1 | tee gpurun_out/${TAG}_update_times.txt
timeout 60 python tools/prof_train.py 1024 bf16 bf16 | tee -a gpurun_out/${TAG}_update_times.txt
timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_train.py 4096 bf16 bf16 > gpurun_out/${TAG}_ncu.out 2>&1; echo "ncu rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = r[vi]
ids = list(per)
last = [i for i in ids if "adam_kernel" in per[i]["k"]]
lo = ids.index(last[-2]) + 1
tot = 0
for i in ids[lo:]:
    d = per[i]; tot += float(d.get("gpu__time_duration.sum", "0").replace(",", ""))
    print(i, d["k"][:50], d.get("gpu__time_duration.sum"), d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
print("sum of the last update's launches (cold, serialised): %.1f us over %d launches" % (tot / 1e3, len(ids) - lo))
PY
