mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qnet_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_q5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q5.log
tail -6 gpurun_out/pytest_q5.log
python tools/prof_train.py 1024; python tools/prof_train.py 4096
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_train_launches.csv python tools/prof_train.py 1024 > gpurun_out/ncu_t.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r1_train_launches.csv')))
hi = next(i for i,r in enumerate(rows) if r and r[0]=='ID'); hdr=rows[hi]; data=rows[hi+1:]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in data[-200:]:
    if len(r)>vi: agg.setdefault((r[ki][:58], r[gi]), []).append(float(r[vi].replace(',','')))
for k,v in agg.items():
    if sum(v)/len(v) > 12000: print(f"{k[0]:58s} grid={k[1]:14s} n={len(v):3d} mean={sum(v)/len(v)/1000:8.1f} us")
PY
