#!/bin/bash
mkdir -p gpurun_out
for CH in 1 2 3 4; do
  DQ_HOST_CHUNKS=$CH timeout 200 python bench.py --cpu-seconds 0.5 --no-dqn --steps 20 --warmup 5 > gpurun_out/chunks_$CH.out 2> gpurun_out/chunks_$CH.err
  python - <<PY
import json
for l in open("gpurun_out/chunks_$CH.out"):
    if l.startswith("{"):
        d = json.loads(l); print("chunks $CH: e2e %.4g packed %.4g two_handles %.4g" % (d["e2e"]["value"], d["e2e_packed"]["value"], d["e2e"]["two_handles"]["value"]))
PY
done
