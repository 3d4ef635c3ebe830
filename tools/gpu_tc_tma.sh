#!/bin/bash
# bf16 update: dW GEMMs fed by TMA vs by cp.async (DQ_TC_DW_TMA=0), parity first
TAG=${1:-tctma}
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_qnet_gpu.py -x -q -k "bf16_training" > gpurun_out/${TAG}_pytest_tc.out 2>&1; echo "pytest tc (TMA) rc=$?"; tail -8 gpurun_out/${TAG}_pytest_tc.out
for rep in 1 2; do
for T in 1 0; do echo "DQ_TC_DW_TMA=$T"; DQ_TC_DW_TMA=$T timeout 60 python tools/prof_train.py 4096 bf16 bf16; DQ_TC_DW_TMA=$T timeout 60 python tools/prof_train.py 1024 bf16 bf16; done
done 2>&1 | tee gpurun_out/${TAG}_update_times.txt
timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tc_dw -c 40 --csv --log-file gpurun_out/${TAG}_dw_tma.csv python tools/prof_train.py 4096 bf16 bf16 > /dev/null 2>&1
DQ_TC_DW_TMA=0 timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tc_dw -c 40 --csv --log-file gpurun_out/${TAG}_dw_cpasync.csv python tools/prof_train.py 4096 bf16 bf16 > /dev/null 2>&1
for f in dw_tma dw_cpasync; do echo $f; grep gpu__time_duration gpurun_out/${TAG}_$f.csv | tail -5 | awk -F'","' '{print $5, $NF}'; done
