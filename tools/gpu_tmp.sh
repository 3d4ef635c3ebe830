mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qnet_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -3
timeout 300 python tools/prof_qnet.py 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|head_dueling" --csv --log-file gpurun_out/qnet_pipe.csv python tools/prof_qnet.py > gpurun_out/ncu_q.log 2>&1; echo rc=$?
