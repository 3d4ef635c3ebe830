#!/bin/bash
# One GPU call: runs the named steps (arguments after the tag) with a per-step cap, each into gpurun_out/<tag>_<step>.{out,err}.
# Usage (under gpurun): bash tools/gpu_shot2.sh <tag> <step> [<step> ...]     steps: see the case statement
TAG=${1:-shot}; shift
T0=$(date +%s)
run() { name=$1; cap=$2; shift 2; timeout -s KILL $cap "$@" > gpurun_out/${TAG}_$name.out 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$? t=$(( $(date +%s) - T0 ))s"; }
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
nproc > gpurun_out/${TAG}_nproc.txt
for step in "$@"; do
case $step in
  pytest_env)   run pytest_env 200 python -m pytest tests/test_env_gpu.py tests/test_zz_packed_host_gpu.py -x -q; tail -2 gpurun_out/${TAG}_pytest_env.out ;;
  pytest_all)   run pytest_all 420 python -m pytest tests -m gpu -x -q; tail -2 gpurun_out/${TAG}_pytest_all.out ;;
  pytest_rest)  run pytest_rest 300 python -m pytest tests -m gpu -x -q --deselect tests/test_env_gpu.py; tail -2 gpurun_out/${TAG}_pytest_rest.out ;;
  tiles)        for n in 4736 9472 14208 16384 18944; do DQ_CALLS=8 DQ_ONLY_ROLLOUT=256 DQ_N=$n run tiles_$n 30 python tools/prof_rollout.py; cat gpurun_out/${TAG}_tiles_$n.out; done ;;
  ab)           run ab_new 60 python tools/prof_rollout.py; cut -c1-600 gpurun_out/${TAG}_ab_new.out
                for v in $(ls build/variants/ 2>/dev/null | sed -n 's/^libdq_\(.*\)\.so$/\1/p'); do
                    DQ_CALLS=12 DQ_ONLY_ROLLOUT=256 DQ_DECODING_LIB=build/variants/libdq_$v.so run ab_$v 30 python tools/prof_rollout.py; echo "$v $(cut -c1-300 gpurun_out/${TAG}_ab_$v.out)"; done ;;
  d7)           DQ_D=7 DQ_N=8192 run d7 40 python tools/prof_rollout.py; cut -c1-600 gpurun_out/${TAG}_d7.out ;;
  big)          DQ_CALLS=6 DQ_ONLY_ROLLOUT=64 DQ_N=262144 run big 40 python tools/prof_rollout.py; cat gpurun_out/${TAG}_big.out ;;
  hostprof)     run hostprof 200 python tools/prof_host_path.py; cut -c1-3000 gpurun_out/${TAG}_hostprof.out ;;
  hostexp)      run hostexp 120 python tools/prof_host_expand.py; cat gpurun_out/${TAG}_hostexp.out ;;
  bench20)      run bench20 300 python bench.py --gpus 1 --steps 20 --warmup 5; cut -c1-2500 gpurun_out/${TAG}_bench20.out ;;
  bench)        run bench 300 python bench.py --cpu-seconds 3; cut -c1-2500 gpurun_out/${TAG}_bench.out ;;
  bench_nodqn)  run bench_nodqn 200 python bench.py --cpu-seconds 3 --no-dqn --steps 20 --warmup 5; cut -c1-2500 gpurun_out/${TAG}_bench_nodqn.out ;;
  benchref)     run benchref 120 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5; cut -c1-800 gpurun_out/${TAG}_benchref.out ;;
  ncu_list)     run ncu_list 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_rollout.csv python bench.py --steps 1024 --warmup 16 --cpu-seconds 0.2 --no-dqn ;;
  ncu_full)     DQ_ONLY_ROLLOUT=64 run ncu_full 120 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_rollout64 python tools/prof_rollout.py ;;
  ncu_single)   run ncu_single 120 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 30 -c 1 -f -o gpurun_out/${TAG}_single python tools/prof_rollout.py ;;
  ncu_qnet)     run ncu_qnet 180 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_qnet_launches.csv python tools/prof_qnet.py ;;
  smoke)        run smoke 90 python __graft_entry__.py smoke; tail -2 gpurun_out/${TAG}_smoke.out ;;
  mgpu_tests)   run mgpu_tests 400 python -m pytest tests/test_multi_gpu.py -x -q; tail -3 gpurun_out/${TAG}_mgpu_tests.out ;;
  benchN)       N=$(nvidia-smi -L | wc -l); run benchN 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5; grep '^{' gpurun_out/${TAG}_benchN.out | cut -c1-3000 ;;
  benchrefN)    N=$(nvidia-smi -L | wc -l); run benchrefN 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 20 --warmup 5; grep '^{' gpurun_out/${TAG}_benchrefN.out | cut -c1-600 ;;
  *)            echo "unknown step $step" ;;
esac
done
