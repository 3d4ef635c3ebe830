#!/bin/bash
# One short GPU call for a rebuilt env kernel: parity first, then a same-box A/B of the default build against an earlier
# revision's kernel (libdq_old.so) and the tuning variants (tools/build_variants.sh), the bench line, the ncu launch list and
# one full ncu capture, then the remaining GPU tests.  Usage (under gpurun): bash tools/gpu_shot.sh <tag> <seconds available>
TAG=${1:-shot}; BUDGET=${2:-220}; T0=$(date +%s)
left() { echo $(( BUDGET - ($(date +%s) - T0) )); }
run() {   # name, max seconds, command...: skipped when less than 10 s remain
    name=$1; cap=$2; shift 2
    l=$(left); [ $l -lt 10 ] && { echo "skip $name (no time)"; return; }
    [ $cap -gt $l ] && cap=$l
    timeout -s KILL $cap "$@" > gpurun_out/${TAG}_$name.out 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$? t=$(( $(date +%s) - T0 ))s"
}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
run pytest_env 150 python -m pytest tests/test_env_gpu.py -x -q
tail -3 gpurun_out/${TAG}_pytest_env.out
grep -q " passed" gpurun_out/${TAG}_pytest_env.out && ! grep -q "failed" gpurun_out/${TAG}_pytest_env.out || { echo "env parity failed: stopping"; exit 1; }
run ab_new 60 python tools/prof_rollout.py
[ -f build/variants/libdq_old.so ] && DQ_DECODING_LIB=build/variants/libdq_old.so run ab_old 60 python tools/prof_rollout.py
[ -n "$SKIP_AB" ] || for v in $(ls build/variants/ | sed -n 's/^libdq_\(.*\)\.so$/\1/p' | grep -v '^old$'); do
    DQ_CALLS=12 DQ_ONLY_ROLLOUT=256 DQ_DECODING_LIB=build/variants/libdq_$v.so run ab_$v 30 python tools/prof_rollout.py
done
DQ_CALLS=6 DQ_ONLY_ROLLOUT=64 DQ_N=262144 run big 40 python tools/prof_rollout.py      # every SM fully loaded
DQ_D=7 DQ_N=8192 run d7 40 python tools/prof_rollout.py
for v in d7w3g4 d7w2g6; do DQ_CALLS=12 DQ_ONLY_ROLLOUT=256 DQ_D=7 DQ_N=8192 DQ_DECODING_LIB=build/variants/libdq_$v.so run d7_$v 30 python tools/prof_rollout.py; done
run hostprof 200 python tools/prof_host_path.py
run bench 150 python bench.py --cpu-seconds 3 --no-dqn
run bench20 100 python bench.py --cpu-seconds 1 --no-dqn --steps 20 --warmup 3
run bench_full 300 python bench.py --cpu-seconds 3
run ncu_list 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_rollout.csv python bench.py --steps 1024 --warmup 16 --cpu-seconds 0.2 --no-dqn
DQ_ONLY_ROLLOUT=64 run ncu_full 90 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_rollout64 python tools/prof_rollout.py
run ncu_full_single 90 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 30 -c 1 -f -o gpurun_out/${TAG}_single python tools/prof_rollout.py
run fold_head 90 python tools/check_fold_head.py
run pytest_rest 240 python -m pytest tests -m gpu -x -q --deselect tests/test_env_gpu.py
run smoke 60 python __graft_entry__.py smoke
for f in gpurun_out/${TAG}_ab_*.out gpurun_out/${TAG}_big.out gpurun_out/${TAG}_d7*.out; do echo "$f $(cut -c1-400 $f)"; done
