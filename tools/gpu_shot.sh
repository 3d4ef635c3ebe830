#!/bin/bash
# One short GPU call for a rebuilt env kernel: parity first, then a same-box A/B of the default build against the previous
# kernel and the tuning variants (tools/build_variants.sh), parity of the fastest build, the bench line and the ncu launch
# list on it, then whatever else fits.  Usage (under gpurun): bash tools/gpu_shot.sh <tag> <seconds available>
# (SKIP_AB=1 skips the first A/B loop: a second call can then spend its time on the later steps)
TAG=${1:-shot}; BUDGET=${2:-220}; T0=$(date +%s)
left() { echo $(( BUDGET - ($(date +%s) - T0) )); }
run() {   # name, max seconds, command...: skipped when less than 10 s remain
    name=$1; cap=$2; shift 2
    l=$(left); [ $l -lt 10 ] && { echo "skip $name (no time)"; return; }
    [ $cap -gt $l ] && cap=$l
    timeout $cap "$@" > gpurun_out/${TAG}_$name.out 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$? t=$(( $(date +%s) - T0 ))s"
}
mkdir -p gpurun_out
run pytest_env 90 python -m pytest tests/test_env_gpu.py -x -q
tail -2 gpurun_out/${TAG}_pytest_env.out
run ab_new 40 python tools/prof_rollout.py
DQ_DECODING_LIB=build/variants/libdq_old.so run ab_old 40 python tools/prof_rollout.py
[ -n "$SKIP_AB" ] || for v in so lk df dfso df2 df2so df2solk df2sopf1 df2sopf2 df2sot160 dfsopf1 bb2 bb2so bb2somi bb bbso bbsomi bbe32 pf1 mb8; do
    DQ_CALLS=12 DQ_ONLY_ROLLOUT=256 DQ_DECODING_LIB=build/variants/libdq_$v.so run ab_$v 25 python tools/prof_rollout.py
done
BEST=$(python - "$TAG" <<'PY'
import glob, json, sys
tag, best = sys.argv[1], ("new", 1e9)
for f in glob.glob("gpurun_out/%s_ab_*.out" % tag):
    name = f.split("_ab_")[1][:-4]
    try:
        us = json.loads(open(f).read().strip().splitlines()[-1])["rollout_256_us_per_step"]
    except Exception:
        continue
    if name != "old" and us < best[1]:
        best = (name, us)
print(best[0])
PY
)
echo "fastest build: $BEST"
if [ "$BEST" != "new" ]; then
    export DQ_DECODING_LIB=build/variants/libdq_$BEST.so
    run pytest_env_best 60 python -m pytest tests/test_env_gpu.py -x -q
    tail -2 gpurun_out/${TAG}_pytest_env_best.out
fi
DQ_CALLS=6 DQ_ONLY_ROLLOUT=64 DQ_N=262144 run big_best 40 python tools/prof_rollout.py      # the same build with every SM fully loaded
run bench_best 90 python bench.py --cpu-seconds 3 --no-dqn --no-experiments
run ncu_list 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_rollout.csv python bench.py --steps 1024 --warmup 16 --cpu-seconds 0.2 --no-dqn --no-experiments
DQ_ONLY_ROLLOUT=64 run ncu_full 60 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_rollout64 python tools/prof_rollout.py
unset DQ_DECODING_LIB
[ "$BEST" != "new" ] && run bench_default 90 python bench.py --cpu-seconds 3 --no-dqn --no-experiments
run pytest_rest 90 python -m pytest tests -m gpu -x -q --deselect tests/test_env_gpu.py
run smoke 40 python __graft_entry__.py smoke
DQ_HOST_EXPAND=1 run bench_host_expand 90 python bench.py --cpu-seconds 3 --no-dqn --no-experiments      # e2e with observations moved bit-packed + expanded on the host
DQ_HOST_EXPAND=1 run pytest_host_expand 60 python -m pytest tests/test_env_gpu.py -x -q -k "host_buffer"
run fold_head 90 python tools/check_fold_head.py      # opt-in folded head of the bf16 acting path: Q against unfolded / fp32, forward time
for v in e16t96 bb2t96 pf2 e8t64mb14 e32t256mb4; do
    DQ_ONLY_ROLLOUT=256 DQ_DECODING_LIB=build/variants/libdq_$v.so run ab_$v 25 python tools/prof_rollout.py
done
for f in gpurun_out/${TAG}_ab_*.out; do echo "$f $(cut -c1-300 $f)"; done
