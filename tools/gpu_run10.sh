mkdir -p gpurun_out
timeout 600 python -m cProfile -o gpurun_out/fit.prof tools/train_demo.py --model X --p 0.007 --steps 6e6 --eps-steps 2e6 --test-episodes 64 --out gpurun_out/train_prof.json > /dev/null 2>&1
python - <<'PY'
import pstats
p = pstats.Stats("gpurun_out/fit.prof"); p.sort_stats("tottime").print_stats(28)
PY
