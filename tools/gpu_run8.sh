mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_qnet_gpu.py -m gpu -q --timeout 600 -k "tensor_core or lifetime" > gpurun_out/pytest_q4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q4.log
tail -5 gpurun_out/pytest_q4.log
timeout 300 python tools/prof_qnet.py
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 8192 --warmup 64 --cpu-seconds 2 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -5 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_2gpu.json") if l.startswith("{")][-1])
print("n_gpus", j["n_gpus"], "value %.3e e2e %.3e ms/step %.4f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in j["dqn"].items() if "env_steps" in k or "_ms" in k})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 50 --warmup 3 | tail -1 | cut -c1-300
