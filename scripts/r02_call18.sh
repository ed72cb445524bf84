set -x
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo rc=$?; tail -3 gpurun_out/r02_bench_8gpu.err
python - <<'PY'
import json
c=json.load(open("gpurun_out/r02_bench_8gpu.json"))
print("bench 8gpu", c["value"], c["ms_per_step"], c["e2e"]["value"], c["config"]["step_ms"], c["config"]["hiccup"])
print(c["extras"].get("llama7b_config3"), c["extras"].get("error"))
PY
