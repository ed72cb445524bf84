set -x
# 2 GPUs: the default bench line (weak scaling svd + extras incl. the sharded llama7b run) and the checksum against N = 1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo rc=$?; tail -3 gpurun_out/r02_bench_2gpu.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload llama7b --no-cpu-baseline > gpurun_out/r02_bench_llama7b_1gpu.json 2> gpurun_out/r02_bench_llama7b_1gpu.err; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload llama7b > gpurun_out/r02_bench_llama7b_2gpu.json 2> gpurun_out/r02_bench_llama7b_2gpu.err; echo rc=$?
python - <<'PY'
import json
a=json.load(open("gpurun_out/r02_bench_llama7b_1gpu.json")); b=json.load(open("gpurun_out/r02_bench_llama7b_2gpu.json")); c=json.load(open("gpurun_out/r02_bench_2gpu.json"))
for d in (a,b): print(d["n_gpus"], d["config"]["seconds"], d["config"]["decompose_s"], d["config"]["exchange_s"], d["config"]["factors_checksum"])
print("bench 2gpu", c["value"], c["ms_per_step"], c["e2e"]["value"], c["extras"].get("llama7b_config3",{}).get("seconds"), c["extras"].get("llama7b_config3",{}).get("factors_checksum"), c["extras"].get("error"))
PY
# the OPT-125m CLI pipeline (config 1 shape) on 1 and 2 GPUs: calibration, sensitivity sweep (unit-sharded, batched ppl), search, final pass
mkdir -p gpurun_out/cli && cd gpurun_out/cli && rm -rf cache output
( time CUDA_VISIBLE_DEVICES=0 timeout 400 python ../../asvd.py --synthetic_model opt-125m --calib_dataset synthetic --act_aware --alpha 0.5 --n_calib_samples 16 --scaling_method abs_mean --param_ratio_target 0.9 ) > ../r02_cli_opt125m_1gpu.log 2>&1
rm -rf cache output
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 ../../asvd.py --synthetic_model opt-125m --calib_dataset synthetic --act_aware --alpha 0.5 --n_calib_samples 16 --scaling_method abs_mean --param_ratio_target 0.9 ) > ../r02_cli_opt125m_2gpu.log 2>&1
cd ../.. && rm -rf gpurun_out/cli
tail -n 8 gpurun_out/r02_cli_opt125m_1gpu.log; tail -n 8 gpurun_out/r02_cli_opt125m_2gpu.log
# the CPU arm of config 3 (bounded sample, extrapolated)
timeout 600 python bench.py --impl reference --workload llama7b > gpurun_out/r02_bench_llama7b_reference_arm.json 2>/dev/null; cat gpurun_out/r02_bench_llama7b_reference_arm.json | cut -c1-900
