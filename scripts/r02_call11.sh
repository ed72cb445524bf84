set -x
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo rc=$?; tail -5 gpurun_out/r02_bench_2gpu.err; tail -c 1400 gpurun_out/r02_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload llama7b > gpurun_out/r02_bench_llama7b_2gpu.json 2> gpurun_out/r02_bench_llama7b_2gpu.err; echo rc=$?; tail -3 gpurun_out/r02_bench_llama7b_2gpu.err; cat gpurun_out/r02_bench_llama7b_2gpu.json
timeout 600 python -m pytest tests -x -q -m gpu -k "two_gpu or 2gpu or multi_device or second" 2>&1 | tail -5
