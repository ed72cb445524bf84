cd /root/repo
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload llama7b > gpurun_out/r02_bench_llama7b_${N}gpu.json 2> gpurun_out/r02_bench_llama7b_${N}gpu.err; echo rc=$?
tail -c 700 gpurun_out/r02_bench_llama7b_${N}gpu.json
