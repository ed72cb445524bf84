set -x
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/r02_pytest_gpu_tri.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r02_bench_tri_try.json 2> gpurun_out/r02_bench_tri_try.err
tail -c 3000 gpurun_out/r02_bench_tri_try.json
