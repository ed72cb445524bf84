set -x
cd /root/repo
( time timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/r02_bench_1gpu.err
( time timeout 900 python bench.py --workload llama7b > gpurun_out/r02_bench_llama7b_1gpu.json 2> gpurun_out/r02_bench_llama7b_1gpu.err ) 2>&1 | tail -3
tail -c 300 gpurun_out/r02_bench_llama7b_1gpu.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_smoke.log
