set -x
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck_svd.log 2>&1; echo memcheck svd rc=$?; tail -3 gpurun_out/r02_sanitizer_memcheck_svd.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo bench rc=$?; tail -2 gpurun_out/r02_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cut -c1-400 gpurun_out/r02_bench_reference_arm.json
timeout 300 python bench.py --workload forward --steps 10 --warmup 3 > gpurun_out/r02_bench_forward.json 2> gpurun_out/r02_bench_forward.err; echo fwd rc=$?; cut -c1-600 gpurun_out/r02_bench_forward.json
