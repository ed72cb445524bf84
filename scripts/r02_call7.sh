set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_try2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_try2.json 2> gpurun_out/r02_bench_try2.err; echo bench rc=$?; tail -3 gpurun_out/r02_bench_try2.err
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_forward.py > gpurun_out/r02_sanitizer_memcheck_forward.log 2>&1; echo memcheck forward rc=$?; tail -4 gpurun_out/r02_sanitizer_memcheck_forward.log
PROF='ncu --set full --clock-control none --import-source on'
PROF_BATCH=18 PROF_FORWARD=0 PROF_SWEEPS=6 timeout 300 $PROF -k regex:update_tc -s 300 -c 1 -o gpurun_out/r02_ncu_update_b18 -f python scripts/prof_one.py > /dev/null 2>&1
PROF_BATCH=18 PROF_FORWARD=0 PROF_SWEEPS=6 timeout 300 $PROF -k regex:gram_tc -s 300 -c 1 -o gpurun_out/r02_ncu_gram_precise_b18 -f python scripts/prof_one.py > /dev/null 2>&1
PROF_BATCH=18 PROF_FORWARD=0 PROF_SWEEPS=1 timeout 300 $PROF -k regex:gram_tc -s 10 -c 1 -o gpurun_out/r02_ncu_gram_single_b18 -f python scripts/prof_one.py > /dev/null 2>&1
PROF_BATCH=18 PROF_FORWARD=0 PROF_SWEEPS=1 timeout 300 $PROF -k regex:solve_quad_kernel -s 10 -c 1 -o gpurun_out/r02_ncu_solve_quad_b18 -f python scripts/prof_one.py > /dev/null 2>&1
ls -la gpurun_out/*b18*
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.json 2>/dev/null; echo launches rc=$?; wc -l gpurun_out/launches_r02.csv
