set -x
# 1. post-fix profiles of the forward kernels at r = 256
ASVD_B200_FWD=fused PROF_R=256 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lowrank_fused -s 2 -c 1 -o gpurun_out/r02_ncu_fwd_fused256b python scripts/prof_fwd.py 2>&1 | tail -2
PROF_R=256 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn2 -s 4 -c 2 -o gpurun_out/r02_ncu_fwd_pair256b python scripts/prof_fwd.py 2>&1 | tail -2
# 2. the inner solve, source-level
PROF_BATCH=9 PROF_SWEEPS=1 PROF_FORWARD=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_quad_kernel -s 5 -c 1 -o gpurun_out/r02_ncu_solve_quad_b9 python scripts/prof_one.py 2>&1 | tail -2
ASVD_B200_SOLVE=lean PROF_BATCH=9 PROF_SWEEPS=1 PROF_FORWARD=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_quad_g_kernel -s 5 -c 1 -o gpurun_out/r02_ncu_solve_lean_g_b9 python scripts/prof_one.py 2>&1 | tail -2
# 3. batch sizes on the rectangles and the square
timeout 200 python scripts/ab_batch.py 11008x4096 4 9 2>&1 | tee gpurun_out/r02_ab_batch.jsonl
timeout 200 python scripts/ab_batch.py 4096x11008 4 9 2>&1 | tee -a gpurun_out/r02_ab_batch.jsonl
timeout 200 python scripts/ab_batch.py 4096x4096 8 9 18 2>&1 | tee -a gpurun_out/r02_ab_batch.jsonl
ASVD_B200_SOLVE=lean timeout 200 python scripts/ab_batch.py 4096x4096 9 18 2>&1 | tee -a gpurun_out/r02_ab_batch.jsonl
# 4. the new bench, default line with extras
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_try1.json 2> gpurun_out/r02_bench_try1.err; echo bench rc=$?; tail -c 1500 gpurun_out/r02_bench_try1.json; tail -5 gpurun_out/r02_bench_try1.err
# 5. the whole GPU suite
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_try1.log
