set -x
python bench.py > gpurun_out/r01_bench_1gpu.json 2> gpurun_out/r01_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_r01.bench.log 2>&1
PROF='ncu --set full --clock-control none --import-source on'
PROF_FORWARD=0 PROF_SWEEPS=6 $PROF -k regex:update_tc -s 300 -c 1 -o gpurun_out/p_update -f python scripts/prof_one.py > gpurun_out/p_update.log 2>&1
PROF_FORWARD=0 PROF_SWEEPS=6 $PROF -k regex:solve_quad -s 300 -c 1 -o gpurun_out/p_solve -f python scripts/prof_one.py > gpurun_out/p_solve.log 2>&1
PROF_FORWARD=0 PROF_SWEEPS=6 $PROF -k regex:gram_tc -s 300 -c 1 -o gpurun_out/p_gram -f python scripts/prof_one.py > gpurun_out/p_gram.log 2>&1
PROF_FORWARD=0 PROF_SWEEPS=1 $PROF -k regex:gram_tc -s 20 -c 1 -o gpurun_out/p_gram_single -f python scripts/prof_one.py > gpurun_out/p_gram_single.log 2>&1
$PROF -k regex:gemm_tn -s 4 -c 2 -o gpurun_out/p_fwd -f python scripts/prof_fwd.py > gpurun_out/p_fwd.log 2>&1
python scripts/check_shapes.py > gpurun_out/r01_shapes_check.jsonl 2>&1
python scripts/gpu_forward_check.py > gpurun_out/r01_forward_check.log 2>&1
tail -c 600 gpurun_out/r01_bench_1gpu.json
