# First GPU call of the next round (about 4 GPU-minutes): decides whether the lean solve and two-wave batches become
# defaults.  Everything goes to gpurun_out/r02_*.
set -x
# 1. parity of the lean solve (gated test) and a memcheck / racecheck pass over it
ASVD_B200_TEST_LEAN=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k lean 2>&1 | tail -5 | tee gpurun_out/r02_lean_test.log
SAN_LEAN=1 timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r02_lean_memcheck.log 2>&1; echo memcheck rc=$?
SAN_LEAN=1 SAN_SWEEPS=1 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r02_lean_racecheck.log 2>&1; echo racecheck rc=$?
# 2. timings, three repetitions: square at 4 / 8 / 9 with both solves
AB_REPS=3 timeout 120 python scripts/ab_lean.py quad:4 quad:8 quad:9 lean:8 lean:9 lean:10 2>&1 | tee gpurun_out/r02_ab_lean.jsonl
# 3. the Llama rectangles at one and two waves (Gram pre-conditioner inside), quad and lean
# (default = odd-even outer solve on 2.7:1 shapes, quad inside the pre-conditioner; lean forces the lean solve in both)
timeout 120 python scripts/ab_overlap.py 11008x4096x4 11008x4096x8 4096x11008x4 4096x11008x8 2>&1 | tee gpurun_out/r02_rect_default.jsonl
ASVD_B200_SOLVE=lean timeout 120 python scripts/ab_overlap.py 11008x4096x4 11008x4096x8 4096x11008x4 4096x11008x8 2>&1 | tee gpurun_out/r02_rect_lean.jsonl
# 4. the bench with the candidate defaults (lean where the batch has more pairs than SMs, nine weights per step)
ASVD_B200_LEAN_AUTO=1 timeout 200 python bench.py --batch 9 --no-cpu-baseline > gpurun_out/r02_bench_1gpu_lean9.json 2> gpurun_out/r02_bench_1gpu_lean9.err; tail -c 300 gpurun_out/r02_bench_1gpu_lean9.json
# 5. the bench as the driver runs it
timeout 200 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 500 gpurun_out/r02_bench_1gpu.json
