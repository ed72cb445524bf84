set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python scripts/fwd_ab.py > gpurun_out/r02_fwd_ab.jsonl 2> gpurun_out/r02_fwd_ab.err; echo fwd rc=$?; tail -30 gpurun_out/r02_fwd_ab.jsonl; tail -5 gpurun_out/r02_fwd_ab.err
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "forward" 2>&1 | tail -5 | tee gpurun_out/r02_pytest_forward.log
for s in quad lean; do for b in 8 9 16; do ASVD_B200_SOLVE=$s PROF_BATCH=$b timeout 100 python scripts/time_classes.py; done; done 2>&1 | tee gpurun_out/r02_classes_lean_quad.log
timeout 200 python scripts/gpu_baselines.py 2>&1 | tee gpurun_out/r02_gpu_baselines.jsonl
