set -x
timeout 120 python scripts/ab_co.py 4096x4096 18
for b in 9 18; do ASVD_B200_SOLVE=quad PROF_BATCH=$b timeout 100 python scripts/time_classes.py; ASVD_B200_SOLVE=lean PROF_BATCH=$b timeout 100 python scripts/time_classes.py; done
timeout 200 python scripts/solve_timing.py
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "lean_solve or golden_cases or batched_equals or mixed_convergence or full_size_4096 or overlapped" 2>&1 | tail -3
