cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" 2>&1 | tail -3
PROF_BATCH=27 timeout 300 python scripts/check_tri.py tri 2>&1 | tee gpurun_out/r02_tri_b27.log | cut -c1-100,380-800
python scripts/tri_timing.py 2>&1 | head -10 | tee gpurun_out/r02_tri_timing3.log
