cd /root/repo
PROF_BATCH=27 timeout 300 python scripts/check_tri.py tri 2>&1 | tee gpurun_out/r02_tri_b27.log
