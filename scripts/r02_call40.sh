set -x
cd /root/repo
timeout 300 python scripts/check_tri.py small 2>&1 | tee gpurun_out/r02_tri_small.log
timeout 300 python scripts/check_tri.py big 2>&1 | tee gpurun_out/r02_tri_big.log
