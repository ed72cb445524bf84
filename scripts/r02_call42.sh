set -x
cd /root/repo
timeout 300 python scripts/check_tri.py small 2>&1 | tee gpurun_out/r02_tri_small.log
timeout 300 python scripts/check_tri.py big 2>&1 | tee gpurun_out/r02_tri_big.log
export ASVD_B200_SOLVE=tri PROF_BATCH=9 PROF_FORWARD=0 PROF_SWEEPS=2
PROF='ncu --set full --clock-control none --import-source on'
$PROF -k regex:solve_tri_g -s 40 -c 1 -o gpurun_out/p_tri_g -f python scripts/prof_one.py > gpurun_out/p_tri_g.log 2>&1
