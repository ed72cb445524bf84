set -x
cd /root/repo
export ASVD_B200_SOLVE=tri PROF_BATCH=9 PROF_FORWARD=0 PROF_SWEEPS=2
PROF='ncu --set full --clock-control none --import-source on'
$PROF -k regex:solve_tri_g -s 40 -c 1 -o gpurun_out/p_tri_g -f python scripts/prof_one.py > gpurun_out/p_tri_g.log 2>&1
$PROF -k regex:solve_quad_r -s 40 -c 1 -o gpurun_out/p_tri_r -f python scripts/prof_one.py > gpurun_out/p_tri_r.log 2>&1
tail -n 2 gpurun_out/p_tri_g.log gpurun_out/p_tri_r.log
