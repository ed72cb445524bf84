set -x
ASVD_B200_FWD=fused PROF_R=256 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lowrank_fused -s 2 -c 1 -o gpurun_out/r02_ncu_fwd_fused256c -f python scripts/prof_fwd.py 2>&1 | tail -1
