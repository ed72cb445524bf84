set -x
PROF_R=256 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lowrank_fused -s 2 -c 1 -o gpurun_out/r02_ncu_fwd_fused256 python scripts/prof_fwd.py 2>&1 | tail -3
ASVD_B200_FWD=pair PROF_R=256 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn2 -s 4 -c 2 -o gpurun_out/r02_ncu_fwd_pair256 python scripts/prof_fwd.py 2>&1 | tail -3
ASVD_B200_FWD=1cta PROF_R=256 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 4 -c 2 -o gpurun_out/r02_ncu_fwd_1cta256 python scripts/prof_fwd.py 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
