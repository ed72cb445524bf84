set -x
for r in 512 1024 1843; do
  PROF_R=$r timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn2 -s 4 -c 2 -o gpurun_out/r02_ncu_fwd_pair$r -f python scripts/prof_fwd.py > /dev/null 2>&1; echo ncu $r rc=$?
done
timeout 600 python scripts/check_illcond.py 2>&1 | tee gpurun_out/r02_illcond_check.jsonl
