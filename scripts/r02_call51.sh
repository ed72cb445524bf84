cd /root/repo
python scripts/ab_env.py ASVD_B200_NEAR_PCT unset 50 70 80 -- 4096x4096x27 2048x2048x16 768x768x32 4096x11008x8 2>&1 | tee gpurun_out/r02_ab_near_pct_tri2.jsonl
for v in unset 60; do
  if [ $v = unset ]; then unset ASVD_B200_NEAR_PCT; else export ASVD_B200_NEAR_PCT=$v; fi
  echo "== NEAR_PCT $v" | tee -a gpurun_out/r02_illcond_near_pct.jsonl
  timeout 600 python scripts/check_illcond.py 2>&1 | tee -a gpurun_out/r02_illcond_near_pct.jsonl
done
