set -x
for b in 9 18; do
  ASVD_B200_SOLVE=quad PROF_BATCH=$b timeout 100 python scripts/time_classes.py
  ASVD_B200_SOLVE=lean PROF_BATCH=$b timeout 100 python scripts/time_classes.py
  ASVD_B200_SOLVE=lean ASVD_B200_LEAN_ONE=1 PROF_BATCH=$b timeout 100 python scripts/time_classes.py
done 2>&1 | grep -v "^+" | tee gpurun_out/r02_lean_one_cta.log
