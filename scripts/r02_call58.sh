cd /root/repo
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
PROF_BATCH=27 timeout 300 python scripts/check_tri.py tri 2>&1 | tee gpurun_out/r02_tri_b27.log | cut -c1-100,380-800
SAN_TRI=1 SAN_SWEEPS=1 timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck_tri.full.log 2>&1
grep -c "Race reported" gpurun_out/r02_sanitizer_racecheck_tri.full.log
grep "Race reported\|hazard" gpurun_out/r02_sanitizer_racecheck_tri.full.log | sed 's/.*between//' | cut -c1-120 | sort | uniq -c | sort -rn | head -12 > gpurun_out/r02_sanitizer_racecheck_tri.summary.txt
grep "RACECHECK SUMMARY" gpurun_out/r02_sanitizer_racecheck_tri.full.log >> gpurun_out/r02_sanitizer_racecheck_tri.summary.txt
cat gpurun_out/r02_sanitizer_racecheck_tri.summary.txt | cut -c1-200
