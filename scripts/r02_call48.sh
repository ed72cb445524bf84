set -x
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu.log
SAN_TRI=1 SAN_SWEEPS=2 timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_memcheck_tri.log
SAN_TRI=1 SAN_SWEEPS=1 timeout 1200 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck_tri.full.log 2>&1
grep -c "Race reported\|hazard" gpurun_out/r02_sanitizer_racecheck_tri.full.log
grep "Race reported\|hazard" gpurun_out/r02_sanitizer_racecheck_tri.full.log | sed 's/.*between//' | sort | uniq -c | sort -rn | head -20 > gpurun_out/r02_sanitizer_racecheck_tri.summary.txt
grep -A3 "RACECHECK SUMMARY\|ERROR SUMMARY" gpurun_out/r02_sanitizer_racecheck_tri.full.log | tail -5 >> gpurun_out/r02_sanitizer_racecheck_tri.summary.txt
head -c 3000 gpurun_out/r02_sanitizer_racecheck_tri.full.log > gpurun_out/r02_sanitizer_racecheck_tri.head.log
rm -f gpurun_out/r02_sanitizer_racecheck_tri.full.log.big
PROF='ncu --set full --clock-control none --import-source on'
export PROF_BATCH=27 PROF_FORWARD=0
PROF_SWEEPS=2 timeout 300 $PROF -k regex:solve_tri_g -s 40 -c 1 -o gpurun_out/r02_ncu_solve_tri_g_b27 -f python scripts/prof_one.py > /dev/null 2>&1
PROF_SWEEPS=2 timeout 300 $PROF -k regex:solve_tri_r -s 40 -c 1 -o gpurun_out/r02_ncu_solve_tri_r_b27 -f python scripts/prof_one.py > /dev/null 2>&1
PROF_SWEEPS=2 timeout 300 $PROF -k regex:update_tc -s 40 -c 1 -o gpurun_out/r02_ncu_update_b27 -f python scripts/prof_one.py > /dev/null 2>&1
PROF_SWEEPS=1 timeout 300 $PROF -k regex:gram_tc -s 10 -c 1 -o gpurun_out/r02_ncu_gram_single_b27 -f python scripts/prof_one.py > /dev/null 2>&1
PROF_SWEEPS=6 timeout 400 $PROF -k regex:gram_tc -s 320 -c 1 -o gpurun_out/r02_ncu_gram_precise_b27 -f python scripts/prof_one.py > /dev/null 2>&1
ls -la gpurun_out/*b27*
