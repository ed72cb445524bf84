cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lean_solve or tri_solve or inner_orderings or mixed_convergence" 2>&1 | tail -3
PROF_BATCH=27 timeout 300 python scripts/check_tri.py tri 2>&1 | tee gpurun_out/r02_tri_b27.log | cut -c1-100,380-800
export ASVD_B200_SOLVE=tri PROF_BATCH=9 PROF_FORWARD=0 PROF_SWEEPS=2
ncu --set full --clock-control none --import-source on -k regex:solve_tri_r -s 40 -c 1 -o gpurun_out/p_tri_r -f python scripts/prof_one.py > /dev/null 2>&1
ncu -i gpurun_out/p_tri_r.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for w in ['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','sm__cycles_elapsed.max']:
    print(w, r[h.index(w)])
"
