PROF='ncu --set full --clock-control none --import-source on'
PROF_FORWARD=0 PROF_SWEEPS=6 $PROF -k regex:update_tc -s 300 -c 1 -o gpurun_out/p_update -f python scripts/prof_one.py > gpurun_out/p_update.log 2>&1
PROF_FORWARD=0 PROF_SWEEPS=6 $PROF -k regex:solve_quad -s 300 -c 1 -o gpurun_out/p_solve -f python scripts/prof_one.py > gpurun_out/p_solve.log 2>&1
PROF_FORWARD=0 PROF_SWEEPS=6 $PROF -k regex:gram_tc -s 300 -c 1 -o gpurun_out/p_gram -f python scripts/prof_one.py > gpurun_out/p_gram.log 2>&1
python scripts/ab_solve.py
tail -n 2 gpurun_out/p_gram.log
