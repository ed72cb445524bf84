cd /root/repo
export PROF_BATCH=27 PROF_FORWARD=0 PROF_SWEEPS=2
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none -k regex:"solve_tri|gram_tc|update_tc" -s 160 -c 8 --csv python scripts/prof_one.py 2>/dev/null | grep -E "solve_tri|gram_tc|update_tc" | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-40,150-260
