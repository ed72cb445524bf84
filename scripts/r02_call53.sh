cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" 2>&1 | tail -3
PROF_BATCH=27 timeout 300 python scripts/check_tri.py tri 2>&1 | tee gpurun_out/r02_tri_b27.log | cut -c1-100,380-800
PROF='ncu --set full --clock-control none --import-source on'
export PROF_BATCH=27 PROF_FORWARD=0
PROF_SWEEPS=8 timeout 400 $PROF -k regex:gram_tc -s 480 -c 1 -o gpurun_out/r02_ncu_gram_precise_b27 -f python scripts/prof_one.py > /dev/null 2>&1
ncu -i gpurun_out/r02_ncu_gram_precise_b27.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for w in ['gpu__time_duration.sum','dram__bytes_read.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__block_size']:
    print(w, r[h.index(w)])
"
