set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "mixed_convergence or sharded_final or batched_equals or fused_kernel or gram_precond" 2>&1 | tail -25 | tee gpurun_out/r02_pytest_mixed.log
ASVD_B200_FWD=fused timeout 120 python scripts/fwd_ab.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if 'check' in d:
        if d['fwd'] == 'fused': print('CHECK', d['check'], d['dtype'][-8:], round(d['max_abs_err'], 5))
    elif d.get('r', 0) <= 256 or 'cublas' in d.get('impl', ''): print(d)
" | tee gpurun_out/r02_fwd_ab4.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_try3.json 2> gpurun_out/r02_bench_try3.err; echo bench rc=$?; tail -3 gpurun_out/r02_bench_try3.err
