cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r02_bench_try2.json 2> gpurun_out/r02_bench_try2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_try2.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step', d['ms_per_step'],'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['class_ms_full_factorisation'], d['clocks'], d['run']['decaying_spectrum_input'])
PY
