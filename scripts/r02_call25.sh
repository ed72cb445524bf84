set -x
ls oracle/_ref/modules/
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['cpu_baseline']['kind'], d['run'])"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['value'], d['run']['hiccup'], d['clocks'])"
