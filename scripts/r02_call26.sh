set -x
for i in 1 2; do ASVD_BENCH_NO_SAMPLER=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no sampler', d['value'], d['run']['step_ms'])"; done
for i in 1 2; do timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sampler   ', d['value'], d['run']['step_ms'])"; done
