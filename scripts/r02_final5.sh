set -x
cd /root/repo
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo bench rc=$?
python - <<'PY'
import json
def last(p): return json.loads(open(p).read().strip().splitlines()[-1])
d=last("gpurun_out/r02_bench_1gpu.json")
print("value", d["value"], "e2e", d["e2e"]["value"])
print("roofline", d["roofline"]["frac"], d["roofline"]["implementation_frac"], d["roofline"]["avg_launch_us"], d["roofline"]["solve_share_of_round"], "hiccup", d["run"]["hiccup"], d["clocks"])
print(d["roofline"]["class_ms_full_factorisation"], d["roofline"]["class_ms_first_2_sweeps"], d["run"]["decaying_spectrum_input"]["matrices_per_s"])
print({k:(v["us"], v["speedup_vs_cublas_pair"]) for k,v in d["extras"]["forward_config4"]["per_rank"].items()}, d["extras"]["llama7b_config3"]["seconds"])
PY
