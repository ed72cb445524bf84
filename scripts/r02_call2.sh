set -x
timeout 240 python scripts/fwd_ab.py > gpurun_out/r02_fwd_ab2.jsonl 2> gpurun_out/r02_fwd_ab2.err; echo fwd rc=$?; grep -c check gpurun_out/r02_fwd_ab2.jsonl; python - <<'PY'
import json
for l in open("gpurun_out/r02_fwd_ab2.jsonl"):
    d = json.loads(l)
    if "check" in d:
        if d["max_abs_err"] > 1e-3 * (2 if "bfloat" in d["dtype"] else 1) * 3: print("BAD", d)
    else: print(d)
PY
tail -5 gpurun_out/r02_fwd_ab2.err
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "forward" 2>&1 | tail -5 | tee gpurun_out/r02_pytest_forward2.log
