"""Timing of the experimental lean solve against the quad solve at several batch sizes (4096^2 fp16), one warm-up and
one timed factorisation each; prints a line per configuration as soon as it is known."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
n = int(os.environ.get("PROF_N", "4096"))
g = torch.Generator(device=dev).manual_seed(233)
Wall = [(torch.randn(n, n, device=dev, generator=g) * 0.02).half() for _ in range(9)]
Sall = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(9)]
cfgs = [("quad", 4), ("lean", 4), ("lean", 9), ("lean", 8), ("quad", 8)]
if len(sys.argv) > 1:
    cfgs = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]]
reps = int(os.environ.get("AB_REPS", "1"))
for solve, B in cfgs:
    os.environ["ASVD_B200_SOLVE"] = solve
    f = _lib.scaled_svd(Wall[:B], Sall[:B]); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f = _lib.scaled_svd(Wall[:B], Sall[:B]); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    dt = min(ts)
    print(json.dumps({"solve": solve, "batch": B, "ms": round(dt * 1e3, 1), "ms_per_matrix": round(dt * 1e3 / B, 2), "sweeps": f.sweeps}), flush=True)
