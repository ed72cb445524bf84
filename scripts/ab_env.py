"""A/B wall-clock of one environment switch, in one process: python scripts/ab_env.py VAR v1 v2 ... [-- MxNxB ...]
(diagnostic; prints one JSON line per shape with times, sweeps and the sigma error against fp64 per value)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
args = sys.argv[1:]
shapes = [(4096, 4096, 4)]
if "--" in args:
    k = args.index("--")
    shapes = [tuple(int(v) for v in a.split("x")) for a in args[k + 1:]]
    args = args[:k]
var, values = args[0], args[1:]
dev = "cuda"
for m, n, B in shapes:
    g = torch.Generator(device=dev).manual_seed(233)
    Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
    Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
    ref = torch.linalg.svdvals(Ws[0].double() * Ss[0].double())
    out = {"shape": [m, n], "batch": B, "var": var}
    for v in values:
        if v == "unset":
            os.environ.pop(var, None)
        else:
            os.environ[var] = v
        f = _lib.scaled_svd(Ws, Ss)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            f = _lib.scaled_svd(Ws, Ss)
            torch.cuda.synchronize()
            ts.append(round((time.perf_counter() - t0) * 1e3, 1))
        err = ((f.sigma(0).double() - ref).abs() / ref).max().item()
        out[v] = {"ms": ts, "sweeps": f.sweeps, "sigma_rel_err_fp64": err}
    print(json.dumps(out), flush=True)
