"""Wall clock per matrix of asvd_scaled_svd at several batch sizes for one shape (default path), best of 2 timed runs.
  python scripts/ab_batch.py 11008x4096 4 9"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
m, n = (int(v) for v in sys.argv[1].split("x"))
for B in (int(v) for v in sys.argv[2:]):
    g = torch.Generator(device=dev).manual_seed(233)
    Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
    Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
    f = _lib.scaled_svd(Ws, Ss); torch.cuda.synchronize()
    ts = []
    for _ in range(2):
        t0 = time.perf_counter(); f = _lib.scaled_svd(Ws, Ss); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(json.dumps({"shape": [m, n], "batch": B, "solve": os.environ.get("ASVD_B200_SOLVE", "default"), "ms_per_matrix": round(min(ts) * 1e3 / B, 2),
                      "sweeps": f.sweeps}), flush=True)
    del Ws, Ss, f
    torch.cuda.empty_cache()
