"""Co-residency experiment: wall clock per matrix (4096^2, batch from argv) with the library / environment given by the
caller, plus a digest of the singular values (bitwise comparison across configurations)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
m, n = (int(v) for v in sys.argv[1].split("x"))
B = int(sys.argv[2])
g = torch.Generator(device=dev).manual_seed(233)
Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
f = _lib.scaled_svd(Ws, Ss); torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); f = _lib.scaled_svd(Ws, Ss); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
dig = sum(int(f.sigma(b).view(torch.int32).long().sum().item()) for b in range(B))
print(json.dumps({"shape": [m, n], "batch": B, "lib": os.path.basename(os.environ.get("ASVD_B200_LIBPATH", "default")),
                  "overlap": os.environ.get("ASVD_B200_OVERLAP", "0"), "solve": os.environ.get("ASVD_B200_SOLVE", "default"),
                  "ms_per_matrix": round(min(ts) * 1e3 / B, 2), "all_ms": [round(t * 1e3, 1) for t in ts], "sweeps": f.sweeps[:4], "sigma_digest": dig}), flush=True)
