"""A/B wall-clock of a batch-4 4096x4096 factorisation under an environment switch (diagnostic)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(233)
B = int(os.environ.get("PROF_BATCH", "4"))
n = int(os.environ.get("PROF_N", "4096")); m = int(os.environ.get("PROF_M", "4096"))
Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
f = _lib.scaled_svd(Ws, Ss)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    f = _lib.scaled_svd(Ws, Ss)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print(os.environ.get("ASVD_B200_SOLVE", "quad"), f"{m}x{n} batch {B}", "ms:", [round(t * 1e3, 1) for t in ts], "sweeps", f.sweeps)
