"""One short batch-4 4096x4096 factorisation (2 sweeps) + one forward, for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(233)
B = int(os.environ.get("PROF_BATCH", "4"))
Ws = [(torch.randn(4096, 4096, device=dev, generator=g) * 0.02).half() for _ in range(B)]
Ss = [_lib.scaling_vector(torch.exp(torch.randn(4096, device=dev, generator=g)).half(), None, 0.5, 4096, dev) for _ in range(B)]
f = _lib.scaled_svd(Ws, Ss, max_sweeps=int(os.environ.get("PROF_SWEEPS", "2")))
A, Bm = f.extract(1843, "UV", torch.float16, 0)
if os.environ.get("PROF_FORWARD", "1") == "1":
    x = (torch.randn(32 * 2048, 4096, device=dev, generator=g) * 0.125).half()
    y = _lib.lowrank_forward(x, A[:, :256].contiguous(), Bm[:256].contiguous(), None)
torch.cuda.synchronize()
print("done", f.sweeps)
