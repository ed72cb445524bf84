"""Small factorisations + a forward for compute-sanitizer runs (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
for (m, n, B) in [(512, 384, 2), (384, 640, 1), (2304, 1024, 1)]:
    Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
    Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
    f = _lib.scaled_svd(Ws, Ss, max_sweeps=int(os.environ.get("SAN_SWEEPS", "3")), allow_status=(0, 5))
    A, Bm = f.extract(min(m, n) // 2, "UV", torch.float16, 0)
    x = (torch.randn(300, n, device=dev, generator=g) * 0.125).half()
    y = _lib.lowrank_forward(x, A, Bm, None)
torch.cuda.synchronize()
print("done")
