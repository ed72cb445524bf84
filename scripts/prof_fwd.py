"""One forward at BASELINE config 4 (r from PROF_R) for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(233)
r = int(os.environ.get("PROF_R", "256")); M, n, m = 32 * 2048, 4096, 4096
x = (torch.randn(M, n, device=dev, generator=g) * 0.125).half()
B = (torch.randn(r, n, device=dev, generator=g) / n ** 0.5).half(); A = (torch.randn(m, r, device=dev, generator=g) / r ** 0.5).half()
for _ in range(3):
    y = _lib.lowrank_forward(x, A, B, None)
torch.cuda.synchronize()
print("done")
