"""Small factorisations + a forward for compute-sanitizer runs (memcheck / racecheck / synccheck).
Covers: batched and single launches, the Gram pre-conditioned tall shape, shapes far below one 64-vector block
(single row / column, odd sizes), the overlapped half-batch schedule (ASVD_B200_OVERLAP=1), the three calibration
statistics on ragged shapes, and ragged forwards."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
sweeps = int(os.environ.get("SAN_SWEEPS", "3"))
cases = [(512, 384, 2, "0"), (384, 640, 1, "0"), (2304, 1024, 1, "0"), (512, 512, 3, "1"), (640, 384, 2, "1"),
         (1, 5, 1, "0"), (5, 1, 1, "0"), (7, 13, 1, "0"), (129, 65, 2, "1"), (1, 300, 1, "0"), (300, 1, 1, "0")]
if os.environ.get("SAN_LEAN") == "1":           # the experimental lean solve (racecheck it before it becomes a default)
    os.environ["ASVD_B200_SOLVE"] = "lean"
    cases = [(512, 512, 2, "0"), (384, 640, 1, "0"), (1024, 1024, 3, "1"), (129, 65, 2, "0")]
if os.environ.get("SAN_TRI") == "1":            # the default (triangular) solve alone: short list for racecheck / synccheck
    os.environ.pop("ASVD_B200_SOLVE", None)
    cases = [(512, 512, 2, "0"), (384, 640, 1, "0"), (1024, 1024, 3, "0"), (129, 65, 2, "0")]
for (m, n, B, overlap) in cases:
    os.environ["ASVD_B200_OVERLAP"] = overlap
    Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
    Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
    f = _lib.scaled_svd(Ws, Ss, max_sweeps=sweeps, allow_status=(0, 5))
    r = max(1, min(m, n) // 2)
    for fuse in ("UV", "V"):
        A, Bm = f.extract(r, fuse, torch.float16, B - 1)
    f.sigma(0)
    x = (torch.randn(300, n, device=dev, generator=g) * 0.125).half()
    y = _lib.lowrank_forward(x, A, Bm, None)
for dt in (torch.float16, torch.float32, torch.bfloat16):
    for (L, n) in [(333, 1001), (1, 7), (2048, 768)]:
        x = torch.randn(L, n, device=dev, generator=g).to(dt)
        for method in ("abs_mean", "abs_max", "sq_mean"):
            acc = torch.zeros(n, dtype=dt, device=dev)
            _lib.absstat_accum(x, acc, method)
torch.cuda.synchronize()
print("done")
