"""Host cost of one SVDLinear.forward call at tiny M (back-to-back calls, one sync at the end) next to torch's pair."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
from asvd4llm_b200.modules.svd_linear import SVDLinear
dev = "cuda"; n = m = 4096; r = 256
g = torch.Generator(device=dev).manual_seed(1)
B = (torch.randn(r, n, device=dev, generator=g) / 64).half(); A = (torch.randn(m, r, device=dev, generator=g) / 16).half()
mod = SVDLinear._from_factors(A, B, None)
for M in (16, 2048):
    x = (torch.randn(M, n, device=dev, generator=g) * 0.125).half()
    for name, fn in (("ours(module)", lambda: mod(x)), ("ours(_lib)", lambda: _lib.lowrank_forward(x, A, B, None)),
                     ("torch pair", lambda: torch.nn.functional.linear(torch.nn.functional.linear(x, B), A))):
        with torch.no_grad():
            for _ in range(20): fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(500): fn()
            t_issue = time.perf_counter() - t0
            torch.cuda.synchronize()
            t_all = time.perf_counter() - t0
        print(json.dumps({"M": M, "impl": name, "host_issue_us_per_call": round(t_issue / 500 * 1e6, 1), "total_us_per_call": round(t_all / 500 * 1e6, 1)}), flush=True)
