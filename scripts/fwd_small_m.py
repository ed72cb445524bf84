"""Forward latency at small / medium token counts (decode, one calibration sample, a batch of eight): CTA-pair GEMMs vs
the fused kernel vs the cuBLAS pair.  d = 4096."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"; n = m = 4096
def timeit(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return round(sorted(ts)[len(ts) // 2], 1)
for r in (128, 256):
    g = torch.Generator(device=dev).manual_seed(r)
    B = (torch.randn(r, n, device=dev, generator=g) / n ** 0.5).half(); A = (torch.randn(m, r, device=dev, generator=g) / r ** 0.5).half()
    for M in (16, 256, 2048, 8192, 16384):
        x = (torch.randn(M, n, device=dev, generator=g) * 0.125).half()
        out = {"r": r, "M": M}
        for mode in ("pair", "fused"):
            os.environ["ASVD_B200_FWD"] = mode
            out[mode + "_us"] = timeit(lambda: _lib.lowrank_forward(x, A, B, None))
        os.environ.pop("ASVD_B200_FWD")
        out["cublas_pair_us"] = timeit(lambda: torch.nn.functional.linear(torch.nn.functional.linear(x, B), A))
        print(json.dumps(out), flush=True)
