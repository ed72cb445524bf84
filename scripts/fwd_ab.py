"""Forward (a7) parity + timing at BASELINE config 4 for the CTA-pair kernel, the single-CTA kernel and the cuBLAS pair
(upstream's module on the same box).  One JSON line per measurement."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
torch.manual_seed(0)


def check(M, n, r, m, bias=True, dtype=torch.float16):
    x = (torch.randn(M, n, device=dev) * 0.125).to(dtype); B = (torch.randn(r, n, device=dev) / n ** 0.5).to(dtype)
    A = (torch.randn(m, r, device=dev) / r ** 0.5 * 0.5).to(dtype); b = (torch.randn(m, device=dev) * 0.1).to(dtype) if bias else None
    y = _lib.lowrank_forward(x, A, B, b)
    t = (x.double() @ B.double().t()).to(dtype).double()
    ref = t @ A.double().t() + (0 if b is None else b.double())
    err = (y.double() - ref).abs().max().item()
    print(json.dumps({"check": [M, n, r, m], "dtype": str(dtype), "bias": bias, "max_abs_err": err, "ymax": ref.abs().max().item(),
                      "fwd": os.environ.get("ASVD_B200_FWD", "default"), "bn": os.environ.get("ASVD_B200_FWD_BN")}), flush=True)
    return err


if "--no-check" not in sys.argv:
    os.environ["ASVD_B200_FWD"] = "fused"                 # fused kernel for r <= 256
    for shp in [(128, 64, 64, 128), (1, 64, 8, 32), (256, 512, 128, 256), (300, 512, 100, 384), (1000, 1024, 200, 1000),
                (2048, 4096, 256, 4096), (2048 * 3 + 77, 768, 153, 3072), (19000, 1024, 256, 1024), (40000, 512, 64, 520)]:
        check(*shp)
    check(512, 1024, 256, 1024, dtype=torch.bfloat16)
    check(513, 1024, 77, 1024, bias=False, dtype=torch.bfloat16)
    os.environ["ASVD_B200_FWD"] = "pair"
    for bn in (None, "64", "192"):
        if bn is None: os.environ.pop("ASVD_B200_FWD_BN", None)
        else: os.environ["ASVD_B200_FWD_BN"] = bn
        for shp in [(128, 64, 64, 128), (1, 64, 8, 32), (256, 512, 128, 256), (300, 512, 100, 384), (1000, 1024, 345, 1000),
                    (700, 4096, 1843, 4096), (4096, 4096, 512, 4096)]:
            check(*shp)
    os.environ.pop("ASVD_B200_FWD_BN", None)
    check(512, 1024, 256, 1024, dtype=torch.bfloat16)
    check(513, 1024, 77, 1024, bias=False, dtype=torch.bfloat16)

M, n, m = 32 * 2048, 4096, 4096
x = (torch.randn(M, n, device=dev) * 0.125).half()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()                                   # evict L2 between iterations
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3


for r in (128, 256, 512, 1024, 1843):
    B = (torch.randn(r, n, device=dev) / n ** 0.5).half(); A = (torch.randn(m, r, device=dev) / r ** 0.5).half()
    Ak = _lib.pad_rank_stride(A)
    fl = 2.0 * M * r * (n + m)
    cfgs = [("fused", None), ("pair", None), ("1cta", None)] if r <= 256 else [("pair", None), ("1cta", None)]
    for fwd, bn in cfgs:
        os.environ["ASVD_B200_FWD"] = fwd
        if bn is None: os.environ.pop("ASVD_B200_FWD_BN", None)
        else: os.environ["ASVD_B200_FWD_BN"] = bn
        if fwd == "1cta" and r % 8: continue
        med, best = timeit(lambda: _lib.lowrank_forward(x, A, B, None, A_kernel=Ak))
        print(json.dumps({"r": r, "impl": fwd, "bn": bn, "us_median": round(med, 1), "us_best": round(best, 1), "tflops_median": round(fl / med / 1e6, 1)}), flush=True)
    os.environ.pop("ASVD_B200_FWD", None); os.environ.pop("ASVD_B200_FWD_BN", None)
    med, best = timeit(lambda: torch.nn.functional.linear(torch.nn.functional.linear(x, B), A))
    print(json.dumps({"r": r, "impl": "cublas_pair(torch F.linear x2)", "us_median": round(med, 1), "us_best": round(best, 1), "tflops_median": round(fl / med / 1e6, 1)}), flush=True)
