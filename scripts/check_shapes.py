"""Accuracy + time of one batch per shape (Llama-2-13B and OPT-125m shapes by default) against an fp64 SVD (diagnostic)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ.get("SHAPES", "5120x5120,13824x5120,5120x13824,768x768,3072x768,768x3072,50272x768").split(",")]
ratio = float(os.environ.get("RATIO", "0.95"))
out = []
for m, n in shapes:
    B = min(_lib.suggest_batch(m, n, dev), int(os.environ.get("MAXB", "4")))
    g = torch.Generator(device=dev).manual_seed(233)
    Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
    Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
    f = _lib.scaled_svd(Ws, Ss); torch.cuda.synchronize()
    t0 = time.perf_counter(); f = _lib.scaled_svd(Ws, Ss); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    r = _lib.rank_for_ratio(m, n, ratio, 1)
    ref = torch.linalg.svdvals(Ws[0].double() * Ss[0].double(), driver="gesvd")
    sig = f.sigma(0).double()
    rel = ((sig[:r] - ref[:r]).abs() / ref[:r]).max().item()
    A, Bm = f.extract(r, "UV", torch.float32, 0)
    Wd = Ws[0].double()
    rec = (A.double() @ Bm.double() - Wd) * Ss[0].double()
    best = (ref[r:] ** 2).sum().sqrt().item()
    out.append({"shape": [m, n], "batch": B, "ms_per_matrix": round(dt / B * 1e3, 2), "rank": r, "sweeps": list(f.sweeps),
                "sigma_rel_err_vs_fp64": rel, "scaled_recon_err": rec.norm().item(), "eckart_young_floor": best})
    print(json.dumps(out[-1]), flush=True)
