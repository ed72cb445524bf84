"""Convergence trace (ASVD_B200_TRACE=1) of the Gaussian and the decaying-spectrum bench inputs, two weights each, plus
the kept-sigma error vs fp64 for a set of ASVD_B200_TOL_PRE values."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"; n = 4096
g = torch.Generator(device=dev).manual_seed(233)
def gauss(): return (torch.randn(n, n, device=dev, generator=g) * 0.02).half()
def decay():
    u = torch.randn(n, 64, device=dev, generator=g); v = torch.randn(64, n, device=dev, generator=g)
    return ((torch.randn(n, n, device=dev, generator=g) + (u * torch.logspace(0, -2, 64, device=dev) * 8.0) @ v) * 0.02).half()
for name, mk in (("gauss", gauss), ("decay", decay)):
    Ws = [mk() for _ in range(2)]
    Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in Ws]
    ref = [torch.linalg.svdvals(W.double() * s.double()) for W, s in zip(Ws, Ss)]
    for tol_pre in (None, "1e-4", "5e-4", "2e-3"):
        if tol_pre is None: os.environ.pop("ASVD_B200_TOL_PRE", None)
        else: os.environ["ASVD_B200_TOL_PRE"] = tol_pre
        os.environ["ASVD_B200_TRACE"] = "1" if tol_pre is None else "0"
        if tol_pre is not None: os.environ.pop("ASVD_B200_TRACE")
        print(f"--- {name} tol_pre={tol_pre}", file=sys.stderr, flush=True)
        f = _lib.scaled_svd(Ws, Ss)
        r = 1843
        errs, recs = [], []
        for b in range(2):
            sig = f.sigma(b).double()
            errs.append(((sig[:r] - ref[b][:r]).abs() / ref[b][:r]).max().item())
            A, Bm = f.extract(r, "UV", torch.float32, b)
            rec = ((A.double() @ Bm.double() - Ws[b].double()) * Ss[b].double()).norm().item()
            floor = (ref[b][r:] ** 2).sum().sqrt().item()
            recs.append(rec / floor - 1.0)
        print(json.dumps({"input": name, "tol_pre": tol_pre, "sweeps": f.sweeps, "kept_sigma_rel_err": errs, "recon_over_eckart_young_minus_1": recs}), flush=True)
