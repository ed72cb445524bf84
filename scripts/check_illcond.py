"""Robustness at scale: power-law spectra and outlier activation scales, direct vs Gram-pre-conditioned (diagnostic)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
def make(m, n, decay, seed, outliers):
    g = torch.Generator(device=dev).manual_seed(seed)
    k = min(m, n)
    U, _ = torch.linalg.qr(torch.randn(m, k, device=dev, generator=g, dtype=torch.float64))
    V, _ = torch.linalg.qr(torch.randn(n, k, device=dev, generator=g, dtype=torch.float64))
    sig = torch.arange(1, k + 1, device=dev, dtype=torch.float64) ** (-decay)
    W = ((U * sig) @ V.t() * 0.5).half()
    sdm = torch.exp(torch.randn(n, device=dev, generator=g))
    if outliers:
        idx = torch.randperm(n, device=dev, generator=g)[: n // 100]
        sdm[idx] *= 100.0
    return W, _lib.scaling_vector(sdm.half(), None, 0.5, n, dev)
cases = [(4096, 4096, 1.0, False), (4096, 4096, 1.5, True), (11008, 4096, 1.0, True), (4096, 11008, 1.5, True), (11008, 4096, 0.0, True)]
for (m, n, decay, outl) in cases:
    W, s = make(m, n, decay, 5, outl)
    ref = torch.linalg.svdvals(W.double() * s.double(), driver="gesvd")
    r = _lib.rank_for_ratio(m, n, 0.9, 1)
    for pre in ("1", "0"):
        if pre == "0" and max(m, n) < 2 * min(m, n):
            continue
        os.environ["ASVD_B200_GRAMPRE"] = pre
        t0 = time.perf_counter()
        try:
            f = _lib.scaled_svd([W], [s]); torch.cuda.synchronize()
        except Exception as e:
            print(json.dumps({"shape": [m, n], "decay": decay, "gram_pre": pre, "error": str(e)[:200]}), flush=True)
            continue
        dt = time.perf_counter() - t0
        sig = f.sigma(0).double()
        rel_all = ((sig[:r] - ref[:r]).abs() / ref[:r]).max().item()
        big = ref[:r] > 1e-3 * ref[0]
        rel_big = ((sig[:r] - ref[:r]).abs() / ref[:r])[big].max().item()
        A, B = f.extract(r, "UV", torch.float32, 0)
        rec = ((A.double() @ B.double() - W.double()) * s.double()).norm().item()
        floor = (ref[r:] ** 2).sum().sqrt().item()
        print(json.dumps({"shape": [m, n], "decay": decay, "outliers": outl, "gram_pre": pre, "ms": round(dt * 1e3, 1), "sweeps": list(f.sweeps),
                          "status": f.status, "cond_kept": (ref[0] / ref[r - 1]).item(), "sigma_rel_err_kept": rel_all,
                          "sigma_rel_err_above_1e-3": rel_big, "recon_over_floor": rec / floor}), flush=True)
