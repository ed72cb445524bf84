"""Triangular inner solve (ASVD_B200_SOLVE=tri) against the quad solve and fp64: sigma error, reconstruction, sweeps,
per-class times of two sweeps, wall clock (diagnostic)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"

def case(m, n, B, modes=("quad", "tri"), timing=True):
    g = torch.Generator(device=dev).manual_seed(233)
    Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
    Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
    ref = torch.linalg.svdvals((Ws[0].double() * Ss[0].double()[None, :]))
    out = {}
    for mode in modes:
        os.environ["ASVD_B200_SOLVE"] = mode
        f = _lib.scaled_svd(Ws, Ss)
        torch.cuda.synchronize()
        s = f.sigma(0).double()
        rel = ((s - ref).abs() / ref).max().item()
        r = min(m, n) // 2
        A, Bm = f.extract(r, "UV", torch.float32, 0)
        Wd = Ws[0].double()
        rec = (A.double() @ Bm.double() - Wd)
        U, S, Vh = torch.linalg.svd(Wd * Ss[0].double()[None, :], full_matrices=False)
        best = ((U[:, :r] * S[:r]) @ Vh[:r]) / Ss[0].double()[None, :] - Wd
        rec_s = (rec * Ss[0].double()[None, :]).norm().item(); best_s = (best * Ss[0].double()[None, :]).norm().item()
        rec_ratio = rec_s / best_s
        res = {"mode": mode, "shape": [m, n], "batch": B, "sweeps": f.sweeps, "status": f.status, "sigma_rel_err": rel,
               "recon_over_eckart_young": rec_ratio}
        if timing:
            ts = []
            for _ in range(2):
                t0 = time.perf_counter(); f = _lib.scaled_svd(Ws, Ss); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            res["ms"] = [round(t * 1e3, 1) for t in ts]
            _lib.profile_enable(True)
            before = _lib.profile_read()
            _lib.scaled_svd(Ws, Ss, max_sweeps=2)
            torch.cuda.synchronize()
            after = _lib.profile_read()
            _lib.profile_enable(False)
            res["class_ms_2_sweeps"] = {k: round(after[k][0] - before[k][0], 2) for k in after if after[k][1] > before[k][1]}
        del f
        print(json.dumps(res), flush=True)
        out[mode] = res
    return out

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        case(256, 256, 1, timing=False)
        case(512, 384, 2, timing=False)
        case(1024, 1024, 3, timing=False)
    elif which == "big":
        case(4096, 4096, int(os.environ.get("PROF_BATCH", "9")))
    elif which == "tri":
        case(4096, 4096, int(os.environ.get("PROF_BATCH", "27")), modes=("tri",))
