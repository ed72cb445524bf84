"""Same-box GPU baselines for BASELINE config 2: upstream's own call, torch.svd_lowrank on CUDA
(modules/svd_linear.py:65, q = rank, niter 2), and torch.linalg.svd on CUDA (north_star's oracle), one 4096x4096 weight."""
import sys, os, json, time
import torch
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(233)
n = 4096
W = (torch.randn(n, n, device=dev, generator=g) * 0.02).half()
s = torch.exp(torch.randn(n, device=dev, generator=g)).half().float() ** 0.5 + 1e-6
Ws = W.float() * s
r = int(n * n * 0.9 / (2 * n))
out = {}
for name, fn in (("torch.svd_lowrank(cuda, q=%d, niter=2)" % r, lambda: torch.svd_lowrank(Ws, q=r)),
                 ("torch.linalg.svd(cuda, fp32, full_matrices=False)", lambda: torch.linalg.svd(Ws, full_matrices=False)),
                 ("torch.linalg.svd(cuda, fp32, driver=gesvdj)", lambda: torch.linalg.svd(Ws, full_matrices=False, driver="gesvdj"))):
    try:
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(2):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(json.dumps({"baseline": name, "s_per_matrix": round(min(ts), 4), "matrices_per_s": round(1 / min(ts), 3)}), flush=True)
    except Exception as e:
        print(json.dumps({"baseline": name, "error": str(e)[:200]}), flush=True)
