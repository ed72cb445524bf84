"""Experiment: does ordering the vectors by norm before the first sweep change the sweep count? (diagnostic)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
m = int(os.environ.get("PROF_M", "4096")); n = int(os.environ.get("PROF_N", "4096")); B = 4
g = torch.Generator(device=dev).manual_seed(233)
Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
def run(tag, perms):
    W2 = [W[:, p].contiguous() if m >= n else W[p].contiguous() for W, p in zip(Ws, perms)]
    S2 = [s[p].contiguous() if m >= n else s for s, p in zip(Ss, perms)]
    _lib.scaled_svd(W2, S2); torch.cuda.synchronize()
    t0 = time.perf_counter(); f = _lib.scaled_svd(W2, S2); torch.cuda.synchronize()
    print(tag, f"{(time.perf_counter()-t0)*1e3:.1f} ms", f.sweeps, flush=True)
nv = min(m, n)
ident = [torch.arange(nv, device=dev) for _ in range(B)]
run("as given   ", ident)
norms = [((W.float() * s).norm(dim=0) if m >= n else (W.float() * s).norm(dim=1)) for W, s in zip(Ws, Ss)]
run("descending ", [torch.argsort(x, descending=True) for x in norms])
run("ascending  ", [torch.argsort(x) for x in norms])
# interleaved: block k gets every (nv/64)-th vector of the sorted order (every block spans the whole norm range)
il = []
for x in norms:
    o = torch.argsort(x, descending=True)
    il.append(o.view(64, nv // 64).t().reshape(-1) if nv % 64 == 0 else o)
run("interleaved", il)
