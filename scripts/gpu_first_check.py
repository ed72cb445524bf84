"""Verbose first-contact check of every kernel on a real GPU (diagnostics, not a test)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
from oracle import asvd_oracle as O

torch.manual_seed(0)
dev = "cuda"

def check_svd(m, n, kind="gauss", dtype=torch.float16, ratio=0.9, batch=1, fuse="UV"):
    Ws, Ss = [], []
    for b in range(batch):
        W, s = O.synthetic_weight(m, n, seed=233 + b, kind=kind, dtype=dtype)
        Ws.append(W); Ss.append((s ** 0.5 + 1e-6).float())
    torch.cuda.synchronize(); t = time.perf_counter()
    f = _lib.scaled_svd([w.to(dev) for w in Ws], [s.to(dev) for s in Ss])
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    r = min(O.rank_for_ratio(m, n, ratio), min(m, n))
    for b in range(batch):
        sig = f.sigma(b).cpu()
        ref = torch.linalg.svdvals(Ws[b].double() * Ss[b].double())
        A, B = f.extract(r, fuse, torch.float32, b)
        A, B = A.cpu().double(), B.cpu().double()
        Wd = Ws[b].double(); sd = Ss[b].double()
        U, S, Vh = torch.linalg.svd(Wd * sd, full_matrices=False)
        Wtr = (U[:, :r] * S[:r]) @ Vh[:r] / sd
        rec = ((A @ B - Wtr) * sd).norm() / (Wd * sd).norm()
        print(f"svd {m}x{n} {kind} {dtype} b={b}/{batch}: status={f.status} sweeps={f.sweeps[b]} time={dt*1e3:.1f}ms "
              f"sigma rel err kept={((sig[:r].double()-ref[:r]).abs()/ref[:r]).max():.2e} all={((sig.double()-ref).abs()/ref).max():.2e} "
              f"recon-vs-exact-trunc={rec:.2e}", flush=True)

for (m, n) in [(96, 64), (64, 96), (128, 128), (200, 80), (80, 200), (256, 384), (512, 512)]:
    check_svd(m, n, dtype=torch.float32)
check_svd(1024, 1024); check_svd(1024, 1024, kind="power"); check_svd(1024, 1024, batch=3)
check_svd(2048, 1024); check_svd(1024, 2048)
check_svd(4096, 4096); check_svd(4096, 4096)
check_svd(4096, 4096, batch=4)

# forward
x = (torch.randn(4, 300, 512) * 0.125).half().to(dev); A = (torch.randn(384, 100) / 10).half().to(dev)
B = (torch.randn(100, 512) / 22).half().to(dev); bias = torch.randn(384).half().to(dev)
y = _lib.lowrank_forward(x, A, B, bias)
ref = O.lowrank_forward(x.cpu(), A.cpu(), B.cpu(), bias.cpu(), compute_dtype=torch.float64)
print("forward max abs err vs fp64:", (y.cpu().double() - ref).abs().max().item(), "ymax", ref.abs().max().item())
# absstat
xx = torch.randn(1, 2048, 4096).half().to(dev); acc = torch.zeros(4096, dtype=torch.half, device=dev)
_lib.absstat_accum(xx, acc, "abs_mean"); _lib.absstat_accum(xx, acc, "abs_mean")
want = xx.abs().mean(dim=-2).view(-1); want = want + want
print("absstat mean max rel err:", ((acc.float() - want.float()).abs() / want.float()).max().item())
acc2 = torch.zeros(4096, dtype=torch.half, device=dev); _lib.absstat_accum(xx, acc2, "abs_max")
print("absstat max equal:", torch.equal(acc2, xx.abs().amax(dim=-2).view(-1)))
sdm = (torch.rand(4096) * 3).half(); sdm[::7] = 0
sv = _lib.scaling_vector(sdm.to(dev), None, 0.5, 4096, dev).cpu()
want = O.scaling_vector(sdm, None, 0.5).float()
print("scaling vector max abs diff:", (sv - want).abs().max().item())
