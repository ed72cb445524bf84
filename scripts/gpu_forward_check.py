"""Forward kernel check + timing at BASELINE config 4 (diagnostics)."""
import sys, os, time
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
    print(f"forward M={M} n={n} r={r} m={m} {dtype}: max abs err {err:.3e} (|y|max {ref.abs().max().item():.2f})", flush=True)
for shp in [(128, 64, 64, 128), (256, 512, 128, 256), (300, 512, 100 * 8 // 8, 384), (1000, 1024, 256, 1000), (4096, 4096, 512, 4096)]:
    check(*shp)
check(512, 1024, 256, 1024, dtype=torch.bfloat16)
# timing, config 4
M, n, m = 32 * 2048, 4096, 4096
x = (torch.randn(M, n, device=dev) * 0.125).half()
for r in (256, 512, 1024):
    B = (torch.randn(r, n, device=dev) / n ** 0.5).half(); A = (torch.randn(m, r, device=dev) / r ** 0.5).half()
    for name, fn in (("ours", lambda: _lib.lowrank_forward(x, A, B, None)),
                     ("cublas(ref module)", lambda: torch.nn.functional.linear(torch.nn.functional.linear(x, B), A))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * M * r * (n + m)
        print(f"r={r} {name}: {ms*1e3:.0f} us  {fl/ms/1e9:.0f} TFLOP/s", flush=True)
