"""Small forwards through all three tensor-core paths (CTA-pair GEMMs, fused kernel, single-CTA GEMMs) for
compute-sanitizer memcheck: ragged M / r / m, ranks that are not multiples of 8 (padded pitch + padded copy of A),
bias, bf16, the split tail tiles of the fused schedule."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
for mode in ("pair", "fused", "1cta"):
    os.environ["ASVD_B200_FWD"] = mode
    for (M, n, r, m, dt) in [(300, 512, 100, 384, torch.float16), (1, 64, 13, 32, torch.float16), (700, 256, 345, 520, torch.bfloat16),
                             (513, 1024, 256, 1000, torch.float16), (40000, 128, 64, 136, torch.float16)]:
        x = (torch.randn(M, n, device=dev, generator=g) * 0.125).to(dt)
        B = (torch.randn(r, n, device=dev, generator=g) / n ** 0.5).to(dt)
        A = (torch.randn(m, r, device=dev, generator=g) / r ** 0.5).to(dt)
        bias = torch.randn(m, device=dev, generator=g).to(dt)
        y = _lib.lowrank_forward(x, A, B, bias)
        ref = torch.nn.functional.linear(torch.nn.functional.linear(x, B), A, bias)
        err = (y.float() - ref.float()).abs().max().item()
        assert err < 3e-2 * max(1.0, ref.float().abs().max().item()), (mode, M, n, r, m, err)
torch.cuda.synchronize()
print("done")
