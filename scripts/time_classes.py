"""Per-class CUDA-event times of the first two sweeps (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(233)
B = int(os.environ.get("PROF_BATCH", "4"))
Ws = [(torch.randn(4096, 4096, device=dev, generator=g) * 0.02).half() for _ in range(B)]
Ss = [_lib.scaling_vector(torch.exp(torch.randn(4096, device=dev, generator=g)).half(), None, 0.5, 4096, dev) for _ in range(B)]
_lib.scaled_svd(Ws, Ss, max_sweeps=1)
_lib.profile_enable(True)
before = _lib.profile_read()
_lib.scaled_svd(Ws, Ss, max_sweeps=2)
torch.cuda.synchronize()
after = _lib.profile_read()
_lib.profile_enable(False)
print(os.environ.get("ASVD_B200_DBG_STEPS"), {k: (round(after[k][0], 2), after[k][1] - before[k][1]) for k in after if after[k][1] > before[k][1]})
