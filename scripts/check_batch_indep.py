"""Is a weight's factorisation bitwise independent of its batch-mates, and run-to-run reproducible? (diagnostic)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
from oracle import asvd_oracle as O
m, n, batch = 2048, 2048, 4
Ws, Ss = [], []
for b in range(batch):
    W, s = O.synthetic_weight(m, n, seed=80 + b)
    Ws.append(W.cuda()); Ss.append((s ** 0.5 + 1e-6).float().cuda())
for mode in ("quad", "tri"):
    os.environ["ASVD_B200_SOLVE"] = mode
    for ms in (1, 2, 3, 0):
        full = [_lib.scaled_svd(Ws, Ss, max_sweeps=ms, allow_status=(0, 5)) for _ in range(2)]
        alone = [_lib.scaled_svd(Ws[-1:], Ss[-1:], max_sweeps=ms, allow_status=(0, 5)) for _ in range(2)]
        s = [f.sigma(batch - 1) for f in full] + [f.sigma(0) for f in alone]
        print(mode, "max_sweeps", ms, "batch run-to-run", torch.equal(s[0], s[1]), "alone run-to-run", torch.equal(s[2], s[3]),
              "batch vs alone", torch.equal(s[0], s[2]), "sweeps", full[0].sweeps, alone[0].sweeps, flush=True)
