import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib as L
from oracle import asvd_oracle as O
m = n = 2048
Ws, Ss = [], []
for b in range(4):
    W, s = O.synthetic_weight(m, n, seed=90 + b)
    Ws.append(W.cuda()); Ss.append((s ** 0.5 + 1e-6).float().cuda())
os.environ["ASVD_B200_GRAM_CHUNKS"] = "2"
batch = L.scaled_svd(Ws, Ss)
for i in range(4):
    alone = L.scaled_svd(Ws[i:i+1], Ss[i:i+1])
    d = (batch.sigma(i) - alone.sigma(0)).abs().max().item()
    print(i, "equal", torch.equal(batch.sigma(i), alone.sigma(0)), d, batch.sweeps, alone.sweeps, flush=True)
for ms in (1, 2, 4, 6, 8, 10):
    b2 = L.scaled_svd(Ws, Ss, max_sweeps=ms, allow_status=(0, 5)); a2 = L.scaled_svd(Ws[2:3], Ss[2:3], max_sweeps=ms, allow_status=(0, 5))
    print("max_sweeps", ms, torch.equal(b2.sigma(2), a2.sigma(0)), flush=True)
