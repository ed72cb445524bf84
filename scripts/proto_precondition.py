"""Numerical experiment (CPU, not shipped): does a QR pre-conditioner (Drmac-Veselic) cut the number of block-Jacobi
sweeps on the SQUARE problems, where the Gram pre-conditioner of the rectangles has nothing to reduce?
Uses the design simulator of proto_jacobi2.py (one inner cyclic sweep per visit, as the CUDA path does).

    python scripts/proto_precondition.py 1024 64 gauss,power
"""
import sys, time
import numpy as np, scipy.linalg, torch
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from proto_jacobi2 import block_jacobi, make

n = int(sys.argv[1]); b = int(sys.argv[2]); kinds = sys.argv[3].split(",")
isw = int(sys.argv[4]) if len(sys.argv) > 4 else 1
for kind in kinds:
    Ws = make(kind, n, n)
    sv64 = torch.linalg.svdvals(Ws.double())
    r = int(n * n * 0.9) // (2 * n)
    Wd = Ws.double().numpy()
    Q, R = np.linalg.qr(Wd)
    Qp, Rp, piv = scipy.linalg.qr(Wd, pivoting=True)
    # rows sorted by norm descending (cheap stand-in for pivoting)
    order = np.argsort(-np.linalg.norm(Wd, axis=0))
    Qs, Rs = np.linalg.qr(Wd[:, order])
    variants = {
        "direct: columns of W": Ws.T.contiguous(),
        "QR: rows of R": torch.from_numpy(R).float(),
        "QR: columns of R": torch.from_numpy(R.T.copy()).float(),
        "pivoted QR: rows of R": torch.from_numpy(Rp).float(),
        "pivoted QR: columns of R": torch.from_numpy(Rp.T.copy()).float(),
        "norm-sorted QR: rows of R": torch.from_numpy(Rs).float(),
    }
    for name, X0 in variants.items():
        t = time.time()
        X, hist = block_jacobi(X0, b, inner="jacobi", sort=True, inner_sweeps=isw, verbose=False, tol=2e-5)
        sj = torch.linalg.norm(X.double(), dim=1).sort(descending=True).values
        err = ((sj[:r] / sv64[:r] - 1).abs().max()).item()
        print(f"{kind:6s} n={n} b={b} {name:28s} sweeps {len(hist):2d}  kept-sigma rel err {err:.2e}  trace "
              + " ".join(f"{h:.0e}" for h in hist) + f"  ({time.time() - t:.0f}s)", flush=True)
