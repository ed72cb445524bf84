"""Design experiment (CPU, not shipped): classical threshold Jacobi at block-pair granularity.  A pair whose largest
cosine at visit time is below tau is left alone (its Gram is still computed); tau = clamp(c * max cosine of the previous
sweep, tol_skip, tau_max), so that while large rotations are still going on elsewhere the many nearly orthogonal pairs are
not re-solved and re-streamed.  Counts sweeps and pair updates (solve + update work) per schedule.

    python scripts/proto_threshold.py 1024 64 gauss,power
"""
import sys, time
import torch
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from proto_jacobi2 import rr_rounds, make
from proto_cross_only import full_steps, inner
torch.set_grad_enabled(False)


def block_jacobi(X, b, c, tau_max, max_sweeps=40, tol=2e-5, skip=4e-6):
    X = X.clone().float(); nv, m = X.shape; nb = nv // b
    rounds = rr_rounds(nb); ar = torch.arange(b)
    fs = full_steps(2 * b)
    hist, updates, taus = [], [], []
    prev = 1.0
    for sw in range(max_sweeps):
        tau = min(max(c * prev, skip), tau_max) if c > 0 else skip
        maxoff, nupd = 0.0, 0
        for rnd in rounds:
            I = torch.tensor([p[0] for p in rnd]); J = torch.tensor([p[1] for p in rnd])
            rows = torch.cat([I[:, None] * b + ar, J[:, None] * b + ar], 1)
            Pn = X[rows]
            G = Pn @ Pn.transpose(1, 2)
            d = torch.diagonal(G, dim1=1, dim2=2).clamp_min(1e-37).sqrt()
            C = G.abs() / (d[:, :, None] * d[:, None, :]); C = C - torch.diag_embed(torch.diagonal(C, dim1=1, dim2=2))
            po = C.amax(dim=(1, 2)); maxoff = max(maxoff, po.max().item())
            act = po >= tau
            if act.any():
                nupd += int(act.sum())
                R = inner(G[act], fs)
                X[rows[act]] = R.transpose(1, 2) @ Pn[act]
        hist.append(maxoff); updates.append(nupd); taus.append(tau)
        prev = maxoff
        if maxoff < tol:
            break
    return X, hist, updates, taus


if __name__ == "__main__":
    n = int(sys.argv[1]); b = int(sys.argv[2]); kinds = sys.argv[3].split(",")
    scheds = [(0.0, 0.0), (0.02, 1e-2), (0.05, 1e-2), (0.1, 1e-2), (0.1, 3e-2), (0.2, 5e-2)]
    for kind in kinds:
        Ws = make(kind, n, n)
        sv64 = torch.linalg.svdvals(Ws.double())
        r = int(n * n * 0.9) // (2 * n)
        X0 = Ws.T.contiguous(); X0 = X0[torch.argsort(X0.norm(dim=1))]
        nb = n // b; per_sweep = (nb - 1) * (nb // 2)
        for c, tmax in scheds:
            t = time.time()
            X, hist, upd, taus = block_jacobi(X0, b, c, tmax)
            sj = torch.linalg.norm(X.double(), dim=1).sort(descending=True).values
            err = ((sj[:r] / sv64[:r] - 1).abs().max()).item()
            # cost model from the measured round at 18 weights: gram 217, solve 468, update 409 (us) -> gram always, rest per update
            cost = sum(217 * per_sweep + (468 + 409) * u for u in upd) / (1094.0 * per_sweep)
            print(f"{kind:6s} n={n} c={c:4.2f} tau_max={tmax:5.0e} sweeps {len(hist):2d} updates {sum(upd):5d} ({sum(upd) / per_sweep:5.2f} sweeps' worth) "
                  f"cost {cost:5.2f} full-sweep equivalents  kept-sigma err {err:.1e}  upd/sweep " + " ".join(str(u) for u in upd) + f" ({time.time() - t:.0f}s)", flush=True)
