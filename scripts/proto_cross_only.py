"""Design experiment (CPU, not shipped): inner sweep restricted to the CROSS pivots of a block pair.
A full cyclic sweep of the 128x128 pair problem has 127 steps; 63 of them (4032 of 8128 pivots) rotate vectors of the
SAME block against each other, and every block meets that work again in each of the 63 rounds of an outer sweep.  Variant:
round 0 of every outer sweep runs the full inner sweep (every block is in exactly one pair of round 0, so intra-block
pivots are visited once per outer sweep), all other rounds only the 64 bipartite steps (i in I) x (j in J).
Counts outer sweeps to convergence for both schedules.

    python scripts/proto_cross_only.py 1024 64 gauss,power
"""
import sys, time
import torch
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from proto_jacobi2 import rr_rounds, make
torch.set_grad_enabled(False)


def full_steps(k):
    idx = list(range(k)); steps = []
    for _ in range(k - 1):
        steps.append([(min(idx[i], idx[k - 1 - i]), max(idx[i], idx[k - 1 - i])) for i in range(k // 2)])
        idx = [idx[0]] + [idx[-1]] + idx[1:-1]
    return steps


def cross_steps(b):
    return [[(i, b + (i + s) % b) for i in range(b)] for s in range(b)]


def inner(G, steps, thresh=1e-8):
    P, k, _ = G.shape
    G = G.clone(); R = torch.eye(k).expand(P, k, k).clone()
    for st in steps:
        p = torch.tensor([a for a, _ in st]); q = torch.tensor([c for _, c in st])
        app = G[:, p, p]; aqq = G[:, q, q]; apq = G[:, p, q]
        tau = (aqq - app) / (2 * apq)
        t = torch.sign(tau) / (tau.abs() + torch.sqrt(1 + tau * tau))
        t = torch.where(tau == 0, torch.ones_like(t), t)
        small = apq.abs() <= thresh * torch.sqrt((app * aqq).abs())
        t = torch.where(small | ~torch.isfinite(t), torch.zeros_like(t), t)
        c = 1 / torch.sqrt(1 + t * t); s = t * c

        def rot(M):
            Mp = M[:, :, p]; Mq = M[:, :, q]
            M[:, :, p] = c[:, None, :] * Mp - s[:, None, :] * Mq
            M[:, :, q] = s[:, None, :] * Mp + c[:, None, :] * Mq
            return M
        G = rot(G); G = G.transpose(1, 2).contiguous(); G = rot(G); R = rot(R)
    # sort columns by new diagonal, descending (de Rijk)
    order = torch.argsort(torch.diagonal(G, dim1=1, dim2=2), dim=1, descending=True)
    R = torch.gather(R, 2, order[:, None, :].expand(P, k, k))
    return R


def block_jacobi(X, b, mode, max_sweeps=30, tol=2e-5, skip=4e-6, sort_cross=True):
    X = X.clone().float(); nv, m = X.shape; nb = nv // b
    rounds = rr_rounds(nb); ar = torch.arange(b)
    fs, cs = full_steps(2 * b), cross_steps(b)
    hist, nsteps = [], 0
    for sw in range(max_sweeps):
        maxoff = 0.0
        for ri, rnd in enumerate(rounds):
            I = torch.tensor([p[0] for p in rnd]); J = torch.tensor([p[1] for p in rnd])
            rows = torch.cat([I[:, None] * b + ar, J[:, None] * b + ar], 1)
            Pn = X[rows]
            G = Pn @ Pn.transpose(1, 2)
            d = torch.diagonal(G, dim1=1, dim2=2).clamp_min(1e-37).sqrt()
            C = G.abs() / (d[:, :, None] * d[:, None, :]); C = C - torch.diag_embed(torch.diagonal(C, dim1=1, dim2=2))
            po = C.amax(dim=(1, 2)); maxoff = max(maxoff, po.max().item())
            act = po >= skip
            if act.any():
                full = mode == "full" or (mode == "cross" and ri == 0) or (mode == "cross2" and ri == 0 and sw % 2 == 0)
                steps = fs if full else cs
                nsteps += len(steps)
                R = inner(G[act], steps)
                X[rows[act]] = R.transpose(1, 2) @ Pn[act]
        hist.append(maxoff)
        if maxoff < tol:
            break
    return X, hist, nsteps


if __name__ == "__main__":
    n = int(sys.argv[1]); b = int(sys.argv[2]); kinds = sys.argv[3].split(",")
    modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["full", "cross"]
    for kind in kinds:
        Ws = make(kind, n, n)
        sv64 = torch.linalg.svdvals(Ws.double())
        r = int(n * n * 0.9) // (2 * n)
        # ascending-norm initial order, as the CUDA path
        X0 = Ws.T.contiguous(); X0 = X0[torch.argsort(X0.norm(dim=1))]
        for mode in modes:
            t = time.time()
            X, hist, nsteps = block_jacobi(X0, b, mode)
            sj = torch.linalg.norm(X.double(), dim=1).sort(descending=True).values
            err = ((sj[:r] / sv64[:r] - 1).abs().max()).item()
            print(f"{kind:6s} n={n} b={b} {mode:6s} sweeps {len(hist):2d} inner steps {nsteps:6d} kept-sigma rel err {err:.2e} trace "
                  + " ".join(f"{h:.0e}" for h in hist) + f" ({time.time() - t:.0f}s)", flush=True)
