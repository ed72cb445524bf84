"""Design simulator #2: emulates the CUDA path's structure (batched pairs, inner two-sided Jacobi with
fixed-position round-robin, optional sort-swap, threshold skipping). CPU/torch. Not shipped."""
import torch, numpy as np, sys, time
torch.set_grad_enabled(False)

def rr_rounds(nb):
    idx = list(range(nb)); rounds = []
    for _ in range(nb - 1):
        rounds.append([(min(idx[i], idx[nb-1-i]), max(idx[i], idx[nb-1-i])) for i in range(nb // 2)])
        idx = [idx[0]] + [idx[-1]] + idx[1:-1]
    return rounds

def step_perm(k):
    """position permutation applied after every inner step: newpos_of[oldpos]"""
    h = k // 2; top = [2*i for i in range(h)]; bot = [2*i+1 for i in range(h)]
    src_top = [top[0], bot[0]] + top[1:h-1]           # new_top[i] comes from ...
    src_bot = bot[1:] + [top[h-1]]
    if h == 1: src_top = [top[0]]; src_bot = [bot[0]]
    src = [0]*k
    for i in range(h): src[2*i] = src_top[i]; src[2*i+1] = src_bot[i]
    return torch.tensor(src)   # new[pos] = old[src[pos]]

def inner_jacobi(G, max_sweeps=8, tol=1e-7, sort=True, stats=None):
    """G: [P,k,k] fp32 symmetric. returns R [P,k,k] with R^T G R ~ diag; columns ordered by label"""
    P, k, _ = G.shape; h = k // 2
    G = G.clone(); R = torch.eye(k, dtype=G.dtype).expand(P, k, k).clone()
    lab = torch.arange(k).expand(P, k).clone()
    src = step_perm(k)
    ev = torch.arange(0, k, 2); od = ev + 1
    nsw = 0
    for sw in range(max_sweeps):
        d = torch.diagonal(G, dim1=1, dim2=2)
        off = (G.abs() / torch.sqrt(d[:, :, None] * d[:, None, :]).clamp_min(1e-37))
        off = off - torch.diag_embed(torch.diagonal(off, dim1=1, dim2=2))
        if off.amax() < tol: break
        nsw += 1
        for st in range(k - 1):
            app = G[:, ev, ev]; aqq = G[:, od, od]; apq = G[:, ev, od]
            # symmetric Schur: small-angle rotation
            tau = (aqq - app) / (2 * apq)
            t = torch.sign(tau) / (tau.abs() + torch.sqrt(1 + tau * tau))
            t = torch.where(tau == 0, torch.ones_like(t), t)
            small = apq.abs() <= 1e-9 * torch.sqrt((app * aqq).abs())   # relative threshold
            t = torch.where(small | ~torch.isfinite(t), torch.zeros_like(t), t)
            c = 1 / torch.sqrt(1 + t * t); s = t * c
            if sort:
                npp = app - t * apq; nqq = aqq + t * apq
                lp = lab[:, ev]; lq = lab[:, od]
                swap = ((npp < nqq) & (lp < lq)) | ((npp > nqq) & (lp > lq))
                c2 = torch.where(swap, -s, c); s2 = torch.where(swap, c, s)   # extra 90deg: [c -s; s c]*[0 1;-1 0]... 
                c, s = c2, s2
            # J = [[c, s],[-s, c]] acting on columns (p,q): newp = c*p - s*q ; newq = s*p + c*q
            def rot_cols(M):
                Mp = M[:, :, ev]; Mq = M[:, :, od]
                return c[:, None, :] * Mp - s[:, None, :] * Mq, s[:, None, :] * Mp + c[:, None, :] * Mq
            Gp, Gq = rot_cols(G); G[:, :, ev] = Gp; G[:, :, od] = Gq
            G = G.transpose(1, 2).contiguous()
            Gp, Gq = rot_cols(G); G[:, :, ev] = Gp; G[:, :, od] = Gq
            Rp, Rq = rot_cols(R); R[:, :, ev] = Rp; R[:, :, od] = Rq
            # permute positions
            G = G[:, src][:, :, src]; R = R[:, :, src]; lab = lab[:, src]
    if stats is not None: stats.append(nsw)
    inv = torch.argsort(lab, dim=1)
    R = torch.gather(R, 2, inv[:, None, :].expand(P, k, k))
    # Newton-Schulz polish: restores orthogonality lost by fp32 rotation accumulation
    R = R @ (1.5 * torch.eye(k) - 0.5 * (R.transpose(1, 2) @ R))
    return R

def block_jacobi(X, b, max_sweeps=20, tol=2e-6, inner="jacobi", sort=True, inner_sweeps=8, verbose=True):
    X = X.clone().float(); nv, m = X.shape; nb = nv // b; k = 2 * b
    rounds = rr_rounds(nb); hist = []
    ar = torch.arange(b)
    # pre-orthogonalise inside blocks is implied by full-G solves
    for sw in range(max_sweeps):
        maxoff = 0.0; nupd = 0; isw = []
        for rnd in rounds:
            I = torch.tensor([p[0] for p in rnd]); J = torch.tensor([p[1] for p in rnd])
            rows = torch.cat([I[:, None] * b + ar, J[:, None] * b + ar], 1)
            Pn = X[rows]
            G = Pn @ Pn.transpose(1, 2)
            d = torch.diagonal(G, dim1=1, dim2=2).clamp_min(1e-37).sqrt()
            C = G.abs() / (d[:, :, None] * d[:, None, :]); C = C - torch.diag_embed(torch.diagonal(C, dim1=1, dim2=2))
            po = C.amax(dim=(1, 2)); maxoff = max(maxoff, po.max().item())
            act = po >= tol
            if act.any():
                Ga = G[act]
                if inner == "eigh":
                    w, R = torch.linalg.eigh(Ga.double()); R = R.flip(2).float()
                else:
                    R = inner_jacobi(Ga, max_sweeps=inner_sweeps, sort=sort, stats=isw)
                X[rows[act]] = R.transpose(1, 2) @ Pn[act]
                nupd += int(act.sum())
        hist.append(maxoff)
        if verbose: print(f"  sweep {sw}: maxoff {maxoff:.3e} updates {nupd}/{len(rounds)*len(rounds[0])} inner sweeps avg {np.mean(isw) if isw else 0:.2f}", flush=True)
        if maxoff < tol: break
    return X, hist

def make(kind, n, m, seed=233):
    g = torch.Generator().manual_seed(seed)
    if kind == "gauss":
        W = (torch.randn(m, n, generator=g) * 0.02).half().float()
    else:
        U, _ = torch.linalg.qr(torch.randn(m, n, generator=g)); V, _ = torch.linalg.qr(torch.randn(n, n, generator=g))
        sv = torch.arange(1, n + 1).float() ** -1.0
        W = (U * sv) @ V.T; W = (W / W.abs().max()).half().float()
    s = torch.exp(torch.randn(n, generator=g))
    return W * (s ** 0.5 + 1e-6)

if __name__ == "__main__":
    n = int(sys.argv[1]); b = int(sys.argv[2]); inner = sys.argv[3]; sort = sys.argv[4] == "sort"
    isw = int(sys.argv[5]) if len(sys.argv) > 5 else 8
    kinds = sys.argv[6].split(",") if len(sys.argv) > 6 else ("gauss", "power")
    for kind in kinds:
        Ws = make(kind, n, n)
        sv64 = torch.linalg.svdvals(Ws.double())
        r = int(n * n * 0.9) // (2 * n)
        t = time.time(); print(kind, n, b, inner, sort, isw)
        X, hist = block_jacobi(Ws.T.contiguous(), b, inner=inner, sort=sort, inner_sweeps=isw)
        sj = torch.linalg.norm(X.double(), dim=1).sort(descending=True).values
        print(f"  rel err kept {((sj[:r]/sv64[:r]-1).abs().max()):.3e} all {((sj/sv64-1).abs().max()):.3e} ({time.time()-t:.0f}s)")
