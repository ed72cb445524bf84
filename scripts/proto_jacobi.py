"""Numerics prototype (CPU, numpy) of the block one-sided Jacobi used by the CUDA path.
Design simulator only: decides block width, sweep counts, tolerances. Not shipped, not an oracle."""
import numpy as np, sys, time

def rr_rounds(nb):
    """round-robin tournament: nb even -> nb-1 rounds of nb/2 disjoint pairs"""
    idx = list(range(nb)); rounds = []
    for _ in range(nb - 1):
        rounds.append([(min(idx[i], idx[nb-1-i]), max(idx[i], idx[nb-1-i])) for i in range(nb // 2)])
        idx = [idx[0]] + [idx[-1]] + idx[1:-1]
    return rounds

def block_jacobi(X, b, max_sweeps=15, tol=1e-6, inner="eigh", gram_dtype=np.float32, verbose=True):
    X = X.astype(np.float32).copy()
    nv, m = X.shape; nb = nv // b
    rounds = rr_rounds(nb)
    hist = []
    for sw in range(max_sweeps):
        maxoff = 0.0
        for rnd in rounds:
            I = np.array([p[0] for p in rnd]); J = np.array([p[1] for p in rnd])
            rows = (np.concatenate([I[:, None] * b + np.arange(b), J[:, None] * b + np.arange(b)], 1))  # P x 2b
            P = X[rows]                      # P x 2b x m
            Pg = P.astype(gram_dtype)
            G = np.einsum('pim,pjm->pij', Pg, Pg, optimize=True).astype(np.float32)
            d = np.sqrt(np.maximum(np.einsum('pii->pi', G), 1e-30))
            C = np.abs(G) / (d[:, :, None] * d[:, None, :])
            C[:, np.arange(2*b), np.arange(2*b)] = 0
            maxoff = max(maxoff, C.max())
            if inner == "eigh":
                w, R = np.linalg.eigh(G.astype(np.float64))
                R = R[:, :, ::-1].astype(np.float32)      # descending
            X[rows] = np.einsum('pij,pim->pjm', R, P, optimize=True)   # new_j = sum_i R[i,j] old_i
        hist.append(maxoff)
        if verbose: print(f"  sweep {sw}: max off-cos at visit {maxoff:.3e}", flush=True)
        if maxoff < tol: break
    return X, hist

def make(kind, n, m, seed=233):
    rng = np.random.default_rng(seed)
    if kind == "gauss":
        W = (rng.standard_normal((m, n)) * 0.02).astype(np.float16).astype(np.float32)
    else:
        U, _ = np.linalg.qr(rng.standard_normal((m, n))); V, _ = np.linalg.qr(rng.standard_normal((n, n)))
        sv = (np.arange(1, n + 1) ** -1.0)
        W = ((U * sv) @ V.T).astype(np.float32)
        W = (W / np.abs(W).max()).astype(np.float16).astype(np.float32)
    s = np.exp(rng.standard_normal(n)).astype(np.float32)
    return W * (s ** 0.5 + 1e-6)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    for kind in ("gauss", "power"):
        Ws = make(kind, n, n)
        sv64 = np.linalg.svd(Ws.astype(np.float64), compute_uv=False)
        sv32 = np.linalg.svd(Ws, compute_uv=False)
        r = int(n * n * 0.9) // (2 * n)
        print(kind, "kappa_top_r", sv64[0] / sv64[r-1], "fp32 lapack rel err kept", np.abs(sv32[:r]/sv64[:r]-1).max())
        t = time.time()
        X, hist = block_jacobi(Ws.T, b)
        sj = np.sort(np.linalg.norm(X.astype(np.float64), axis=1))[::-1]
        print(f"  jacobi rel err kept {np.abs(sj[:r]/sv64[:r]-1).max():.3e} all {np.abs(sj/sv64-1).max():.3e}  ({time.time()-t:.1f}s)")
