"""A/B wall-clock of ASVD_B200_OVERLAP (two half-batches on two streams) against the one-stream schedule, in one
process, on the bench shape and the Llama-2-7B rectangles (diagnostic; prints one JSON line per case)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
cases = [(4096, 4096, 4), (4096, 4096, 2), (11008, 4096, 4), (4096, 11008, 4), (4096, 4096, 3)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for m, n, B in cases:
    g = torch.Generator(device=dev).manual_seed(233)
    Ws = [(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for _ in range(B)]
    Ss = [_lib.scaling_vector(torch.exp(torch.randn(n, device=dev, generator=g)).half(), None, 0.5, n, dev) for _ in range(B)]
    out = {"shape": [m, n], "batch": B}
    sig = {}
    for ov in ("0", "1", "0", "1"):
        os.environ["ASVD_B200_OVERLAP"] = ov
        f = _lib.scaled_svd(Ws, Ss)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            f = _lib.scaled_svd(Ws, Ss)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        out.setdefault("ms_overlap" + ov, []).extend(round(t * 1e3, 1) for t in ts)
        out["sweeps" + ov] = f.sweeps
        s = torch.stack([f.sigma(b) for b in range(B)])
        if ov in sig:
            out["bitwise_repeat" + ov] = bool(torch.equal(sig[ov], s))
        sig[ov] = s
    out["sigma_bitwise_equal"] = bool(torch.equal(sig["0"], sig["1"]))
    print(json.dumps(out), flush=True)
