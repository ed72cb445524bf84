"""Probe: do the read-bound x B^T and the write-bound t A^T of the forward overlap when they run side by side on half
of the SMs each?  Two half-size forwards (32 768 tokens each), (a) one after the other on the whole GPU, (b) on two
streams with 37 clusters per kernel, the second stream delayed by one GEMM so that A's second product meets B's first."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib
dev = "cuda"
n = m = 4096; Mh = 32768
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for r in (128, 256, 512):
    g = torch.Generator(device=dev).manual_seed(r)
    xa = (torch.randn(Mh, n, device=dev, generator=g) * 0.125).half(); xb = (torch.randn(Mh, n, device=dev, generator=g) * 0.125).half()
    B = (torch.randn(r, n, device=dev, generator=g) / n ** 0.5).half(); A = (torch.randn(m, r, device=dev, generator=g) / r ** 0.5).half()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    def seq():
        os.environ.pop("ASVD_B200_FWD_MAXCL", None)
        _lib.lowrank_forward(xa, A, B, None); _lib.lowrank_forward(xb, A, B, None)
    def pipe(delay_cycles):
        os.environ["ASVD_B200_FWD_MAXCL"] = "37"
        cur = torch.cuda.current_stream()
        sa.wait_stream(cur); sb.wait_stream(cur)
        with torch.cuda.stream(sa): _lib.lowrank_forward(xa, A, B, None)
        with torch.cuda.stream(sb):
            torch.cuda._sleep(delay_cycles); _lib.lowrank_forward(xb, A, B, None)
        cur.wait_stream(sa); cur.wait_stream(sb)
    def timeit(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize(); ts = []
        for _ in range(7):
            flush.zero_(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        return sorted(ts)[3]
    out = {"r": r, "sequential_full_grid_us": round(timeit(seq), 1)}
    for d_us in (0, 40, 80, 120):
        out[f"two_streams_37cl_delay{d_us}us"] = round(timeit(lambda: pipe(int(d_us * 1900))), 1)
    os.environ.pop("ASVD_B200_FWD_MAXCL", None)
    print(json.dumps(out), flush=True)
