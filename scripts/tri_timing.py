"""Where does the time of the triangular solve's two kernels go?  Needs the timing build (nvcc ... -DASVD_SOLVE_TIMING=1 -o
asvd4llm_b200/csrc/libasvd_b200_timing.so asvd4llm_b200/csrc/*.cu): thread 0 of CTA (0,0) accumulates clock64 deltas between
marks; printed for the last launch of a two-sweep run."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import build as B
B.LIB = os.path.join(B.CSRC, "libasvd_b200_timing.so"); B.stale = lambda: False
from asvd4llm_b200 import _lib
lib = _lib.load()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(233)
Bn = int(os.environ.get("PROF_BATCH", "27"))
Ws = [(torch.randn(4096, 4096, device=dev, generator=g) * 0.02).half() for _ in range(Bn)]
Ss = [_lib.scaling_vector(torch.exp(torch.randn(4096, device=dev, generator=g)).half(), None, 0.5, 4096, dev) for _ in range(Bn)]
_lib.scaled_svd(Ws, Ss, max_sweeps=2, allow_status=(0, 5))
torch.cuda.synchronize()
out = (C.c_ulonglong * 16)()
assert lib.asvd_debug_tri_timing(out) == 0
names = ["G: prologue", "G: first three steps", "G: four steps of a round (x31)", "G: fold", "G: staging writes", "G: wait at the move barrier",
         "G: staging reads", "G: tail", "R: prologue", "R: first step + wait for the record", "R: four steps (x31)", "R: fold + move (x30)",
         "R: (loop exit)", "R: wait for the other warps", "R: sort, norms, store", "-"]
for half, lo in (("G kernel", 0), ("replay kernel", 8)):
    tot = sum(out[lo:lo + 8])
    print(f"{half}: {tot} clocks (batch {Bn}, CTA (0,0), thread 0, last launch)")
    for n, v in zip(names[lo:lo + 8], out[lo:lo + 8]):
        print(f"  {n:44s} {v:8d} clocks  {100.0 * v / max(tot, 1):5.1f} %")
