"""Where does a step of solve_quad_kernel go?  Needs the timing build (nvcc ... -DASVD_SOLVE_TIMING=1 -o
asvd4llm_b200/csrc/libasvd_b200_timing.so asvd4llm_b200/csrc/*.cu): lane 0 of the lead warp of CTA (0,0) accumulates clock64
deltas between marks; printed per step / per round of the last solve launch of a two-sweep run."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import build as B
B.LIB = os.path.join(B.CSRC, "libasvd_b200_timing.so"); B.stale = lambda: False
from asvd4llm_b200 import _lib
lib = _lib.load()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(233)
Bn = int(os.environ.get("PROF_BATCH", "18"))
Ws = [(torch.randn(4096, 4096, device=dev, generator=g) * 0.02).half() for _ in range(Bn)]
Ss = [_lib.scaling_vector(torch.exp(torch.randn(4096, device=dev, generator=g)).half(), None, 0.5, 4096, dev) for _ in range(Bn)]
_lib.scaled_svd(Ws, Ss, max_sweeps=2, allow_status=(0, 5))
torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
assert lib.asvd_debug_solve_timing(out) == 0
names = ["since previous apply (fold / move remainder)", "gather shuffles", "rotation parameters", "shuffle back + publish + arrive",
         "parameter loads (LDS)", "own 128 FMAs (rows + cols)", "fold", "quad move incl. barrier waits"]
tot = sum(out)
print(f"lead lane of CTA (0,0), last launch: {tot} cycles over 127 steps = {tot / 127:.0f} per step")
for n, v in zip(names, out):
    print(f"  {n:48s} {v:8d} cycles  {100.0 * v / tot:5.1f} %   {v / 127:6.0f} per step")
