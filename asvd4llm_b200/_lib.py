"""ctypes binding of include/asvd_b200.h.  PyTorch is used for device memory and streams only.

There is no CPU fallback: every entry point raises if the CUDA library cannot be loaded or no CUDA device is
present (the product path must fail loudly rather than run somewhere else)."""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import List, Optional, Sequence

import torch

from . import build as _build

F32, F16, BF16 = 0, 1, 2
FUSE = {"UV": 0, "U": 1, "V": 2}
STAT_ABS_MEAN, STAT_ABS_MAX, STAT_SQ_MEAN = 0, 1, 2
OK, ERR_INVALID, ERR_WORKSPACE, ERR_CUDA, ERR_NONFINITE, ERR_NOT_CONVERGED = 0, 1, 2, 3, 4, 5

_DTYPES = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}

EXPORTS = [
    "asvd_version", "asvd_last_error", "asvd_rank_for_ratio", "asvd_scaling_vector", "asvd_svd_workspace_bytes",
    "asvd_scaled_svd", "asvd_svd_sigma", "asvd_svd_extract", "asvd_lowrank_forward_scratch_bytes",
    "asvd_lowrank_forward", "asvd_absstat_scratch_bytes", "asvd_absstat_accum", "asvd_linear_forward_stat",
    "asvd_profile_enable", "asvd_profile_read", "asvd_launch_count",
]
KERNEL_CLASSES = ["prep", "gram", "solve", "update", "finalize", "extract", "forward", "absstat"]

_lock = threading.Lock()
_lib = None


class AsvdError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"asvd_b200 status {status}: {message}")
        self.status = status


def library_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Loads (building first if the in-tree library is missing or stale) libasvd_b200.so."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.build() if _build.stale() else _build.LIB
        lib = C.CDLL(path)
        vp, i64, i32, sz, f32, f64 = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_float, C.c_double
        lib.asvd_version.restype = i32
        lib.asvd_last_error.restype = C.c_char_p
        lib.asvd_rank_for_ratio.restype = i32
        lib.asvd_rank_for_ratio.argtypes = [i64, i64, f64, i32]
        lib.asvd_scaling_vector.restype = i32
        lib.asvd_scaling_vector.argtypes = [vp, vp, i32, i32, f64, vp, vp]
        lib.asvd_svd_workspace_bytes.restype = sz
        lib.asvd_svd_workspace_bytes.argtypes = [i32, i32, i32]
        lib.asvd_scaled_svd.restype = i32
        lib.asvd_scaled_svd.argtypes = [C.POINTER(vp), i32, i64, i32, i32, i32, C.POINTER(vp), vp, sz, f32, i32,
                                        C.POINTER(i32), vp]
        lib.asvd_svd_sigma.restype = i32
        lib.asvd_svd_sigma.argtypes = [vp, i32, i32, i32, i32, vp, vp]
        lib.asvd_svd_extract.restype = i32
        lib.asvd_svd_extract.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp, i64, vp]
        lib.asvd_lowrank_forward_scratch_bytes.restype = sz
        lib.asvd_lowrank_forward_scratch_bytes.argtypes = [i64, i32, i32]
        lib.asvd_lowrank_forward.restype = i32
        lib.asvd_lowrank_forward.argtypes = [vp, i64, i64, i32, vp, i64, i32, vp, i64, i32, vp, vp, i64, i32, vp, sz, vp]
        lib.asvd_absstat_scratch_bytes.restype = sz
        lib.asvd_absstat_scratch_bytes.argtypes = [i32]
        lib.asvd_absstat_accum.restype = i32
        lib.asvd_absstat_accum.argtypes = [vp, i64, i64, i32, i32, i32, vp, vp, sz, vp]
        lib.asvd_linear_forward_stat.restype = i32
        lib.asvd_linear_forward_stat.argtypes = [vp, i64, i64, i32, vp, i64, i32, vp, vp, i64, i32, i32, vp, vp, sz, vp]
        lib.asvd_profile_enable.restype = None
        lib.asvd_profile_enable.argtypes = [i32]
        lib.asvd_profile_read.restype = i32
        lib.asvd_profile_read.argtypes = [C.POINTER(f64), C.POINTER(C.c_uint64)]
        lib.asvd_launch_count.restype = C.c_uint64
        _lib = lib
        return lib


def _check(status: int, allow: Sequence[int] = ()):
    if status != OK and status not in allow:
        raise AsvdError(status, load().asvd_last_error().decode())
    return status


def _require_cuda(*tensors: torch.Tensor):
    if not torch.cuda.is_available():
        raise RuntimeError("asvd4llm_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("asvd4llm_b200 kernels take CUDA tensors; got a tensor on " + str(t.device))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(dt: torch.dtype) -> int:
    if dt not in _DTYPES:
        raise TypeError(f"unsupported dtype {dt}; expected float32, float16 or bfloat16")
    return _DTYPES[dt]


def rank_for_ratio(out_features: int, in_features: int, param_ratio: float, rank_align: int = 1) -> int:
    return load().asvd_rank_for_ratio(out_features, in_features, float(param_ratio), int(rank_align))


def scaling_vector(sdm: Optional[torch.Tensor], fisher: Optional[torch.Tensor], alpha: float, n: int,
                   device) -> torch.Tensor:
    """fp32 [n] = sdm**alpha * fisher**alpha + 1e-6 with the upstream rounding (svd_linear.py:48-59)."""
    ref = sdm if sdm is not None else fisher
    code = F32 if ref is None else dtype_code(ref.dtype)
    if sdm is not None and fisher is not None and sdm.dtype != fisher.dtype:
        fisher = fisher.to(sdm.dtype)
    sdm = None if sdm is None else sdm.to(device).contiguous()
    fisher = None if fisher is None else fisher.to(device).contiguous()
    _require_cuda(sdm, fisher)
    out = torch.empty(n, dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):          # launch on the tensors' device, not the process' current one
        _check(load().asvd_scaling_vector(None if sdm is None else sdm.data_ptr(), None if fisher is None else fisher.data_ptr(),
                                          code, n, float(alpha), out.data_ptr(), _stream()))
    return out


class Factorisation:
    """Finished activation-scaled SVDs of `batch` same-shape weights; owns the workspace (SURVEY.md F3:
    one SVD serves every truncation rank)."""

    def __init__(self, m: int, n: int, batch: int, workspace: torch.Tensor, sweeps: List[int], status: int):
        self.m, self.n, self.batch, self.workspace, self.sweeps, self.status = m, n, batch, workspace, sweeps, status

    def sigma(self, b: int = 0) -> torch.Tensor:
        out = torch.empty(min(self.m, self.n), dtype=torch.float32, device=self.workspace.device)
        with torch.cuda.device(self.workspace.device):
            _check(load().asvd_svd_sigma(self.workspace.data_ptr(), self.m, self.n, self.batch, b, out.data_ptr(), _stream()))
        return out

    def extract(self, r: int, sigma_fuse: str = "UV", dtype: torch.dtype = torch.float16, b: int = 0):
        """(ALinear.weight [m, r], BLinear.weight [r, n]) in `dtype` — svd_linear.py:8-24,69-70,102."""
        dev = self.workspace.device
        A = torch.empty(self.m, r, dtype=dtype, device=dev)
        B = torch.empty(r, self.n, dtype=dtype, device=dev)
        with torch.cuda.device(dev):
            _check(load().asvd_svd_extract(self.workspace.data_ptr(), self.m, self.n, self.batch, b, r, FUSE[sigma_fuse],
                                           dtype_code(dtype), A.data_ptr(), r, B.data_ptr(), self.n, _stream()))
        return A, B


SOLVE_CTAS_PER_SM = 3          # solve_tri_g_kernel: 160 threads x 128 registers, 74 KB of shared memory


def suggest_batch(m: int, n: int, device=None, limit_bytes: int = 32 << 30) -> int:
    """How many same-shape weights to factorise per call.  Every kernel of a Jacobi round works one block pair (128
    vectors) per CTA.  The inner eigen-solve (solve_tri_g_kernel) is a chain of 127 dependent rotation steps per pair that
    takes about as long for one CTA as for a full SM, and THREE of its CTAs share an SM, so batches are sized in whole
    solve waves of 3 x SMs block pairs: two of them by default (27 weights at 4096^2 on 148 SMs = 864 of 888 slots, 2.92
    of 3 waves of the two-CTA replay kernel, 5.84 of 6 waves of the streaming kernels).  Measured ms per matrix at
    4096^2 (profiles/r02_tri_*.log): 9 weights 44.2, 13 weights 45.0, 18 weights 44.3, 27 weights 40.4.
    ASVD_B200_WAVES overrides the number of solve waves; at most 32 weights, bounded by the workspace budget."""
    _require_cuda()
    sms = torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count
    pairs = (min(m, n) + 127) // 128
    lib = load()
    try:
        waves = max(1, int(os.environ.get("ASVD_B200_WAVES", "2")))
    except ValueError:
        waves = 2
    waves *= SOLVE_CTAS_PER_SM
    b = int(max(1, min(waves * sms // max(pairs, 1), 32)))
    while b > 1 and lib.asvd_svd_workspace_bytes(int(m), int(n), b) > limit_bytes:      # part of the workspace is per call
        b -= 1
    return b


def balanced_batches(n_items: int, cap: int):
    """Split n_items into ceil(n_items / cap) batches of nearly equal size (128 layers with cap 18 -> 8 x 16, not
    7 x 18 + 2: the inner solve of a 2-weight batch costs as much as that of a full wave)."""
    if n_items <= 0:
        return []
    nb = (n_items + cap - 1) // cap
    base, extra = divmod(n_items, nb)
    return [base + (1 if i < extra else 0) for i in range(nb)]


def scaled_svd(weights: Sequence[torch.Tensor], scales: Optional[Sequence[Optional[torch.Tensor]]] = None,
               tol: float = 0.0, max_sweeps: int = 0, allow_status: Sequence[int] = (ERR_NOT_CONVERGED,)) -> Factorisation:
    """Exact SVD of W_b * diag(scale_b) for a batch of same-shape CUDA weights [m, n]."""
    lib = load()
    _require_cuda(*weights)
    w0 = weights[0]
    m, n = w0.shape
    batch = len(weights)
    ws_list = []
    for w in weights:
        if tuple(w.shape) != (m, n) or w.dtype != w0.dtype or w.device != w0.device:
            raise ValueError("all weights of a batch must share shape, dtype and device")
        ws_list.append(w if (w.is_contiguous() and w.data_ptr() % 16 == 0) else w.contiguous().clone())
    sc_list = []
    for b in range(batch):
        s = None if scales is None else scales[b]
        if s is not None:
            s = s.to(device=w0.device, dtype=torch.float32).contiguous()
            if s.numel() != n:
                raise ValueError("scale must have in_features elements")
        sc_list.append(s)
    nbytes = lib.asvd_svd_workspace_bytes(m, n, batch)
    workspace = torch.empty(nbytes, dtype=torch.uint8, device=w0.device)
    Wp = (C.c_void_p * batch)(*[w.data_ptr() for w in ws_list])
    Sp = (C.c_void_p * batch)(*[(None if s is None else s.data_ptr()) for s in sc_list])
    sweeps = (C.c_int * batch)()
    with torch.cuda.device(w0.device):
        status = lib.asvd_scaled_svd(Wp, dtype_code(w0.dtype), ws_list[0].stride(0), m, n, batch, Sp, workspace.data_ptr(),
                                     nbytes, float(tol), int(max_sweeps), sweeps, _stream())
    _check(status, allow=allow_status)
    return Factorisation(m, n, batch, workspace, list(sweeps), status)


def pad_rank_stride(A: torch.Tensor) -> torch.Tensor:
    """[m, r] view of A whose row pitch is r rounded up to 64 elements (whole 128-byte lines for 16-bit types) -- what
    the tensor-core forward reads without a per-call copy.  Returns A itself when its rows are already 16-byte aligned."""
    m, r = A.shape
    if A.dtype == torch.float32 or (A.stride(1) == 1 and A.stride(0) % 8 == 0 and A.data_ptr() % 16 == 0):
        return A
    buf = torch.zeros(m, (r + 63) // 64 * 64, dtype=A.dtype, device=A.device)
    buf[:, :r].copy_(A)
    return buf[:, :r]


def _lowrank_forward_raw(x2: torch.Tensor, A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    lib = load()
    r, n = B.shape
    m = A.shape[0]
    M = x2.shape[0]
    y = torch.empty(M, m, dtype=x2.dtype, device=x2.device)
    if M == 0:
        return y
    nbytes = lib.asvd_lowrank_forward_scratch_bytes(M, r, m)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x2.device)        # caching allocator: 512-byte aligned
    with torch.cuda.device(x2.device):
        _check(lib.asvd_lowrank_forward(x2.data_ptr(), x2.stride(0), M, n, B.data_ptr(), B.stride(0), r, A.data_ptr(),
                                        A.stride(0), m, None if bias is None else bias.data_ptr(), y.data_ptr(), m,
                                        dtype_code(x2.dtype), scratch.data_ptr(), nbytes, _stream()))
    return y


class _LowRankForward(torch.autograd.Function):
    """Autograd wrapper: the forward is the C-ABI call; the backward (upstream's module is differentiable, but nothing
    on the hot path differentiates through a decomposed layer -- calib_fisher_info runs on the raw model) is three
    library matmuls."""

    @staticmethod
    def forward(ctx, x2, A, B, bias, A_kernel):
        ctx.save_for_backward(x2, A, B)
        ctx.has_bias = bias is not None
        return _lowrank_forward_raw(x2, A_kernel, B, bias)

    @staticmethod
    def backward(ctx, g):
        x2, A, B = ctx.saved_tensors
        g = g.contiguous()
        gt = g @ A                                             # [M, r]
        gx = gt @ B if ctx.needs_input_grad[0] else None
        gA = g.t() @ (x2 @ B.t()) if ctx.needs_input_grad[1] else None
        gB = gt.t() @ x2 if ctx.needs_input_grad[2] else None
        gb = g.sum(0) if (ctx.has_bias and ctx.needs_input_grad[3]) else None
        return gx, gA, gB, gb, None


def lowrank_forward(x: torch.Tensor, A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor],
                    A_kernel: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = (x B^T) A^T + bias — svd_linear.py:105-109.  A_kernel: optional pad_rank_stride(A) kept by the caller."""
    _require_cuda(x, A, B, bias)
    r, n = B.shape
    m = A.shape[0]
    if A.shape[1] != r or x.shape[-1] != n or (bias is not None and bias.numel() != m):
        raise RuntimeError(f"shape mismatch: x {tuple(x.shape)}, BLinear.weight {tuple(B.shape)}, ALinear.weight {tuple(A.shape)}")
    for name, t in (("BLinear.weight", B), ("ALinear.weight", A), ("ALinear.bias", bias)):
        if t is None:
            continue
        if t.dtype != x.dtype:      # what F.linear raises for upstream's nn.Linear children
            raise RuntimeError(f"expected input and {name} to have the same dtype, but got: {x.dtype} != {t.dtype}")
        if t.device != x.device:
            raise RuntimeError(f"expected all tensors to be on the same device, but found {x.device} and {t.device} ({name})")
    dtype_code(x.dtype)
    x2 = x.reshape(-1, n)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    A = A if A.stride(1) == 1 else A.contiguous()
    B = B if B.is_contiguous() else B.contiguous()
    if bias is not None and not bias.is_contiguous():
        bias = bias.contiguous()
    Ak = A if A_kernel is None else A_kernel
    needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, A, B, bias))
    if needs_grad:
        y = _LowRankForward.apply(x2, A, B, bias, Ak.detach())
    else:
        y = _lowrank_forward_raw(x2, Ak, B, bias)
    return y.reshape(*x.shape[:-1], m)


def absstat_accum(x: torch.Tensor, acc: torch.Tensor, method: str) -> None:
    """One hook call of act_aware_utils.py:64-74: acc [n] (same dtype as x) updated in place from x [.., L, n].
    method "sq_mean" is the Fisher statistic of act_aware_utils.py:31 (x = a weight gradient [m, n])."""
    _require_cuda(x, acc)
    lib = load()
    n = x.shape[-1]
    x2 = x.reshape(-1, n)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    if acc.dtype != x.dtype or acc.numel() != n or not acc.is_contiguous():
        raise ValueError("acc must be a contiguous [n] tensor of the activation dtype")
    if method == "sq_mean":
        mode = STAT_SQ_MEAN
    elif "abs_mean" in method:
        mode = STAT_ABS_MEAN
    elif "abs_max" in method:
        mode = STAT_ABS_MAX
    else:
        raise ValueError(f"unknown statistic {method!r}")
    nbytes = lib.asvd_absstat_scratch_bytes(n)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib.asvd_absstat_accum(x2.data_ptr(), x2.stride(0), x2.shape[0], n, dtype_code(x.dtype), mode,
                                      acc.data_ptr(), scratch.data_ptr(), nbytes, _stream()))


def linear_stat_eligible(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> bool:
    """Can asvd_linear_forward_stat take this layer call?  16-bit CUDA tensors of one dtype, rows 16-byte aligned."""
    if not (x.is_cuda and weight.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and weight.dtype == x.dtype):
        return False
    if bias is not None and (bias.dtype != x.dtype or not bias.is_cuda):
        return False
    m, n = weight.shape
    return (n % 8 == 0 and m % 8 == 0 and weight.stride(1) == 1 and weight.stride(0) % 8 == 0 and weight.data_ptr() % 16 == 0
            and x.shape[-1] == n and x.numel() > 0)


def linear_forward_stat(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], acc: torch.Tensor,
                        method: str) -> torch.Tensor:
    """One calibration step of one nn.Linear in one kernel: y = x W^T + bias, and acc [n] updated from |x| exactly as
    absstat_accum would (act_aware_utils.py:64-74) -- the statistic is a side output of the GEMM that consumes x."""
    _require_cuda(x, weight, bias, acc)
    lib = load()
    m, n = weight.shape
    x2 = x.reshape(-1, n)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    if acc.dtype != x.dtype or acc.numel() != n or not acc.is_contiguous():
        raise ValueError("acc must be a contiguous [n] tensor of the activation dtype")
    mode = STAT_ABS_MEAN if "abs_mean" in method else STAT_ABS_MAX
    M = x2.shape[0]
    y = torch.empty(M, m, dtype=x.dtype, device=x.device)
    scratch = torch.empty(4 * n, dtype=torch.uint8, device=x.device)
    if bias is not None and not bias.is_contiguous():
        bias = bias.contiguous()
    with torch.cuda.device(x.device):
        _check(lib.asvd_linear_forward_stat(x2.data_ptr(), x2.stride(0), M, n, weight.data_ptr(), weight.stride(0), m,
                                            None if bias is None else bias.data_ptr(), y.data_ptr(), m, dtype_code(x.dtype), mode,
                                            acc.data_ptr(), scratch.data_ptr(), 4 * n, _stream()))
    return y.reshape(*x.shape[:-1], m)


def launch_count() -> int:
    return int(load().asvd_launch_count())


def profile_enable(on: bool) -> None:
    load().asvd_profile_enable(1 if on else 0)


def profile_read():
    """{class: (total ms, launches since load)} — ms is only accumulated while profiling is enabled."""
    ms = (C.c_double * 8)()
    ln = (C.c_uint64 * 8)()
    load().asvd_profile_read(ms, ln)
    return {k: (ms[i], int(ln[i])) for i, k in enumerate(KERNEL_CLASSES)}
