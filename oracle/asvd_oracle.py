"""CPU oracle for the ASVD hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (asvd4llm_b200) never does and fails loudly when its CUDA library is missing.

It restates, in plain CPU torch / numpy, what hahnyuan/ASVD4LLM computes on this path.  Citations are
`file:line` inside the upstream tree (mounted at /root/reference while the repo was built; it does not
exist on the GPU box, so nothing here reads it).

Pinning status: the upstream repository ships no tests, golden vectors or fixtures for this path
(SURVEY.md F4), so parity is pinned two ways instead:
  * tests/golden/*.pt hold outputs of the UNMODIFIED upstream code run in the build container
    (tests/golden/make_golden.py is the generating script); tests/test_oracle_golden.py checks every
    function below against them;
  * the arithmetic dependency outside the upstream tree is PyTorch (requirements.txt:3, unpinned; 2.11.0
    here): torch.svd_lowrank (torch/_lowrank.py, Halko et al. 2009 Alg. 4.4 + 5.1, niter=2, q = rank, no
    oversampling) calling torch.linalg.qr / torch.linalg.svd.  `factorise_lowrank` restates that published
    algorithm; `factorise_exact` is the grading oracle BASELINE.json's north_star names
    (torch.linalg.svd of the scaled weight), in fp32 and fp64.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

RATIO_CANDIDATES = [0.4, 0.5, 0.6, 0.7, 0.8, 0.9]          # sensitivity.py:39
KV_RATIO_CANDIDATES = [0.1 * i for i in range(1, 20)]       # sensitivity.py:37


# ----------------------------------------------------------------------------- a2: rank formula
def rank_for_ratio(out_features: int, in_features: int, param_ratio: float, rank_align: int = 1) -> int:
    """modules/svd_linear.py:39-44 — r = ceil((int(m*n*ratio) // (m+n)) / align) * align."""
    compressed = int(out_features * in_features * param_ratio)
    rank = compressed // (in_features + out_features)
    return int(np.ceil(rank / rank_align) * rank_align)


# ----------------------------------------------------------------------------- a3: scaling vector
def scaling_vector(sdm: Optional[torch.Tensor], fisher: Optional[torch.Tensor], alpha: float):
    """modules/svd_linear.py:48-59 — s = sdm**alpha * fisher**alpha + 1e-6, evaluated in the dtype of the
    statistics (fp16 on the upstream GPU path), exactly as the in-place python expressions do."""
    s = 1
    if sdm is not None:
        s = s * sdm ** alpha
    if fisher is not None:
        s = s * fisher ** alpha
    s = s + 1e-6
    return s


# ----------------------------------------------------------------------------- a6: sigma fusion
def fuse_sigma(U: torch.Tensor, S: torch.Tensor, V: torch.Tensor, sigma_fuse: str = "UV"):
    """modules/svd_linear.py:16-24 — returns (ALinear.weight [m,r], BLinear.weight [r,n])."""
    if sigma_fuse == "UV":
        return U * S.sqrt(), V.t() * S.sqrt().view(-1, 1)
    if sigma_fuse == "U":
        return U * S, V.t().clone()
    if sigma_fuse == "V":
        return U.clone(), V.t() * S.view(-1, 1)
    raise ValueError(sigma_fuse)


# ----------------------------------------------------------------------------- a3-a6 with an exact SVD
def factorise_exact(W: torch.Tensor, param_ratio: float, sdm=None, fisher=None, alpha: float = 1.0,
                    act_aware: bool = False, sigma_fuse: str = "UV", rank_align: int = 1,
                    compute_dtype=torch.float32):
    """The north_star oracle: modules/svd_linear.py:39-102 with `torch.linalg.svd` in place of line 65.

    Returns dict(A [m,r], B [r,n] in compute_dtype, S_all [min(m,n)], rank, U, S, V).  The truncation keeps
    min(rank, min(m,n)) triplets."""
    m, n = W.shape
    rank = rank_for_ratio(m, n, param_ratio, rank_align)
    w = W.detach().cpu().float().to(compute_dtype)
    s = None
    if act_aware:
        s = scaling_vector(sdm, fisher, alpha)
        s = s.detach().cpu().to(compute_dtype) if torch.is_tensor(s) else torch.full((n,), float(s), dtype=compute_dtype)
        w = w * s.view(1, -1)
    U, S, Vh = torch.linalg.svd(w, full_matrices=False)
    r = min(rank, S.numel())
    Ur, Sr, Vr = U[:, :r], S[:r], Vh[:r].t()
    if act_aware:
        Vr = Vr / s.view(-1, 1)                                   # svd_linear.py:69-70
    A, B = fuse_sigma(Ur, Sr, Vr, sigma_fuse)
    return dict(A=A.contiguous(), B=B.contiguous(), S_all=S, rank=r, U=Ur, S=Sr, V=Vr, scaled=w)


# ----------------------------------------------------------------------------- a4 as shipped (randomised)
def svd_lowrank_restated(A: torch.Tensor, q: int, niter: int = 2, generator: Optional[torch.Generator] = None):
    """torch/_lowrank.py:60-82,149-180 restated: Halko Alg. 4.4 (subspace iteration, QR re-orthogonalised)
    + Alg. 5.1 (SVD of the projected matrix).  Draws R from the global torch RNG unless `generator`."""
    m, n = A.shape
    transposed = m < n
    if transposed:
        A = A.t()
    R = torch.randn(A.shape[1], q, dtype=A.dtype, generator=generator)
    Q = torch.linalg.qr(A @ R).Q
    for _ in range(niter):
        Q = torch.linalg.qr(A.t() @ Q).Q
        Q = torch.linalg.qr(A @ Q).Q
    Bm = Q.t() @ A
    U, S, Vh = torch.linalg.svd(Bm, full_matrices=False)
    V = Vh.t()
    U = Q @ U
    if transposed:
        U, V = V, U
    return U, S, V


def factorise_lowrank(W, param_ratio, sdm=None, fisher=None, alpha=1.0, act_aware=False, sigma_fuse="UV",
                      rank_align=1, seed: Optional[int] = None):
    """modules/svd_linear.py:39-102 as shipped (svd_lowrank, q = rank).  RNG-dependent; `seed` pins it."""
    m, n = W.shape
    rank = rank_for_ratio(m, n, param_ratio, rank_align)
    w = W.detach().cpu().float()
    s = None
    if act_aware:
        s = scaling_vector(sdm, fisher, alpha)
        s = s.detach().cpu().float() if torch.is_tensor(s) else torch.full((n,), float(s))
        w = w * s.view(1, -1)
    if seed is not None:
        torch.manual_seed(seed)
    U, S, V = svd_lowrank_restated(w, rank)
    if act_aware:
        V = V / s.view(-1, 1)
    A, B = fuse_sigma(U, S, V, sigma_fuse)
    return dict(A=A.contiguous(), B=B.contiguous(), S=S, rank=S.numel(), scaled=w)


# ----------------------------------------------------------------------------- a7: forward
def lowrank_forward(x: torch.Tensor, A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor] = None,
                    compute_dtype=None):
    """modules/svd_linear.py:105-109 — y = ALinear(BLinear(x)); the [.., r] intermediate is materialised in
    the module dtype (fp16 on the GPU path).  compute_dtype=torch.float64 gives the F7 error oracle."""
    if compute_dtype is not None:
        x, A, B = x.to(compute_dtype), A.to(compute_dtype), B.to(compute_dtype)
        bias = None if bias is None else bias.to(compute_dtype)
    t = torch.nn.functional.linear(x, B)
    return torch.nn.functional.linear(t, A, bias)


# ----------------------------------------------------------------------------- a1: calibration statistic
def abs_stat_update(acc, x: torch.Tensor, method: str):
    """act_aware_utils.py:64-74 — one hook call. acc starts as python int 0 (:80)."""
    if "abs_mean" in method:
        return acc + x.abs().mean(dim=-2).detach().view(-1)
    if "abs_max" in method:
        cur = x.abs().amax(dim=-2).detach().view(-1)
        if not torch.is_tensor(acc):
            acc = torch.full_like(cur, float(acc))
        return torch.where(cur > acc, cur, acc)
    return acc


def calib_input_distribution(model: nn.Module, calib_loader, method: str) -> Dict[str, torch.Tensor]:
    """act_aware_utils.py:47-95 without the cache file: returns {module name: Tensor[n]} and sets
    module.scaling_diag_matrix."""
    model.eval()
    hooks = []

    def hook(module, inp, out):
        module.scaling_diag_matrix = abs_stat_update(module.scaling_diag_matrix, inp[0], method)

    for _, mod in model.named_modules():
        if isinstance(mod, nn.Linear):
            mod.scaling_diag_matrix = 0
            hooks.append(mod.register_forward_hook(hook))
    with torch.no_grad():
        for batch in calib_loader:
            model(**{k: v.to(next(model.parameters()).device) for k, v in batch.items()})
    for h in hooks:
        h.remove()
    return {name: mod.scaling_diag_matrix for name, mod in model.named_modules() if isinstance(mod, nn.Linear)}


def fisher_stat_update(acc, grad: torch.Tensor):
    """act_aware_utils.py:31 — one sample's contribution: acc += grad.pow(2).mean(0), in the gradient's dtype."""
    return acc + grad.detach().pow(2).mean(0)


def calib_fisher_info(model: nn.Module, calib_loader) -> Dict[str, torch.Tensor]:
    """act_aware_utils.py:8-44 without the cache file: fisher_info = sqrt(sum_samples mean_i grad[i, :]^2 / N).
    Returns {module name: Tensor[n]} and sets module.fisher_info."""
    model.eval()
    linears = [(n, m) for n, m in model.named_modules() if isinstance(m, nn.Linear)]
    for _, mod in linears:
        mod.fisher_info = 0
    dev = next(model.parameters()).device
    for batch in calib_loader:
        input_ids = batch["input_ids"][:, :-1].to(dev)
        labels = batch["input_ids"][:, 1:].to(dev)
        out = model(input_ids=input_ids, labels=labels)
        out[0].backward()
        for _, mod in linears:
            mod.fisher_info = fisher_stat_update(mod.fisher_info, mod.weight.grad)
        model.zero_grad()
    for _, mod in linears:
        mod.fisher_info = mod.fisher_info.div(len(calib_loader)).sqrt()
    return {n: m.fisher_info for n, m in linears}


# ----------------------------------------------------------------------------- a10: perplexity
@torch.no_grad()
def evaluate_perplexity(model, dataset: torch.Tensor, limit: int) -> float:
    """evaluate_utils.py:90-115 — exp(mean over samples of mean CE over seqlen-1 shifted tokens), batch 1."""
    nsamples, seqlen = dataset.size()
    nlls = []
    for i in range(nsamples):
        if i == limit:
            break
        ids = dataset[i:i + 1, :-1]
        labels = dataset[i:i + 1, 1:].contiguous()
        logits = model(input_ids=ids)[0]
        loss = nn.CrossEntropyLoss()(logits.view(-1, logits.size(-1)), labels.view(-1))
        nlls.append(loss.float() * seqlen)
    return torch.exp(torch.stack(nlls).sum() / (len(nlls) * seqlen)).item()


# ----------------------------------------------------------------------------- a8/a9 host logic
def enumerate_linears(model: nn.Module) -> List[Tuple[nn.Module, str, str, nn.Linear]]:
    """sensitivity.py:19-33 / binary_search.py:11-27 — explicit-stack DFS, last-registered child first.
    Returns [(father, child_name, full_name, linear)] in sweep order."""
    full_name = {mod: name for name, mod in model.named_modules()}
    out, stack = [], [model]
    while stack:
        sub = stack.pop()
        for name, child in sub.named_children():
            if isinstance(child, nn.Linear):
                out.append((sub, name, full_name[child], child))
            else:
                stack.append(child)
    return out


def allocate_ratios(sensitivity: Dict[str, Dict[float, float]], numel: Dict[str, int], ratio_target: float,
                    compress_kv_cache: bool = False) -> Tuple[Dict[str, float], int]:
    """binary_search.py:29-110, ratio-target mode: returns ({layer: ratio or default}, stale mid).
    Default ratio (= leave the layer raw) is 1, or 2 in kv-cache mode."""
    if compress_kv_cache:
        sensitivity = {k: v for k, v in sensitivity.items() if "k_proj" in k or "v_proj" in k}
        default = 2
    else:
        default = 1
    flat = []
    for layer, table in sensitivity.items():
        for ratio, ppl in table.items():
            if not compress_kv_cache and ratio >= 1:
                continue
            flat.append((layer, ratio, ppl))
    flat = sorted(flat, key=lambda t: -t[2])
    low, high, mid = 0, len(flat) - 1, None
    while low < high:
        mid = (low + high) // 2
        chosen = {k: default for k in sensitivity}
        for layer, ratio, _ in flat[mid:]:
            chosen[layer] = min(chosen[layer], ratio)
        tot = sum(numel[k] for k in chosen)
        comp = sum(numel[k] * r for k, r in chosen.items())
        now = comp / tot
        if compress_kv_cache:
            now /= 2
        if now > ratio_target:
            high = mid
        else:
            low = mid + 1
    chosen = {k: default for k in sensitivity}
    for layer, ratio, _ in flat[mid:]:                      # stale `mid`, binary_search.py:106 (quirk 3)
        chosen[layer] = min(chosen[layer], ratio)
    return chosen, mid


# ----------------------------------------------------------------------------- synthetic inputs
def synthetic_weight(m: int, n: int, seed: int = 233, kind: str = "gauss", dtype=torch.float16):
    """Config-2 style inputs: W ~ N(0, 0.02^2) in fp16 (or a power-law spectrum), s = exp(N(0,1)) fp32."""
    g = torch.Generator().manual_seed(seed)
    if kind == "gauss":
        W = torch.randn(m, n, generator=g) * 0.02
    elif kind == "power":
        k = min(m, n)
        U = torch.linalg.qr(torch.randn(m, k, generator=g)).Q
        V = torch.linalg.qr(torch.randn(n, k, generator=g)).Q
        sv = torch.arange(1, k + 1, dtype=torch.float32) ** -1.0
        W = (U * sv) @ V.t()
        W = W / W.abs().max()
    else:
        raise ValueError(kind)
    s = torch.exp(torch.randn(n, generator=g))
    return W.to(dtype), s


def relative_sigma_error(S: torch.Tensor, S_ref: torch.Tensor, r: int) -> float:
    return ((S[:r].double() - S_ref[:r].double()).abs() / S_ref[:r].double()).max().item()


# ----------------------------------------------------------------------------- module-level restatement
class OracleSVDLinear(nn.Module):
    """modules/svd_linear.py:7-24,105-109 — two nn.Linear children named ALinear / BLinear."""

    def __init__(self, A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor], rank: int):
        super().__init__()
        self.ALinear = nn.Linear(A.size(1), A.size(0), bias=bias is not None)
        if bias is not None:
            self.ALinear.bias.data = bias
        self.BLinear = nn.Linear(B.size(1), B.size(0), bias=False)
        self.ALinear.weight.data = A.contiguous()
        self.BLinear.weight.data = B.contiguous()
        self.truncation_rank = rank

    def forward(self, inp):
        return self.ALinear(self.BLinear(inp))


def from_linear(linear: nn.Linear, param_ratio: float, act_aware=False, alpha=1, sigma_fuse="UV", rank_align=1,
                method: str = "lowrank"):
    """modules/svd_linear.py:26-103.  method='lowrank' is the shipped behaviour (consumes the global torch
    RNG exactly once, like torch.svd_lowrank); method='exact' is the north_star oracle."""
    sdm = getattr(linear, "scaling_diag_matrix", None)
    fisher = getattr(linear, "fisher_info", None)
    fn = factorise_lowrank if method == "lowrank" else factorise_exact
    out = fn(linear.weight.data, param_ratio, sdm=sdm, fisher=fisher, alpha=alpha, act_aware=act_aware,
             sigma_fuse=sigma_fuse, rank_align=rank_align)
    bias = linear.bias.data if linear.bias is not None else None
    mod = OracleSVDLinear(out["A"], out["B"], bias, out["rank"])
    mod.to(linear.weight.dtype)
    return mod


@torch.no_grad()
def calib_sensitivity_ppl(model, calib_loader, alpha: float, n_calib_samples: int, rank_align: int = 1,
                          compress_kv_cache: bool = False, method: str = "lowrank"):
    """sensitivity.py:35-61 without the cache file."""
    model.eval()
    table: Dict[str, Dict[float, float]] = {}
    ratios = KV_RATIO_CANDIDATES if compress_kv_cache else RATIO_CANDIDATES
    ids = torch.cat([b["input_ids"] for b in calib_loader], 0)
    for father, name, full_name, raw in enumerate_linears(model):
        table[full_name] = {}
        for ratio in ratios:
            setattr(father, name, from_linear(raw, ratio, act_aware=True, alpha=alpha, rank_align=rank_align,
                                              method=method))
            table[full_name][ratio] = evaluate_perplexity(model, ids, n_calib_samples)
        setattr(father, name, raw)
    return table


def decompose_model(model, chosen: Dict[str, float], default_ratio, alpha, act_aware, sigma_fuse="UV",
                    rank_align=1, method="lowrank"):
    """binary_search.py:112-128 — final pass; layers at the default ratio stay raw."""
    by_name = dict(model.named_modules())
    info = {lin: (father, name) for father, name, _, lin in enumerate_linears(model)}
    for layer, ratio in chosen.items():
        raw = by_name[layer]
        if ratio == default_ratio:
            continue
        father, name = info[raw]
        setattr(father, name, from_linear(raw, ratio, act_aware=act_aware, alpha=alpha, sigma_fuse=sigma_fuse,
                                          rank_align=rank_align, method=method))


def binary_search_truncation_rank(model, sensitivity: Dict[str, Dict[float, float]], calib_loader, ppl_target: float = -1,
                                  param_ratio_target: float = -1, alpha: float = 0.5, act_aware: bool = True,
                                  sigma_fuse: str = "UV", rank_align: int = 1, n_calib_samples: int = 3,
                                  method: str = "lowrank") -> List[str]:
    """binary_search.py:10-131 in full (weight mode), including the --ppl_target branch (:64-87): every iteration
    decomposes EVERY layer of the table from its raw nn.Linear at the current ratio (ratio 1 included: from_linear then
    gives rank m*n // (m+n)), evaluates the calibration perplexity and bisects on it; the final pass (:104-128) re-uses
    the LAST mid, restores raw layers at the default ratio and decomposes the others.  Returns upstream's log lines.
    method='lowrank' consumes the global torch RNG exactly like upstream's torch.svd_lowrank calls."""
    by_name = dict(model.named_modules())
    info = {lin: (father, name) for father, name, _, lin in enumerate_linears(model)}      # captured before any replacement
    default = 1
    log = [f"=== compress weight target: ppl={ppl_target}, ratio_target={param_ratio_target} ==="]
    flat = [(layer, ratio, ppl) for layer, table in sensitivity.items() for ratio, ppl in table.items() if ratio < 1]
    flat = sorted(flat, key=lambda t: -t[2])
    low, high, mid = 0, len(flat) - 1, None
    ids = torch.cat([b["input_ids"] for b in calib_loader], 0)
    while low < high:
        mid = (low + high) // 2
        chosen = {k: default for k in sensitivity}
        for layer, ratio, _ in flat[mid:]:
            chosen[layer] = min(chosen[layer], ratio)
        tot = comp = 0
        if ppl_target > 0:
            for layer, ratio in chosen.items():
                raw = by_name[layer]
                father, name = info[raw]
                setattr(father, name, from_linear(raw, ratio, act_aware=act_aware, alpha=alpha, sigma_fuse=sigma_fuse,
                                                  rank_align=rank_align, method=method))
                tot += raw.weight.numel()
                comp += raw.weight.numel() * ratio
            ppl = evaluate_perplexity(model, ids, n_calib_samples)
            log.append(f"low={low} mid={mid}, high={high}, ppl={ppl}, param_ratio={comp / tot}")
            if ppl < ppl_target:
                high = mid
            else:
                low = mid + 1
        else:
            for layer, ratio in chosen.items():
                tot += by_name[layer].weight.numel()
                comp += by_name[layer].weight.numel() * ratio
            now = comp / tot
            log.append(f"low={low} mid={mid}, high={high}, now_ratio={now}, params=({comp}/{tot})")
            if now > param_ratio_target:
                high = mid
            else:
                low = mid + 1
    log.append("=== Searching done, decomposing layers... ===")
    chosen = {k: default for k in sensitivity}
    for layer, ratio, _ in flat[mid:]:
        chosen[layer] = min(chosen[layer], ratio)
    for layer, ratio in chosen.items():
        raw = by_name[layer]
        father, name = info[raw]
        setattr(father, name, raw if ratio == default else from_linear(raw, ratio, act_aware=act_aware, alpha=alpha,
                                                                       sigma_fuse=sigma_fuse, rank_align=rank_align,
                                                                       method=method))
    return log
