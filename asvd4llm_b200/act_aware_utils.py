"""calib_input_distribution — upstream act_aware_utils.py:47-95 with the hook arithmetic (:64-74) in the
asvd_absstat_accum kernel — and calib_fisher_info — act_aware_utils.py:8-44 with the per-sample reduction of
the weight gradient (:31) in the same kernel.  Cache file names and formats are upstream's:
cache/{model_id with / -> _}_calib_input_distribution_{method}.pt and ..._calib_fisher_info.pt, each
{module name: Tensor[in_features]}."""
import os

import torch
import torch.nn as nn

from . import _lib


def atomic_save(obj, path):
    """torch.save through a temporary file + os.replace: under torchrun every rank computes the same table and writes the
    same cache path; a reader (or another writer) never sees a half-written file."""
    tmp = f"{path}.tmp{os.getpid()}"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def _cache_file(model, method):
    model_id = model.config._name_or_path
    return f"cache/{model_id.replace('/', '_')}_calib_input_distribution_{method}.pt"


@torch.no_grad()
def calib_input_distribution(model, calib_loader, method, use_cache=True):
    cache_file = _cache_file(model, method)
    if os.path.exists(cache_file) and use_cache:
        table = torch.load(cache_file, map_location="cpu")
        for name, module in model.named_modules():
            if isinstance(module, nn.Linear):
                module.scaling_diag_matrix = table[name].to(module.weight.device)
        return
    model.eval()

    def check_batch1(x):
        if x.dim() > 2 and x.numel() // (x.shape[-1] * x.shape[-2]) != 1:
            # upstream's `.view(-1)` (:66) is only meaningful for batch 1
            raise ValueError("calib_input_distribution expects batch-1 activations, got " + str(tuple(x.shape)))

    def accumulator(module, x):
        acc = module.scaling_diag_matrix
        if not torch.is_tensor(acc):                           # python int 0 until the first call (:80)
            acc = torch.zeros(x.shape[-1], dtype=x.dtype, device=x.device)
            module.scaling_diag_matrix = acc
        return acc

    def hook(module, inp, out):
        x = inp[0].detach()
        check_batch1(x)
        acc = accumulator(module, x)
        if "abs_mean" in method or "abs_max" in method:
            _lib.absstat_accum(x, acc, method)

    def fused_forward(module):
        # the layer's own GEMM produces the statistic of its input as a side output
        # (asvd_linear_forward_stat) -- no hook, no second pass over the activation
        def forward(x):
            if torch.is_grad_enabled() or not _lib.linear_stat_eligible(x, module.weight, module.bias):
                y = nn.functional.linear(x, module.weight, module.bias)
                hook(module, (x,), y)
                return y
            check_batch1(x)
            return _lib.linear_forward_stat(x.detach(), module.weight.data, None if module.bias is None else module.bias.data,
                                            accumulator(module, x), method)
        return forward

    # ASVD_B200_CALIB=hook keeps upstream's structure (model forward by torch, statistic in a forward hook)
    fused = os.environ.get("ASVD_B200_CALIB", "fused") != "hook" and ("abs_mean" in method or "abs_max" in method)
    handles = []
    patched = []
    for _, module in model.named_modules():
        if isinstance(module, nn.Linear):
            module.scaling_diag_matrix = 0
            # (a module whose `forward` is already an instance attribute -- accelerate's device_map hooks wrap it that way --
            # keeps its wrapper and gets the plain hook)
            if (fused and type(module) is nn.Linear and "forward" not in vars(module) and module.weight.is_cuda
                    and module.weight.dtype in (torch.float16, torch.bfloat16)):
                module.forward = fused_forward(module)
                patched.append(module)
            else:
                handles.append(module.register_forward_hook(hook))
    device = getattr(model, "device", None) or next(model.parameters()).device
    try:
        for batch in calib_loader:
            batch = {k: v.to(device) for k, v in batch.items()}
            model(**batch)
    finally:
        for module in patched:
            del module.forward                                  # back to nn.Linear.forward
    table = {}
    for name, module in model.named_modules():
        if isinstance(module, nn.Linear):
            module._forward_hooks.clear()                      # upstream clears every hook (:93)
            table[name] = module.scaling_diag_matrix
    atomic_save(table, cache_file)


def calib_fisher_info(model, calib_loader, use_cache=True):
    """Upstream act_aware_utils.py:8-44: fisher_info[j] = sqrt(mean over samples of mean_i (dL/dW[i, j])^2).
    The backward pass is the model's own (autograd, as upstream); the [m, n] -> [n] reduction of every weight
    gradient runs in asvd_absstat_accum (SQ_MEAN) instead of `grad.pow(2).mean(0)`, which materialises an
    [m, n] temporary per linear and sample."""
    model_id = model.config._name_or_path
    cache_file = f"cache/{model_id.replace('/', '_')}_calib_fisher_info.pt"
    if os.path.exists(cache_file) and use_cache:
        table = torch.load(cache_file, map_location="cpu")
        for name, module in model.named_modules():
            if isinstance(module, nn.Linear):
                module.fisher_info = table[name].to(module.weight.device)
        return
    model.eval()
    linears = [(name, module) for name, module in model.named_modules() if isinstance(module, nn.Linear)]
    for _, module in linears:
        module.fisher_info = 0
    device = getattr(model, "device", None) or next(model.parameters()).device
    for batch in calib_loader:
        input_ids = batch["input_ids"][:, :-1].to(device)
        labels = batch["input_ids"][:, 1:].to(device)
        with torch.enable_grad():
            out = model(input_ids=input_ids, labels=labels)
            out[0].backward()
        with torch.no_grad():
            for _, module in linears:
                g = module.weight.grad
                if g is None:                                   # upstream would raise AttributeError here (:31)
                    raise AttributeError("'NoneType' object has no attribute 'detach'")
                acc = module.fisher_info
                if not torch.is_tensor(acc):                    # python int 0 until the first sample (:21)
                    acc = torch.zeros(g.shape[1], dtype=g.dtype, device=g.device)
                    module.fisher_info = acc
                _lib.absstat_accum(g.detach(), acc, "sq_mean")
        model.zero_grad()
    table = {}
    with torch.no_grad():
        for name, module in linears:
            module.fisher_info = module.fisher_info.div(len(calib_loader)).sqrt()     # :36
            module._forward_hooks.clear()                                             # :42
            table[name] = module.fisher_info
    atomic_save(table, cache_file)
