"""calib_sensitivity_ppl — upstream sensitivity.py:10-61.  Same table ({layer: {ratio: ppl}}, python floats),
same sweep order, same cache file; the six factorisations per layer are six slices of ONE SVD."""
import os

import torch
import torch.nn as nn

from .act_aware_utils import atomic_save
from .evaluate_utils import evaluate_perplexity
from .modules.svd_linear import SVDLinear, clear_cache

RATIOS = [0.4, 0.5, 0.6, 0.7, 0.8, 0.9]                       # sensitivity.py:39
KV_RATIOS = [0.1 * i for i in range(1, 20)]                  # sensitivity.py:37 (float keys like 0.30000000000000004)


def enumerate_linears(model):
    """upstream's explicit-stack DFS (sensitivity.py:19-33): last-registered child first, stops at nn.Linear.
    Yields (father, child name, full name, linear) in sweep order."""
    full_name = {module: name for name, module in model.named_modules()}
    found, stack = [], [model]
    while stack:
        sub = stack.pop()
        for name, child in sub.named_children():
            if isinstance(child, nn.Linear):
                found.append((sub, name, full_name[child], child))
            else:
                stack.append(child)
    return found


def sensitivity_cache_file(model, args):
    model_id = model.config._name_or_path
    return (f"cache/{model_id.replace('/', '_')}_sensitivity_{args.scaling_method}_{args.alpha}_"
            f"{args.n_calib_samples}_{args.calib_dataset}.pt")


@torch.no_grad()
def calib_sensitivity_ppl(model, calib_loader, args, use_cache=True, layer_filter=None, unit_filter=None):
    """Extensions, both off by default:
      * layer_filter: callable(full_name) -> bool, or unit_filter: callable(unit index) -> bool over the flattened
        (layer, ratio) sweep order -- restrict the sweep to a shard; the multi-GPU driver merges the shards
        (asvd4llm_b200.sharding.gather_sensitivity).  With one SVD serving the six ratios of a layer the sweep is > 95 %
        model forwards, so (layer, ratio) units balance the ranks better than layers at the cost of one extra SVD per
        rank that shares a layer (asvd.py deals contiguous unit ranges, so only world - 1 layers are shared);
      * args.eval_batch_size (asvd.py --eval_batch_size): calibration samples per model forward in evaluate_perplexity
        (upstream: 1).  Same per-sample losses up to floating-point reassociation inside the batched GEMMs."""
    cache_file = sensitivity_cache_file(model, args)
    sharded = layer_filter is not None or unit_filter is not None
    if os.path.exists(cache_file) and use_cache and not sharded:
        return torch.load(cache_file, map_location="cpu")
    model.eval()
    ratios = KV_RATIOS if args.compress_kv_cache else RATIOS
    input_ids = torch.cat([b["input_ids"] for b in calib_loader], 0)
    print(f"input_ids.shape={input_ids.shape}")
    eval_bs = max(1, int(getattr(args, "eval_batch_size", 1) or 1))
    table = {}
    unit = 0
    for father, name, full_name, raw in enumerate_linears(model):
        if layer_filter is not None and not layer_filter(full_name):
            unit += len(ratios)
            continue
        for ratio in ratios:
            mine = unit_filter is None or unit_filter(unit)
            unit += 1
            if not mine:
                continue
            svd_linear = SVDLinear.from_linear(raw, param_ratio=ratio, alpha=args.alpha, act_aware=True,
                                               rank_align=args.rank_align)       # act_aware hard-coded (:50)
            setattr(father, name, svd_linear)
            ppl = evaluate_perplexity(model, input_ids, args.n_calib_samples, batch_size=eval_bs)
            table.setdefault(full_name, {})[ratio] = ppl
            print(f"{full_name} {ratio} {ppl}")
        setattr(father, name, raw)
    clear_cache()
    if not sharded:
        atomic_save(table, cache_file)
    return table


@torch.no_grad()
def calib_sensitivity_stable_rank(model, calib_loader, args, use_cache=True):
    """upstream sensitivity.py:64-110 — sensitivity = -sqrt(|W|_F^2 / sigma_max^2) * ratio^0.1 for nine ratios, values
    0-d tensors as upstream.  The singular values come from asvd_scaled_svd (no scaling), which also gives
    |W|_F^2 = sum sigma^2."""
    from . import _lib
    model_id = model.config._name_or_path
    cache_file = (f"cache/{model_id.replace('/', '_')}_sensitivity_stable_rank_{args.scaling_method}_{args.alpha}_"
                  f"{args.n_calib_samples}_{args.calib_dataset}.pt")
    if os.path.exists(cache_file) and use_cache:
        return torch.load(cache_file, map_location="cpu")
    model.eval()
    ratios = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]
    input_ids = torch.cat([b["input_ids"] for b in calib_loader], 0)
    print(f"input_ids.shape={input_ids.shape}")
    table = {}
    for father, name, full_name, raw in enumerate_linears(model):
        w = raw.weight.data
        dev = w.device if w.is_cuda else torch.device("cuda", torch.cuda.current_device())
        sigma = _lib.scaled_svd([w.to(dev)], [None]).sigma(0)
        sr = ((sigma.double() ** 2).sum() / sigma[0].double() ** 2).sqrt().to(w.dtype).to(w.device)
        table[full_name] = {ratio: -sr * ratio ** 0.1 for ratio in ratios}
    atomic_save(table, cache_file)
    return table
