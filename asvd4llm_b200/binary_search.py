"""binary_search_truncation_rank — upstream binary_search.py:10-131: global rank allocation over the
(layer, ratio, ppl) list and the final decomposition of every selected layer.  Quirks kept on purpose
(SURVEY.md §9): ratio >= 1 entries dropped outside kv mode, kv mode halves the ratio, the final allocation
uses the LAST `mid` of the loop, log lines are upstream's."""
import os
import time
from collections import defaultdict

import torch

from . import _lib
from .evaluate_utils import evaluate_perplexity
from .modules.svd_linear import from_linear_batch, clear_cache
from .sensitivity import enumerate_linears


def _ratios_after_cut(sorted_list, cut, layer_names, default_ratio):
    chosen = {name: default_ratio for name in layer_names}
    for layer, ratio, _ in sorted_list[cut:]:
        chosen[layer] = min(chosen[layer], ratio)
    return chosen


class LinearIndex:
    """name -> raw nn.Linear and raw -> (father, child name), captured ONCE before anything is replaced -- upstream's
    module_dict / linear_info (binary_search.py:11-27).  The ppl-target search installs SVDLinears in every iteration;
    every later step must still find the raw layers."""

    def __init__(self, model):
        self.by_name = dict(model.named_modules())
        self.where = {lin: (father, name) for father, name, _, lin in enumerate_linears(model)}
        # a weight tied to another module (OPT: lm_head <-> embed_tokens) must stay where it is: upstream's
        # unconditional `raw_linear.to("cpu")` (:127) would drag the embedding to the CPU with it on a GPU run
        self.uses = defaultdict(int)
        for mod in model.modules():
            for prm in mod._parameters.values():
                if prm is not None:
                    self.uses[id(prm)] += 1


def _install(index, chosen, default_ratio, args, layer_filter=None, batch_limit_bytes=16 << 30, final=True,
             decompose_default=False):
    """Replaces every selected layer by its SVDLinear, same-shape layers batched per kernel call.
    final=True is upstream's last pass (:112-128): default-ratio layers get the RAW linear back and replaced weights
    move to the CPU.  decompose_default=True is the ppl-target loop (:64-76), which decomposes every layer at its ratio."""
    groups = defaultdict(list)
    done = 0
    for layer, ratio in chosen.items():
        raw = index.by_name[layer]
        if layer_filter is not None and not layer_filter(layer):
            if final:                                           # another rank's layer: hold the raw layer until its factors arrive
                father, name = index.where[raw]
                setattr(father, name, raw)
            continue
        if ratio == default_ratio and not decompose_default:
            father, name = index.where[raw]
            setattr(father, name, raw)                          # :116-117 (undoes a ppl-target search's replacement)
            continue
        groups[(tuple(raw.weight.shape), raw.weight.dtype, raw.weight.device)].append((layer, ratio, raw))
    for (shape, _, _), items in groups.items():
        m, n = shape
        cap = _lib.suggest_batch(m, n, limit_bytes=batch_limit_bytes)
        i = 0
        for size in _lib.balanced_batches(len(items), cap):
            part = items[i:i + size]
            i += size
            mods = from_linear_batch([raw for _, _, raw in part], [ratio for _, ratio, _ in part], alpha=args.alpha,
                                     act_aware=args.act_aware, sigma_fuse=args.sigma_fuse, rank_align=args.rank_align)
            for (layer, _, raw), mod in zip(part, mods):
                father, name = index.where[raw]
                setattr(father, name, mod)
                done += 1
                if final and mod is not raw:
                    # Upstream moves the replaced layer to the CPU (:127) to free GPU memory; nothing reads it afterwards
                    # (module_dict / linear_info die with the function).  Here the index simply forgets the layer: the raw
                    # weight is released as soon as nothing else holds it (a weight tied to another module, OPT's
                    # lm_head <-> embed_tokens, stays where it is -- upstream's `.to("cpu")` would drag the embedding
                    # along on a GPU run), without 13 GB of pageable device-to-host copies for a 7B model.
                    # ASVD_B200_RAW_TO_CPU=1 restores upstream's literal behaviour for untied weights.
                    if os.environ.get("ASVD_B200_RAW_TO_CPU") == "1" and index.uses[id(raw.weight)] <= 1:
                        raw.to("cpu")
                    index.by_name.pop(layer, None)
                    index.where.pop(raw, None)
            del mods, part
    clear_cache()
    return done


def search_allocation(model, sensitivity_dict, calib_loader, args, index=None):
    """The search part (binary_search.py:29-110).  Returns ({layer: ratio}, default_ratio)."""
    index = index or LinearIndex(model)
    by_name = index.by_name
    if args.compress_kv_cache:
        ratio_target = args.kv_cache_ratio_target
        sensitivity_dict = {k: v for k, v in sensitivity_dict.items() if "k_proj" in k or "v_proj" in k}
        assert args.ppl_target < 0, "ppl_target is not supported when compressing kv_cache"
        default_ratio = 2
    else:
        ratio_target = args.param_ratio_target
        default_ratio = 1
    print(f"=== {'compress kv_cache' if args.compress_kv_cache else 'compress weight'} target: "
          f"ppl={args.ppl_target}, ratio_target={ratio_target} ===")
    flat = []
    for layer, table in sensitivity_dict.items():
        for ratio, ppl in table.items():
            if not args.compress_kv_cache and ratio >= 1:
                continue
            flat.append((layer, ratio, ppl))
    flat = sorted(flat, key=lambda t: -t[2])
    low, high, mid = 0, len(flat) - 1, None
    assert args.ppl_target > 0 or ratio_target > 0
    input_ids = torch.cat([b["input_ids"] for b in calib_loader], 0)
    while low < high:
        mid = (low + high) // 2
        chosen = _ratios_after_cut(flat, mid, sensitivity_dict.keys(), default_ratio)
        tot = comp = 0
        if args.ppl_target > 0:
            assert not args.compress_kv_cache, "ppl_target is not supported when compressing kv_cache now"
            # every layer of the table is decomposed at its current ratio -- including ratio 1, which upstream's
            # from_linear turns into rank m*n // (m+n) (:64-76) -- always from the RAW layer
            _install(index, chosen, default_ratio, args, final=False, decompose_default=True)
            for layer, ratio in chosen.items():
                numel = by_name[layer].weight.numel()
                tot += numel
                comp += numel * ratio
            ppl = evaluate_perplexity(model, input_ids, args.n_calib_samples)
            print(f"low={low} mid={mid}, high={high}, ppl={ppl}, param_ratio={comp / tot}")
            if ppl < args.ppl_target:
                high = mid
            else:
                low = mid + 1
        else:
            for layer, ratio in chosen.items():
                numel = by_name[layer].weight.numel()
                tot += numel
                comp += numel * ratio
            now_ratio = comp / tot
            if args.compress_kv_cache:
                now_ratio /= 2          # param ratio counts ALinear + BLinear; the rank ratio is half of it
            print(f"low={low} mid={mid}, high={high}, now_ratio={now_ratio}, params=({comp}/{tot})")
            if now_ratio > ratio_target:
                high = mid
            else:
                low = mid + 1
    print("=== Searching done, decomposing layers... ===")
    return _ratios_after_cut(flat, mid, sensitivity_dict.keys(), default_ratio), default_ratio   # stale mid (:106)


def decompose_layers(model, chosen, default_ratio, args, layer_filter=None, batch_limit_bytes=16 << 30, index=None):
    """The final pass (binary_search.py:112-128), batching same-shape layers per kernel call.
    Returns the number of layers replaced.  `index` must be the LinearIndex captured before a ppl-target search."""
    return _install(index or LinearIndex(model), chosen, default_ratio, args, layer_filter=layer_filter,
                    batch_limit_bytes=batch_limit_bytes, final=True)


def binary_search_truncation_rank(model, sensitivity_dict, calib_loader, args):
    index = LinearIndex(model)
    chosen, default_ratio = search_allocation(model, sensitivity_dict, calib_loader, args, index=index)
    st = time.time()
    decompose_layers(model, chosen, default_ratio, args, index=index)
    ed = time.time()
    print(f"decompose time: {ed - st}")
