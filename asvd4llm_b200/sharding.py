"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box; gloo in the CPU
tests), layers sharded at the only granularity the path offers — one nn.Linear = one independent unit
(SURVEY.md §8e).  No collective sits on a kernel's critical path: ranks factorise disjoint layers and exchange
results once per phase.

  calibration : ranks may shard the calibration samples; [n] accumulators are all-reduced (SUM / MAX)
  sensitivity : ranks sweep disjoint layers (calib_sensitivity_ppl(layer_filter=...)); tables are all-gathered
  final pass  : ranks decompose disjoint layers; each layer's factors are broadcast from its owner
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from .sensitivity import enumerate_linears


def layer_cost(out_features: int, in_features: int) -> int:
    """Work of one block-Jacobi SVD ~ max(m,n) * min(m,n)^2 (per sweep)."""
    return out_features * in_features * min(out_features, in_features)


def lpt_partition(costs: Dict[str, int], world_size: int) -> List[List[str]]:
    """Greedy longest-processing-time assignment; deterministic (ties broken by insertion order)."""
    order = sorted(range(len(costs)), key=lambda i: (-list(costs.values())[i], i))
    names = list(costs.keys())
    loads = [0] * world_size
    shards: List[List[str]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(names[i])
        loads[r] += costs[names[i]]
    return shards


def model_layer_costs(model: nn.Module) -> Dict[str, int]:
    return {full: layer_cost(lin.out_features, lin.in_features) for _, _, full, lin in enumerate_linears(model)}


def owner_map(model: nn.Module, world_size: int) -> Dict[str, int]:
    shards = lpt_partition(model_layer_costs(model), world_size)
    return {name: r for r, names in enumerate(shards) for name in names}


def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def allreduce_calibration(model: nn.Module, method: str) -> None:
    """Combine per-rank scaling_diag_matrix accumulators when the calibration samples were sharded."""
    rank, world = _world()
    if world == 1:
        return
    op = dist.ReduceOp.SUM if "abs_mean" in method else dist.ReduceOp.MAX
    for _, mod in model.named_modules():
        if isinstance(mod, nn.Linear) and torch.is_tensor(getattr(mod, "scaling_diag_matrix", None)):
            t = mod.scaling_diag_matrix
            buf = t.float() if t.dtype != torch.float32 else t.clone()
            dist.all_reduce(buf, op=op)
            mod.scaling_diag_matrix = buf.to(t.dtype)


def gather_sensitivity(model: nn.Module, shard: Dict[str, Dict[float, float]]) -> Dict[str, Dict[float, float]]:
    """All-gather the per-rank tables and restore upstream's sweep order (dict order feeds the stable sort of
    binary_search.py:49, so it is part of the contract)."""
    rank, world = _world()
    parts = [shard]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, shard)
    merged = {}
    for part in parts:
        merged.update(part)
    return {full: merged[full] for _, _, full, _ in enumerate_linears(model) if full in merged}


def broadcast_factors(model: nn.Module, owners: Dict[str, int], replaced: Iterable[str]) -> None:
    """After a sharded final pass every rank installs every decomposed layer: the owner broadcasts
    (rank r, ALinear.weight, BLinear.weight); bias tensors are already replicated.  A layer whose factorisation
    failed on its owner is replicated as the plain nn.Linear the owner ended up with."""
    from .modules.svd_linear import SVDLinear
    rank, world = _world()
    if world == 1:
        return
    by_name = dict(model.named_modules())
    where = {full: (father, name) for father, name, full, _ in enumerate_linears(model)}
    for full in replaced:
        src = owners[full]
        mod = by_name[full]
        is_svd = isinstance(mod, SVDLinear)
        dev = (mod.ALinear.weight if is_svd else mod.weight).device
        # rank -1: the owner's factorisation took upstream's failure path (svd_linear.py:66-68,80-98) and left a plain
        # nn.Linear (fresh, or the raw layer with ASVD_B200_KEEP_RAW_ON_FAILURE=1); its weight and bias are
        # replicated instead, so that every rank still holds the same model
        meta = torch.tensor([(mod.truncation_rank if is_svd else -1) if rank == src else 0], dtype=torch.int64, device=dev)
        dist.broadcast(meta, src=src)
        r = int(meta.item())
        if r < 0:
            dist.broadcast(mod.weight.data, src=src)
            if mod.bias is None:
                # upstream's fallback layer always has a bias (nn.Linear default); receivers need the tensor first
                has = torch.tensor([0], dtype=torch.int64, device=dev)
            else:
                has = torch.tensor([1], dtype=torch.int64, device=dev)
            dist.broadcast(has, src=src)
            if int(has.item()):
                if mod.bias is None:
                    mod.bias = nn.Parameter(torch.zeros(mod.out_features, dtype=mod.weight.dtype, device=dev))
                dist.broadcast(mod.bias.data, src=src)
            elif mod.bias is not None:
                mod.bias = None
            continue
        if rank == src:
            A, B = mod.ALinear.weight.data, mod.BLinear.weight.data
        else:
            w = mod.weight.data
            A = torch.empty(w.shape[0], r, dtype=w.dtype, device=w.device)
            B = torch.empty(r, w.shape[1], dtype=w.dtype, device=w.device)
        dist.broadcast(A, src=src)
        dist.broadcast(B, src=src)
        if rank != src:
            bias = mod.bias.data if mod.bias is not None else None
            father, name = where[full]
            setattr(father, name, SVDLinear._from_factors(A, B, bias))


def decompose_sharded(model: nn.Module, chosen: Dict[str, float], default_ratio, args) -> None:
    """binary_search.py:112-128 with the layers split over the ranks by LPT."""
    from .binary_search import decompose_layers
    rank, world = _world()
    owners = owner_map(model, world)
    decompose_layers(model, chosen, default_ratio, args, layer_filter=lambda full: owners[full] == rank)
    replaced = [full for full, ratio in chosen.items() if ratio != default_ratio]
    broadcast_factors(model, owners, replaced)
