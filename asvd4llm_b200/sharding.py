"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box; gloo in the CPU
tests), layers sharded at the only granularity the path offers — one nn.Linear = one independent unit
(SURVEY.md §8e).  No collective sits on a kernel's critical path: ranks factorise disjoint layers and exchange
results once per phase.

  calibration : ranks may shard the calibration samples; [n] accumulators are all-reduced (SUM / MAX)
  sensitivity : ranks sweep disjoint layers (calib_sensitivity_ppl(layer_filter=...)); tables are all-gathered
  final pass  : ranks decompose disjoint layers; each layer's factors are broadcast from its owner
"""
from __future__ import annotations

from typing import Dict, Iterable, List

import torch
import torch.distributed as dist
import torch.nn as nn

from .sensitivity import enumerate_linears


def layer_cost(out_features: int, in_features: int) -> int:
    """Relative time of one factorisation, calibrated on the measured shapes: a square of nv vectors
    costs nv^3; of that 0.68 are the streaming passes, which grow with the vector length, and 0.32 the inner solve, which
    does not.  Rectangles of 2:1 and flatter go through the Gram pre-conditioner (a square problem plus two sweeps on the
    long vectors: 1.5-1.8 x the square at 2.7:1, not 2.7 x); vectors longer than the pre-conditioner takes (16 Ki) run the
    direct path, alone in their batch.  (m * n * min(m, n) -- the streaming work alone -- put too few rectangles on a
    rank: slowest / fastest rank 1.80 / 1.53 s on Llama-2-7B at eight ranks.)"""
    nv, ln = min(out_features, in_features), max(out_features, in_features)
    ratio = ln / max(nv, 1)
    if ratio < 2.0 or nv < 1024:
        factor = 0.68 * ratio + 0.32
    elif ln <= 16384:
        factor = 1.0 + 0.25 * ratio
    else:
        factor = 1.4 * (0.68 * ratio + 0.32)
    return int(nv * nv * nv * factor)


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
        for layer, row in part.items():
            merged.setdefault(layer, {}).update(row)
    # rows sharded by (layer, ratio) unit arrive in pieces: restore the ratio order of the sweep (ascending), which is
    # the insertion order upstream's table has and binary_search.py:41-49 iterates
    return {full: dict(sorted(merged[full].items())) for _, _, full, _ in enumerate_linears(model) if full in merged}


def _pack(tensors):
    """One flat uint8 buffer holding `tensors` back to back, each at a 256-byte aligned offset.  Returns (buffer, offsets)."""
    offs, total = [], 0
    for t in tensors:
        offs.append(total)
        total += (t.numel() * t.element_size() + 255) // 256 * 256
    dev = tensors[0].device if tensors else torch.device("cpu")
    buf = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    for t, o in zip(tensors, offs):
        n = t.numel() * t.element_size()
        buf[o:o + n].copy_(t.contiguous().view(-1).view(torch.uint8))
    return buf, offs


def broadcast_factors(model: nn.Module, owners: Dict[str, int], replaced: Iterable[str], device=None) -> Dict[str, float]:
    """After a sharded final pass every rank installs every decomposed layer.  One exchange per OWNER, not per layer:
    each rank packs the factors of all the layers it decomposed (ALinear.weight, BLinear.weight, 256-byte aligned) into
    one flat buffer; a single all_gather_object carries the directory (layer, rank, dtype, offsets), then one
    dist.broadcast per rank moves its buffer (Llama-2-7B at 0.9: 11.9 GB in `world` collectives instead of 450
    latency-bound ones).  Receivers install VIEWS into the received buffer -- no second copy.  A layer whose
    factorisation failed on its owner (upstream's fallback nn.Linear, svd_linear.py:66-68,80-98) travels as weight + bias.
    Returns {"bytes": total bytes received, "collectives": number of data collectives}."""
    from .modules.svd_linear import SVDLinear
    rank, world = _world()
    stats = {"bytes": 0, "collectives": 0}
    if world == 1:
        return stats
    by_name = dict(model.named_modules())
    where = {full: (father, name) for father, name, full, _ in enumerate_linears(model)}
    mine = [full for full in replaced if owners[full] == rank]
    tensors, directory = [], []
    for full in mine:
        mod = by_name[full]
        if isinstance(mod, SVDLinear):
            A, B = mod.ALinear.weight.data, mod.BLinear.weight.data
            directory.append({"layer": full, "kind": "svd", "dtype": A.dtype, "A": tuple(A.shape), "B": tuple(B.shape), "slots": 2})
            tensors += [A, B]
        else:
            w = mod.weight.data
            entry = {"layer": full, "kind": "linear", "dtype": w.dtype, "W": tuple(w.shape), "slots": 1, "bias": mod.bias is not None}
            tensors.append(w)
            if mod.bias is not None:
                tensors.append(mod.bias.data)
                entry["slots"] = 2
            directory.append(entry)
    if device is None:
        device = tensors[0].device if tensors else next(model.parameters()).device
        if dist.get_backend() == "nccl" and device.type != "cuda":
            device = torch.device("cuda", torch.cuda.current_device())
    tensors = [t.to(device) for t in tensors]
    buf, offs = _pack(tensors) if tensors else (torch.empty(1, dtype=torch.uint8, device=device), [])
    k = 0
    for entry in directory:
        entry["offsets"] = offs[k:k + entry["slots"]]
        k += entry["slots"]
    books = [None] * world
    dist.all_gather_object(books, {"nbytes": int(buf.numel()), "layers": directory})
    for src in range(world):
        book = books[src]
        if not book["layers"]:
            continue
        data = buf if src == rank else torch.empty(book["nbytes"], dtype=torch.uint8, device=device)
        dist.broadcast(data, src=src)
        stats["collectives"] += 1
        if src == rank:
            continue
        stats["bytes"] += book["nbytes"]
        for entry in book["layers"]:
            full = entry["layer"]
            dt = entry["dtype"]

            def view(off, shape):
                n = 1
                for d in shape:
                    n *= d
                return data[off:off + n * dt.itemsize].view(dt).view(*shape)

            cur = by_name[full]
            father, name = where[full]
            tgt = (cur.ALinear.weight if isinstance(cur, SVDLinear) else cur.weight).device
            if entry["kind"] == "svd":
                A, B = view(entry["offsets"][0], entry["A"]), view(entry["offsets"][1], entry["B"])
                if tgt != A.device:
                    A, B = A.to(tgt), B.to(tgt)
                bias = (cur.ALinear.bias if isinstance(cur, SVDLinear) else cur.bias)
                setattr(father, name, SVDLinear._from_factors(A, B, None if bias is None else bias.data))
            else:
                W = view(entry["offsets"][0], entry["W"]).to(tgt)
                lin = nn.Linear(W.shape[1], W.shape[0], bias=entry["bias"], device="meta")
                lin.weight = nn.Parameter(W)
                if entry["bias"]:
                    lin.bias = nn.Parameter(view(entry["offsets"][1], (W.shape[0],)).to(tgt))
                setattr(father, name, lin)
    return stats


def decompose_sharded(model: nn.Module, chosen: Dict[str, float], default_ratio, args, index=None) -> Dict[str, float]:
    """binary_search.py:112-128 with the layers split over the ranks by LPT (cost model layer_cost), followed by the
    factor exchange.  Returns the exchange statistics plus the seconds spent in each part (device-synchronised): this rank's own
    decomposition, its wait for the slowest owner, the exchange proper."""
    import time
    from .binary_search import decompose_layers, LinearIndex
    rank, world = _world()
    index = index or LinearIndex(model)
    owners = owner_map_from_index(index, world)
    sync = torch.cuda.synchronize if torch.cuda.is_available() else (lambda: None)
    sync(); t0 = time.perf_counter()
    decompose_layers(model, chosen, default_ratio, args, layer_filter=lambda full: owners[full] == rank, index=index)
    sync(); t1 = time.perf_counter()
    if world > 1:
        dist.barrier()                   # so that a rank's wait for the slowest owner is not booked as exchange time
        sync()
    t2 = time.perf_counter()
    replaced = [full for full, ratio in chosen.items() if ratio != default_ratio]
    stats = broadcast_factors(model, owners, replaced)
    sync(); t3 = time.perf_counter()
    stats.update(decompose_s=t1 - t0, imbalance_wait_s=t2 - t1, exchange_s=t3 - t2)
    return stats


def owner_map_from_index(index, world_size: int) -> Dict[str, int]:
    """LPT owners computed from the RAW layers captured before any replacement (same result as owner_map on an
    untouched model)."""
    name_of = {mod: name for name, mod in index.by_name.items()}
    costs = {name_of[lin]: layer_cost(lin.out_features, lin.in_features) for lin in index.where}
    shards = lpt_partition(costs, world_size)
    return {name: r for r, names in enumerate(shards) for name in names}
