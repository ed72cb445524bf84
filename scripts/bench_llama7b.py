"""BASELINE config 3: factorise every linear of a Llama-2-7B-shaped model at param_ratio 0.9 (the `decompose time`
loop of binary_search.py:112-128), layers sharded over the ranks by LPT.  Random-init weights of the real shapes.
  python scripts/bench_llama7b.py            (1 GPU)      torchrun --nproc-per-node N scripts/bench_llama7b.py
Prints one JSON line (rank 0) and spot-checks one weight of every shape against torch.linalg.svdvals on the GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvd4llm_b200 import _lib, sharding

def layers(n_blocks=32, hidden=4096, inter=11008, vocab=32000):
    out = []
    for l in range(n_blocks):
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            out.append((f"model.layers.{l}.self_attn.{nm}", hidden, hidden))
        out.append((f"model.layers.{l}.mlp.gate_proj", inter, hidden))
        out.append((f"model.layers.{l}.mlp.up_proj", inter, hidden))
        out.append((f"model.layers.{l}.mlp.down_proj", hidden, inter))
    out.append(("lm_head", vocab, hidden))
    return out

def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_blocks = int(os.environ.get("LLAMA_BLOCKS", 32))
    L = layers(n_blocks)
    costs = {name: sharding.layer_cost(m, n) for name, m, n in L}
    mine = set(sharding.lpt_partition(costs, world)[rank])
    shapes = {}
    for i, (name, m, n) in enumerate(L):
        if name in mine:
            shapes.setdefault((m, n), []).append((i, name))
    _lib.load()
    # warm-up (library load, attribute set-up) on a small problem
    _lib.scaled_svd([torch.randn(256, 256, device=dev).half()], [None])
    torch.cuda.synchronize()
    if dist is not None: dist.barrier()
    t0 = time.perf_counter()
    per_shape, checks, done, sweeps_seen = {}, [], 0, {}
    for (m, n), items in shapes.items():
        cap = _lib.suggest_batch(m, n, dev)
        ts = time.perf_counter()
        r = _lib.rank_for_ratio(m, n, 0.9, 1)
        for j in range(0, len(items), cap):
            part = items[j:j + cap]
            Ws, Ss = [], []
            for idx, name in part:
                g = torch.Generator(device=dev).manual_seed(233 + idx)
                Ws.append((torch.randn(m, n, device=dev, generator=g) * 0.02).half())
                sdm = torch.exp(torch.randn(n, device=dev, generator=g)).half()
                Ss.append(_lib.scaling_vector(sdm, None, 0.5, n, dev))
            fact = _lib.scaled_svd(Ws, Ss)
            outs = [fact.extract(r, "UV", torch.float16, b) for b in range(len(part))]
            done += len(part)
            if j == 0:
                sweeps_seen[(m, n)] = list(fact.sweeps)
            del fact, outs, Ws, Ss
        torch.cuda.synchronize()
        per_shape[f"{m}x{n}"] = {"count": len(items), "seconds": time.perf_counter() - ts}
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # untimed verification: the first weight of every shape this rank owns, against an fp64 SVD (cuSOLVER gesvd)
    if os.environ.get("LLAMA_VERIFY", "1") == "1":
        for (m, n), items in shapes.items():
            idx, name = items[0]
            g = torch.Generator(device=dev).manual_seed(233 + idx)
            W = (torch.randn(m, n, device=dev, generator=g) * 0.02).half()
            sdm = torch.exp(torch.randn(n, device=dev, generator=g)).half()
            sc = _lib.scaling_vector(sdm, None, 0.5, n, dev)
            r = _lib.rank_for_ratio(m, n, 0.9, 1)
            fact = _lib.scaled_svd([W], [sc])
            ref = torch.linalg.svdvals(W.double() * sc.double(), driver="gesvd")
            sig = fact.sigma(0).double()
            rel = ((sig[:r] - ref[:r]).abs() / ref[:r]).max().item()
            A, B = fact.extract(r, "UV", torch.float32, 0)
            An = A / A.norm(dim=0, keepdim=True)
            AB = A @ B
            proj = ((An @ (An.t() @ W.float())) - AB).norm().item() / AB.norm().item()
            kept = (((AB * sc) ** 2).sum().item(), (ref[:r] ** 2).sum().item())
            checks.append({"shape": [m, n], "rank": r, "sigma_rel_err_vs_fp64": rel, "projector_residual": proj,
                           "kept_energy_rel_err": abs(kept[0] - kept[1]) / kept[1], "sweeps": sweeps_seen.get((m, n))})
            del fact, A, B, AB, An
    if dist is not None:
        dt_local = dt
        t = torch.tensor([dt], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        gathered = [None] * world
        dist.all_gather_object(gathered, {"rank": rank, "matrices": done, "seconds": dt_local, "per_shape": per_shape, "checks": checks})
    else:
        gathered = [{"rank": 0, "matrices": done, "seconds": dt, "per_shape": per_shape, "checks": checks}]
    if rank == 0:
        print(json.dumps({"workload": f"Llama-2-7B shapes ({n_blocks} blocks + lm_head), every linear at param_ratio 0.9, fp16 weights, alpha 0.5",
                          "n_gpus": world, "matrices": len(L), "seconds": dt, "matrices_per_s": len(L) / dt, "ranks": gathered}))
    if dist is not None: dist.destroy_process_group()

if __name__ == "__main__":
    main()
