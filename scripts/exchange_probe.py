"""Where does the time of sharding.broadcast_factors go?  (2+ ranks, NCCL)  torchrun ... scripts/exchange_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def sync(): torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
nbytes = int(float(os.environ.get("PROBE_GB", "1.5")) * 1e9)
t = torch.ones(8, device=dev); dist.all_reduce(t); sync()
for rep in range(3):
    out = {}
    t0 = time.perf_counter(); buf = torch.empty(nbytes, dtype=torch.uint8, device=dev); torch.cuda.synchronize(); out["alloc_send"] = time.perf_counter() - t0
    t0 = time.perf_counter(); books = [None] * world; dist.all_gather_object(books, {"n": nbytes, "layers": [{"layer": "x" * 40, "o": [1, 2]}] * 30}); out["all_gather_object"] = time.perf_counter() - t0
    recv = []
    t0 = time.perf_counter()
    for src in range(world):
        if src != rank: recv.append(torch.empty(nbytes, dtype=torch.uint8, device=dev))
    torch.cuda.synchronize(); out["alloc_recv"] = time.perf_counter() - t0
    sync()
    t0 = time.perf_counter(); k = 0
    for src in range(world):
        data = buf if src == rank else recv[k]
        if src != rank: k += 1
        dist.broadcast(data, src=src)
    torch.cuda.synchronize(); out["broadcasts"] = time.perf_counter() - t0
    sync()
    big = torch.empty(world * nbytes, dtype=torch.uint8, device=dev); torch.cuda.synchronize()
    t0 = time.perf_counter(); dist.all_gather_into_tensor(big, buf); torch.cuda.synchronize(); out["all_gather_into_tensor"] = time.perf_counter() - t0
    if rank == 0: print(rep, {k: round(v * 1e3, 1) for k, v in out.items()}, "ms;", round(nbytes * (world - 1) / 1e9 / out["broadcasts"], 1), "GB/s received via broadcasts", flush=True)
    del buf, recv, big
dist.destroy_process_group()
