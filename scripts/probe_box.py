"""One-off box probe: host cores, GPU properties, vendor-library timings (context only)."""
import json, os, time, torch
out = {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()}
try:
    out["cpu_model"] = [l for l in open("/proc/cpuinfo") if "model name" in l][0].split(":")[1].strip()
    out["affinity"] = len(os.sched_getaffinity(0))
except Exception as e:
    out["cpu_model"] = str(e)
p = torch.cuda.get_device_properties(0)
out["gpu"] = {k: getattr(p, k) for k in ("name", "total_memory", "multi_processor_count", "L2_cache_size",
      "shared_memory_per_block_optin", "shared_memory_per_multiprocessor", "major", "minor") if hasattr(p, k)}
torch.manual_seed(233)
W = (torch.randn(4096, 4096) * 0.02).half()
s = torch.exp(torch.randn(4096))
Ws = W.float() * (s ** 0.5 + 1e-6)
def t_cpu(f, n=1):
    f(); ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return sorted(ts)[len(ts) // 2]
out["cpu_linalg_svd_4096_s"] = t_cpu(lambda: torch.linalg.svd(Ws, full_matrices=False))
out["cpu_svd_lowrank_4096_q1843_s"] = t_cpu(lambda: torch.svd_lowrank(Ws, q=1843))
Wg = Ws.cuda()
def t_gpu(f, n=2):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(True); b = torch.cuda.Event(True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / 1e3)
    return sorted(ts)[len(ts) // 2]
for drv in ("gesvdj", "gesvd", "gesvda"):
    try:
        out[f"gpu_linalg_svd_4096_{drv}_s"] = t_gpu(lambda: torch.linalg.svd(Wg, full_matrices=False, driver=drv), 1)
    except Exception as e:
        out[f"gpu_linalg_svd_4096_{drv}_s"] = str(e)[:100]
out["gpu_svd_lowrank_4096_q1843_s"] = t_gpu(lambda: torch.svd_lowrank(Wg, q=1843), 1)
out["gpu_eigh_4096_s"] = t_gpu(lambda: torch.linalg.eigh(Wg.T @ Wg), 1)
x = torch.randn(8192, 4096, device="cuda", dtype=torch.float32)
out["gpu_sgemm_fp32_tflops"] = 2 * 8192 * 4096 * 4096 / t_gpu(lambda: x @ Wg, 5) / 1e12
torch.backends.cuda.matmul.allow_tf32 = True
out["gpu_sgemm_tf32_tflops"] = 2 * 8192 * 4096 * 4096 / t_gpu(lambda: x @ Wg, 5) / 1e12
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_box.json", "w"), indent=1)
print(json.dumps(out, indent=1))
