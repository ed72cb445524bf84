#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs.

  python bench.py --gpus N --steps K --warmup W                     (our arm; torchrun launches N>1, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W    (the reference's CPU algorithm, rank 0 only)
  python bench.py --workload forward | llama7b ...                  (BASELINE configs 4 and 3 as lines of their own)

Default workload `svd` (configs[1], the configuration the metric is quoted on): 4096x4096 fp16 weights, activation
scale s = sdm**0.5 + 1e-6, param_ratio 0.9 -> rank 1843, sigma_fuse "UV", factors written in fp16.  A step = one batch
of same-shape weights per GPU (the batch the product itself picks, _lib.suggest_batch) through the whole hot path
(scaling vector -> scaled SVD -> truncation / un-scaling / sigma fusion / cast).  `value` times it with everything
resident in HBM; `e2e` goes through the public module API (SVDLinear.from_linear semantics, batched) with the weights in
pinned HOST memory and the factors copied back to the host, copies inside the timed region.  Each rank works on its
own weights (no data-path collective): scaling is "weak".

The default line also carries `extras`: config 4 (SVDLinear.forward at B=32, L=2048, d=4096, r in {256, 512, 1024, 1843},
tensor-pipe roofline, cuBLAS pair of the same box beside it) and config 3 (all 225 linears of a Llama-2-7B-shaped model
at 0.9 through sharding.decompose_sharded, factor exchange included; strong scaling over the ranks) so that the
driver-run line holds them too; `--no-extras` skips them.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, N_IN, RATIO, ALPHA = 4096, 4096, 0.9, 0.5
METRIC = "weight-matrices factorised/sec (full-model ASVD wall-clock) at 1/2/4/8 GPU"
UNIT = "matrices/s"
WORKLOAD = "single 4096x4096 fp16 weight: activation-scaled SVD + rank-1843 truncation (param_ratio 0.9, alpha 0.5, sigma_fuse UV)"
FWD_WORKLOAD = "SVDLinear.forward: B=32 L=2048 d=4096 fp16, rank r in {256, 512, 1024, 1843}"
FWD_CONFIG = {"workload": FWD_WORKLOAD, "l2": "x and y are 512 MB each: larger than the 126 MB L2"}
LLAMA_WORKLOAD = "Llama-2-7B shapes (225 linears, random init) full ASVD factorisation at param_ratio 0.9, layers sharded over the ranks, factors exchanged"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p.get("hbm_gbs", 6650.0), "bf16_tflops": p.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", 1400.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi polling in a child process, started BEFORE the warm-up (its NVML start-up takes most of a second and was
    seen to stall a launch-heavy step when it fell inside the timed region); only the samples whose timestamps lie between
    begin() and end() are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.index = None, index
        self.t0 = self.t1 = None
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}

    def start(self):
        if os.environ.get("ASVD_BENCH_NO_SAMPLER") == "1":     # experiments only: is the sampler itself perturbing the run?
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    # context-manager form: the whole `with` body is the sampled region (workloads without a separate warm-up phase)
    def __enter__(self):
        self.start()
        time.sleep(1.0)                     # let NVML come up before the region starts
        self.begin()
        return self

    def __exit__(self, *a):
        self.end()
        self.stop()

    def stop(self):
        if self.proc is None:
            return self.result
        if self.t1 is None:
            self.end()
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return self.result
        import datetime
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.t0 is not None and not (self.t0 - 0.05 <= ts <= self.t1 + 0.05):
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            self.result = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                           "samples": len(sm)}
        return self.result


# ---------------------------------------------------------------------------------------------------------------- CPU arm
_UPSTREAM = "unset"


def upstream_module():
    """upstream's own modules/svd_linear.py staged in oracle/_ref (oracle/make_ref.py), or None."""
    global _UPSTREAM
    if _UPSTREAM == "unset":
        try:
            from oracle import make_ref
            _UPSTREAM = make_ref.load_upstream_svd_linear()
        except Exception:      # noqa
            _UPSTREAM = None
    return _UPSTREAM


def reference_kind():
    return "reference" if upstream_module() is not None else "port"


def reference_step(lin, O, method="lowrank"):
    """One unit of the reference's CPU path (modules/svd_linear.py:26-103: scale, SVD, un-scale, fuse, cast).
    method="lowrank" is what upstream ships (torch.svd_lowrank, q = rank, niter 2): upstream's OWN file when oracle/_ref
    holds it (kind "reference"), else the oracle's restatement (kind "port"); "exact" is torch.linalg.svd, the oracle
    north_star names (restatement only: upstream has no such code path)."""
    up = upstream_module()
    if method == "lowrank" and up is not None:
        return up.SVDLinear.from_linear(lin, RATIO, act_aware=True, alpha=ALPHA, sigma_fuse="UV")
    return O.from_linear(lin, RATIO, act_aware=True, alpha=ALPHA, sigma_fuse="UV", method=method)


def make_cpu_linear(seed, m=M, n=N_IN):
    import torch.nn as nn
    g = torch.Generator().manual_seed(seed)
    lin = nn.Linear(n, m, bias=False)
    lin.weight.data = (torch.randn(m, n, generator=g) * 0.02).half()
    lin.scaling_diag_matrix = torch.exp(torch.randn(n, generator=g)).half()
    return lin


def _host_threads() -> int:
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers, which would make
    the CPU arm single-threaded at N > 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def svd_config(batch, rank):
    """`config` of the svd workload: the SAME dict in both arms (the driver compares the two); everything that depends on
    the run (sweeps, per-step times, ...) goes to the line's `run` key instead."""
    return {"workload": WORKLOAD, "batch_per_step_per_gpu": batch, "rank": rank,
            "l2": f"inputs larger than L2: {batch} x 128 MB fp32 working set per step vs 126 MB L2; two input sets alternate",
            "parallelism": "independent ranks, disjoint weights, no data-path collective"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import asvd_oracle as O
    torch.set_num_threads(_host_threads())             # torchrun exports OMP_NUM_THREADS=1; this arm uses every host thread
    torch.manual_seed(233)
    cores = torch.get_num_threads()
    if args.workload == "forward":
        # upstream's module on the host CPU: two nn.Linear calls in fp32 (fp16 matmul is not a CPU path upstream runs)
        import torch.nn.functional as F
        g = torch.Generator().manual_seed(1)
        Mtok = 2048                                        # bounded sample: one sequence of the 32
        x = torch.randn(Mtok, 4096, generator=g) * 0.125
        ts_all = {}
        for r in (256, 512, 1024, 1843):
            Bw = torch.randn(r, 4096, generator=g) / 64; Aw = torch.randn(4096, r, generator=g) / r ** 0.5
            F.linear(F.linear(x, Bw), Aw)
            t0 = time.perf_counter()
            for _ in range(max(1, args.steps)):
                F.linear(F.linear(x, Bw), Aw)
            ts_all[r] = (time.perf_counter() - t0) / max(1, args.steps)
        tot = sum(ts_all.values())
        value = 4 * Mtok / tot
        _emit({"impl": "reference", "metric": "SVDLinear.forward throughput (mean over the four ranks)", "value": value, "unit": "tokens/s",
               "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot * 1e3, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": FWD_CONFIG,
               "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port",
                                "sample": "one 2048-token sequence per rank r, fp32 F.linear pair (upstream's forward) on the host"},
               "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    if args.workload == "llama7b":
        per_shape, total = llama_cpu_estimate(O)
        value = 225 / total
        _emit({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
               "warmup": 0, "ms_per_step": total * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": {"workload": LLAMA_WORKLOAD},
               "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": reference_kind(),
                                "sample": "one weight of each of the four shapes timed (svd_lowrank q=rank niter=2), multiplied by the shape's count: "
                                          + json.dumps(per_shape)},
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    lin = make_cpu_linear(233)
    for _ in range(max(1, min(args.warmup, 1))):       # one warm-up is enough for a CPU LAPACK path
        reference_step(lin, O)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_step(lin, O)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = f"{args.steps} x one 4096x4096 fp16 weight @0.9 per step (scale, svd_lowrank q=1843 niter=2, un-scale, fuse, cast)"
    _emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": svd_config(_default_batch_static(), 1843),
        "run": {"where": ("host CPU, upstream's own modules/svd_linear.py (oracle/_ref)" if reference_kind() == "reference"
                          else "host CPU, oracle port of the upstream algorithm")
                         + "; one weight per CPU step (a bounded sample of the GPU arm's batch)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": reference_kind(), "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def _default_batch_static():
    """suggest_batch(4096, 4096) on a 148-SM part, without touching a GPU (the CPU arm prints the same config keys)."""
    return 2 * 3 * 148 // 32


def llama_cpu_estimate(O):
    counts = {(4096, 4096): 128, (11008, 4096): 64, (4096, 11008): 32, (32000, 4096): 1}
    per_shape, total = {}, 0.0
    for (m, n), c in counts.items():
        lin = make_cpu_linear(7, m, n)
        t0 = time.perf_counter()
        reference_step(lin, O)
        dt = time.perf_counter() - t0
        per_shape[f"{m}x{n}"] = {"count": c, "s_per_matrix": round(dt, 3)}
        total += c * dt
    return per_shape, total


_JSON_FD = None


def _claim_stdout():
    """Rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL's version banner comes through C stdio
    whatever NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for the run and the JSON line goes to
    the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# ---------------------------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", 0))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.dist is None:
            return v
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def svd_workload(ctx):
    from asvd4llm_b200 import _lib
    from asvd4llm_b200.modules.svd_linear import from_linear_batch
    import torch.nn as nn
    args, dev, rank, world = ctx.args, ctx.dev, ctx.rank, ctx.world
    # the batch the product's own final pass uses for this shape (binary_search._install -> suggest_batch): two solve
    # waves of 3 x SMs block pairs (27 at 4096^2 on 148 SMs); --batch overrides for experiments
    B = args.batch if args.batch > 0 else _lib.suggest_batch(M, N_IN, dev)
    r = _lib.rank_for_ratio(M, N_IN, RATIO, 1)
    g = torch.Generator(device=dev).manual_seed(233 + rank)
    n_pool = 2                                            # alternate two input sets; the B x 128 MB fp32 working set
    pools = []                                            # per step is itself several times the 126 MB L2
    for _ in range(n_pool):
        Ws = [(torch.randn(M, N_IN, device=dev, generator=g) * 0.02).half() for _ in range(B)]
        sdm = [torch.exp(torch.randn(N_IN, device=dev, generator=g)).half() for _ in range(B)]
        pools.append((Ws, sdm))

    def device_step(i):
        Ws, sdm = pools[i % n_pool]
        scales = [_lib.scaling_vector(s, None, ALPHA, N_IN, dev) for s in sdm]
        fact = _lib.scaled_svd(Ws, scales)
        outs = [fact.extract(r, "UV", torch.float16, b) for b in range(B)]
        return fact, outs

    clocks = ClockSampler(ctx.local).start()             # NVML start-up happens during the warm-up, not the timed region
    # Every step drops the previous step's factorisation before it allocates its own workspace (6.9 GB at 27 weights), as a
    # pipeline that consumes the factors would: with two workspaces alive in turn the caching allocator sooner or later
    # splits the free one for the 15 MB factor tensors and has to cudaMalloc a fresh 4.6 GB block mid-run -- the "hiccup"
    # that kept landing on the second timed step (843 -> 900-1270 ms; profiles/r02_bench_hiccup.log).
    fact = outs = None
    for i in range(args.warmup):
        fact = outs = None
        fact, outs = device_step(i)
    ctx.barrier()
    # A freshly booted box has been seen to take one step 40-80 % longer than its neighbours about a second into the load
    # (sw_power_cap flagged, clocks back at maximum right after): keep warming up, untimed, until two consecutive steps
    # agree to 5 % (at most 8 extra steps), so that the event falls outside the timed region.
    # (at least 3 s of load -- the event comes about a second in -- at most 12 extra steps)
    extra, prev, t_load = 0, None, time.perf_counter()
    while extra < 12:
        t0 = time.perf_counter()
        fact = outs = None
        fact, outs = device_step(extra)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        extra += 1
        if prev is not None and abs(dt - prev) <= 0.05 * prev and time.perf_counter() - t_load >= 3.0:
            break
        prev = dt
    ctx.barrier()
    # exactly K steps between two barriers, one CUDA-event bracket; per-step events only to REPORT a hiccup (the number
    # is what the K steps took, whatever happened in them)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ctx.barrier()
    clocks.begin()
    e0.record()
    marks[0].record()
    for i in range(args.steps):
        fact = outs = None
        fact, outs = device_step(i)
        marks[i + 1].record()
    e1.record()
    ctx.barrier()
    clocks.end()
    clocks.stop()
    per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    sweeps = list(fact.sweeps)
    value = world * B * args.steps / (ms / 1e3)
    hiccup = None
    if args.steps >= 3 and max(per_step) > 1.2 * statistics.median(per_step):
        hiccup = {"slow_step_ms": round(max(per_step), 1), "median_step_ms": round(statistics.median(per_step), 1),
                  "note": "reported, not re-measured: value is what the K steps took"}

    # ---- e2e: public module API, host-resident weights, factors read back to the host
    lins = []
    for b in range(B):
        lin = nn.Linear(N_IN, M, bias=False)
        lin.weight.data = pools[0][0][b].cpu().pin_memory()
        lin.scaling_diag_matrix = pools[0][1][b].cpu().pin_memory()
        lins.append(lin)

    def e2e_step():
        mods = from_linear_batch(lins, [RATIO] * B, act_aware=True, alpha=ALPHA, sigma_fuse="UV")
        return sum(float(m.ALinear.weight.data[0, 0]) for m in mods)     # factors are host tensors here

    e2e_step()
    ctx.barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * n_e2e / e2e_s
    h2d = B * (M * N_IN * 2 + N_IN * 2)
    d2h = B * (r * (M + N_IN) * 2)

    # ---- roofline of the SVD "kernel" of SURVEY.md 8(d): the Jacobi round = one launch each of gram_tc_kernel, the inner
    # solve and update_tc_kernel.  ALGORITHMIC bytes per 8(d): the working set X [n', m'] fp32 streamed once in and once
    # out per round, 2 * 4 * m' * n' per matrix (V is recovered by a GEMM afterwards, not accumulated).  The implementation
    # reads X twice (Gram pass, update pass) and writes it once: reported separately as implementation_bytes.
    # Timed with per-class CUDA events (asvd_profile_*) over the first two sweeps, where every block pair is active.
    roofline = None
    if rank == 0:
        Ws, sdm = pools[0]
        scales = [_lib.scaling_vector(s, None, ALPHA, N_IN, dev) for s in sdm]
        _lib.profile_enable(True)
        before = _lib.profile_read()
        _lib.scaled_svd(Ws, scales, max_sweeps=2)
        torch.cuda.synchronize()
        after = _lib.profile_read()
        classes = {k: {"ms": after[k][0], "launches": after[k][1] - before[k][1]} for k in after if after[k][1] > before[k][1]}
        _lib.profile_enable(True)                       # resets the timers
        before = _lib.profile_read()
        _lib.scaled_svd(Ws, scales)
        torch.cuda.synchronize()
        after = _lib.profile_read()
        _lib.profile_enable(False)
        full_run = {k: round(after[k][0], 2) for k in after if after[k][1] > before[k][1]}
        nv = len_ = 4096
        pairs, JK = nv // 128, 128
        x_bytes = B * nv * len_ * 4
        g_bytes = r_bytes = B * pairs * JK * JK * 4
        aux_bytes = B * pairs * 67584                                   # per-pair rotation record: written by the G sweep, read by the replay
        impl = {"gram": x_bytes + g_bytes, "solve": g_bytes + r_bytes + 2 * aux_bytes, "update": 2 * x_bytes + r_bytes}
        pk = peaks()
        peak = pk["hbm_gbs"]
        traffic, tsrc = {}, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
            tsrc = traffic.get("_source", "profiles/traffic.json: one `ncu --set full` capture per kernel at a batch of 4, scaled by batch/4 "
                                          "(every kernel works one block pair per CTA); not a counter of this run")
        tb = float(traffic.get("_batch", 4))
        rounds = max(1, classes["gram"]["launches"])
        per_kernel, t_round = {}, 0.0
        for k in ("gram", "solve", "update"):
            us = classes[k]["ms"] / rounds * 1e3
            t_round += us
            tks = {"gram": ["gram_tc"], "solve": ["solve_tri_g", "solve_tri_r"], "update": ["update_tc"]}[k]
            tv = sum(traffic[t] for t in tks) if all(t in traffic for t in tks) else None
            per_kernel[k] = {"us_per_round": round(us, 1), "launches_per_round": round(classes[k]["launches"] / rounds, 2),
                             "implementation_bytes": impl[k], "implementation_GBps": round(impl[k] / us / 1e3, 1),
                             "frac_of_hbm_peak_implementation_bytes": round(impl[k] / us / 1e3 / peak, 3),
                             "dram_traffic_ncu_scaled": None if tv is None else tv * B / tb}
        alg_bytes = 2 * x_bytes                                       # 2 * 4 * m' * n' per matrix and round
        impl_bytes = sum(impl.values())
        tr = [per_kernel[k]["dram_traffic_ncu_scaled"] for k in ("gram", "solve", "update")]
        roofline = {"bound": "hbm", "kernel": "jacobi_round = gram_tc_kernel + inner solve + update_tc_kernel",
                    "achieved": alg_bytes / t_round / 1e3, "peak": peak, "unit": "GB/s", "frac": alg_bytes / t_round / 1e3 / peak,
                    "traffic": (sum(tr) if all(v is not None for v in tr) else None), "traffic_source": tsrc,
                    "peak_source": pk["source"] + ", hbm_gbs (burst copy)",
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "algorithmic_bytes_rule": "SURVEY.md 8(d): 2 * 4 * m' * n' per matrix per round (X once in, once out; V by GEMM)",
                    "implementation_bytes_per_launch": impl_bytes,
                    "implementation_frac": impl_bytes / t_round / 1e3 / peak,
                    "avg_launch_us": round(t_round, 1),
                    "solve_share_of_round": round(per_kernel["solve"]["us_per_round"] / t_round, 3),
                    "note": "the inner solve (solve_tri_g_kernel + solve_tri_r_kernel) is an on-chip Jacobi eigensolver bounded by its 127 "
                            "dependent rotation steps, not by HBM; gram/update are the streaming passes",
                    "per_kernel": per_kernel,
                    "class_ms_first_2_sweeps": {k: round(v["ms"], 3) for k, v in classes.items()},
                    "class_ms_full_factorisation": full_run}

    # ---- the same shape with a decaying spectrum (trained, activation-scaled weights look like this; the Gaussian input
    # above has a clustered Marchenko-Pastur bulk, the slowest case for a Jacobi method): sweeps and time, rank 0
    decaying = None
    if rank == 0:
        gd = torch.Generator(device=dev).manual_seed(99)
        Wd = []
        for _ in range(B):
            u = torch.randn(M, 64, device=dev, generator=gd); v = torch.randn(64, N_IN, device=dev, generator=gd)
            sv = torch.logspace(0, -2, 64, device=dev) * 8.0
            Wd.append(((torch.randn(M, N_IN, device=dev, generator=gd) + (u * sv) @ v) * 0.02).half())
        sc = [_lib.scaling_vector(s, None, ALPHA, N_IN, dev) for s in pools[0][1]]
        f2 = _lib.scaled_svd(Wd, sc); torch.cuda.synchronize()
        t0 = time.perf_counter(); f2 = _lib.scaled_svd(Wd, sc); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        decaying = {"input": "Gaussian bulk + 64 directions with singular values decaying over two decades", "sweeps": list(f2.sweeps),
                    "ms_per_matrix": round(dt * 1e3 / B, 2), "matrices_per_s": round(B / dt, 2)}
        del Wd, f2

    # ---- baselines on this box, rank 0, N=1 only: the reference's algorithms on the host cores (bounded sample) and on
    # the same GPU through torch (the kernel to beat is the vendor library on the same box)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = svd_baselines(dev, pools[0])

    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": svd_config(B, r),
        "run": dict(sweeps=sweeps, ranks=world,
                    batch_rule="_lib.suggest_batch(4096, 4096): the batch binary_search's final pass uses for this shape",
                    step_ms=[round(t, 1) for t in per_step], extra_untimed_warmup_steps=extra, hiccup=hiccup,
                    decaying_spectrum_input=decaying),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": n_e2e, "api": "asvd4llm_b200.modules.svd_linear.from_linear_batch (SVDLinear.from_linear semantics)"},
        "gpu_launches": int(launches),
        "clocks": clocks.result,
    }


def svd_baselines(dev, pool):
    from oracle import asvd_oracle as O
    torch.set_num_threads(_host_threads())
    torch.manual_seed(233)
    lin = make_cpu_linear(233)
    reference_step(lin, O)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); reference_step(lin, O); ts.append(time.perf_counter() - t0)
    t_low = statistics.median(ts)
    t0 = time.perf_counter(); reference_step(lin, O, "exact"); t_exact = time.perf_counter() - t0
    legs = {"cpu_svd_lowrank": {"s_per_matrix": round(t_low, 3), "matrices_per_s": round(1 / t_low, 3),
                                "what": "upstream's shipped call torch.svd_lowrank(q=1843, niter=2) + un-scale/fuse/cast, median of 3"},
            "cpu_linalg_svd": {"s_per_matrix": round(t_exact, 3), "matrices_per_s": round(1 / t_exact, 3),
                               "what": "torch.linalg.svd (north_star's oracle) + un-scale/fuse/cast, one run"}}
    # same GPU, vendor libraries through torch (not part of the product path)
    W = pool[0][0].float() * (pool[1][0].float() ** 0.5 + 1e-6)
    r = 1843
    for name, fn in (("cuda_svd_lowrank", lambda: torch.svd_lowrank(W, q=r)),
                     ("cuda_linalg_svd", lambda: torch.linalg.svd(W, full_matrices=False))):
        try:
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
            legs[name] = {"s_per_matrix": round(dt, 4), "matrices_per_s": round(1 / dt, 3),
                          "what": "torch." + ("svd_lowrank(q=1843, niter=2)" if "lowrank" in name else "linalg.svd(full_matrices=False)")
                                  + " on this GPU (cuSOLVER / cuBLAS), SVD only, one weight"}
        except Exception as e:      # noqa
            legs[name] = {"error": str(e)[:200]}
    return {"value": 1.0 / t_low, "unit": UNIT, "cores": torch.get_num_threads(), "kind": reference_kind(),
            "sample": "one 4096x4096 fp16 weight @0.9: 3 x upstream's svd_lowrank path (median) and 1 x torch.linalg.svd on the host cores; "
                      "same weight through torch on this GPU for context", "legs": legs}


def forward_workload(ctx, ranks=(256, 512, 1024, 1843), steps=None):
    """BASELINE config 4 through the module API (SVDLinear.forward).  Per rank r: CUDA-event time of K forwards of the
    [32, 2048, 4096] fp16 batch (x and y are 512 MB each: larger than L2), tensor-pipe roofline 2*M*r*(n+m) flop / t
    against the measured cuBLAS bf16 peak, and upstream's own module (two torch F.linear calls = cuBLAS) on the same box."""
    from asvd4llm_b200 import _lib
    from asvd4llm_b200.modules.svd_linear import SVDLinear
    import torch.nn.functional as F
    args, dev = ctx.args, ctx.dev
    K = steps or max(3, min(args.steps, 10))
    W = max(3, args.warmup)
    Mtok, n, m = 32 * 2048, 4096, 4096
    g = torch.Generator(device=dev).manual_seed(7 + ctx.rank)
    x = (torch.randn(32, 2048, n, device=dev, generator=g) * 0.125).half()
    pk = peaks()
    out = {}
    launches0 = _lib.launch_count()
    for r in ranks:
        Bw = (torch.randn(r, n, device=dev, generator=g) / n ** 0.5).half()
        Aw = (torch.randn(m, r, device=dev, generator=g) / r ** 0.5).half()
        mod = SVDLinear._from_factors(Aw, Bw, None)

        def timeit(fn):
            for _ in range(W):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / K * 1e3       # us

        with torch.no_grad():
            us = timeit(lambda: mod(x))
            us_ref = timeit(lambda: F.linear(F.linear(x, Bw), Aw))
        flops = 2.0 * Mtok * r * (n + m)
        tf = flops / us / 1e6
        out[str(r)] = {"us": round(us, 1), "tokens_per_s": round(Mtok / us * 1e6), "tflops": round(tf, 1),
                       "frac_of_bf16_burst_peak": round(tf / pk["bf16_tflops"], 3),
                       "frac_of_bf16_sustained_peak": round(tf / pk["bf16_tflops_sustained"], 3),
                       "kernel": "gemm_tn2_kernel x 2 (tcgen05 cta_group::2, 256-row tiles)",
                       "cublas_pair_us": round(us_ref, 1), "speedup_vs_cublas_pair": round(us_ref / us, 3)}
    return out, _lib.launch_count() - launches0, K, W


def build_llama_like(dev, n_blocks=32, hidden=4096, inter=11008, vocab=32000, seed=233):
    """A module tree with Llama-2-7B's linear shapes and names (random init, fp16), every rank the same values."""
    import torch.nn as nn

    def lin(m, n, g):
        l = nn.Linear(n, m, bias=False, device="meta")
        l.weight = nn.Parameter((torch.randn(m, n, device=dev, generator=g) * 0.02).half(), requires_grad=False)
        l.scaling_diag_matrix = torch.exp(torch.randn(n, device=dev, generator=g)).half()
        return l

    g = torch.Generator(device=dev).manual_seed(seed)
    model = nn.Module()
    model.model = nn.Module()
    layers = []
    for _ in range(n_blocks):
        blk = nn.Module()
        blk.self_attn = nn.Module()
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            setattr(blk.self_attn, nm, lin(hidden, hidden, g))
        blk.mlp = nn.Module()
        blk.mlp.gate_proj = lin(inter, hidden, g)
        blk.mlp.up_proj = lin(inter, hidden, g)
        blk.mlp.down_proj = lin(hidden, inter, g)
        layers.append(blk)
    model.model.layers = nn.ModuleList(layers)
    model.lm_head = lin(vocab, hidden, g)
    return model


def llama_workload(ctx, n_blocks=32):
    """BASELINE config 3 through the repo's API: sharding.decompose_sharded = binary_search's final pass with the layers
    split over the ranks by LPT + the per-owner factor exchange (NCCL broadcasts of packed buffers).  Strong scaling:
    the model is fixed, the ranks share it.  Time = max over ranks of (decompose + exchange), device-synchronised."""
    from asvd4llm_b200 import _lib, sharding
    from asvd4llm_b200.sensitivity import enumerate_linears
    from asvd4llm_b200.modules.svd_linear import SVDLinear
    dev = ctx.dev
    model = build_llama_like(dev, n_blocks=n_blocks)
    names = [full for _, _, full, _ in enumerate_linears(model)]
    chosen = {nm: RATIO for nm in names}
    ns = argparse.Namespace(alpha=ALPHA, act_aware=True, sigma_fuse="UV", rank_align=1)
    _lib.scaled_svd([torch.randn(256, 256, device=dev).half()], [None])        # library load / attribute set-up
    launches0 = _lib.launch_count()
    ctx.barrier()
    t0 = time.perf_counter()
    stats = sharding.decompose_sharded(model, chosen, 1, ns)
    ctx.barrier()
    total = ctx.max_over_ranks(time.perf_counter() - t0)
    dec = ctx.max_over_ranks(stats["decompose_s"])
    dec_min = -ctx.max_over_ranks(-stats["decompose_s"])
    exch = ctx.max_over_ranks(stats["exchange_s"])
    mods = [mod for _, mod in model.named_modules() if isinstance(mod, SVDLinear)]
    assert len(mods) == len(names), (len(mods), len(names))
    # checksum of every factor as installed on this rank: identical on every rank and for every world size
    chk = 0
    for mod in mods:
        for t in (mod.ALinear.weight.data, mod.BLinear.weight.data):
            chk = (chk * 1000003 + int(t.contiguous().view(torch.int16).to(torch.int64).sum().item())) % (1 << 61)
    if ctx.dist is not None:
        t = torch.tensor([chk], device=dev, dtype=torch.int64)
        lo, hi = t.clone(), t.clone()
        ctx.dist.all_reduce(lo, op=ctx.dist.ReduceOp.MIN); ctx.dist.all_reduce(hi, op=ctx.dist.ReduceOp.MAX)
        assert int(lo.item()) == int(hi.item()), "ranks hold different factors after the exchange"
    return {"linears": len(names), "seconds": round(total, 3), "matrices_per_s": round(len(names) / total, 2),
            "decompose_s": round(dec, 3), "decompose_s_fastest_rank": round(dec_min, 3), "exchange_s": round(exch, 3),
            "exchange_GB_received_per_rank": round(stats["bytes"] / 1e9, 3), "exchange_collectives": stats["collectives"],
            "exchange_GBps_per_rank": (round(stats["bytes"] / 1e9 / exch, 1) if exch > 0 and stats["bytes"] else None),
            "factors_checksum": chk, "ranks": ctx.world,
            "api": "asvd4llm_b200.sharding.decompose_sharded (LPT owners, binary_search final pass, per-owner packed broadcast)",
            "gpu_launches": int(_lib.launch_count() - launches0)}


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="svd", choices=["svd", "forward", "llama7b"])
    ap.add_argument("--batch", type=int, default=0, help="same-shape weights per step per GPU (0 = the product's own choice)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 4 / config 3 measurements attached to the default line")
    ap.add_argument("--llama-blocks", type=int, default=32)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    ctx = Ctx(args)
    from asvd4llm_b200 import _lib
    _lib.load()

    if args.workload == "forward":
        with ClockSampler(ctx.local) as clocks:
            res, launches, K, W = forward_workload(ctx, steps=args.steps)
        tot_us = sum(v["us"] for v in res.values())
        value = ctx.world * 4 * 65536 / tot_us * 1e6
        pk = peaks()
        worst = min(res.values(), key=lambda v: v["frac_of_bf16_burst_peak"])
        fl = sum(2.0 * 65536 * int(r) * 8192 for r in res)
        # e2e: activations from pinned host memory, result read back
        from asvd4llm_b200.modules.svd_linear import SVDLinear
        g = torch.Generator(device=ctx.dev).manual_seed(3)
        xh = torch.empty(32, 2048, 4096, dtype=torch.half).pin_memory()
        xh.copy_((torch.randn(32, 2048, 4096, device=ctx.dev, generator=g) * 0.125).half())
        yh = torch.empty(32, 2048, 4096, dtype=torch.half).pin_memory()
        mod = SVDLinear._from_factors((torch.randn(4096, 256, device=ctx.dev, generator=g) / 16).half(),
                                      (torch.randn(256, 4096, device=ctx.dev, generator=g) / 64).half(), None)
        with torch.no_grad():
            for _ in range(2):
                yh.copy_(mod(xh.to(ctx.dev, non_blocking=True)))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                yh.copy_(mod(xh.to(ctx.dev, non_blocking=True)))
            torch.cuda.synchronize()
        e2e_s = ctx.max_over_ranks(time.perf_counter() - t0) / 3
        if ctx.rank == 0:
            _emit({"metric": "SVDLinear.forward throughput (mean over the four ranks)", "value": value, "unit": "tokens/s",
                   "n_gpus": ctx.world, "steps": K, "warmup": W, "ms_per_step": tot_us / 1e3, "higher_is_better": True, "scaling": "weak",
                   "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                   "config": FWD_CONFIG, "run": {"per_rank": res},
                   "roofline": {"bound": "tensor", "achieved": fl / tot_us / 1e6, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                                "frac": fl / tot_us / 1e6 / pk["bf16_tflops"], "traffic": None,
                                "peak_source": pk["source"] + ", bf16_tflops (burst: kernels timed alone)",
                                "note": "flops = 2*M*r*(n+m) summed over the four ranks / summed time; per-rank fractions in config.per_rank; "
                                        f"lowest: {worst['frac_of_bf16_burst_peak']} (r = 256 is HBM-bound: 1 GB of x and y per 275 GFLOP)"},
                   "cpu_baseline": None,
                   "e2e": {"value": ctx.world * 65536 / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": 32 * 2048 * 4096 * 2,
                           "d2h_bytes_per_step": 32 * 2048 * 4096 * 2, "api": "SVDLinear.forward (r = 256), x in pinned host memory, y copied back"},
                   "gpu_launches": int(launches), "clocks": clocks.result})
    elif args.workload == "llama7b":
        with ClockSampler(ctx.local) as clocks:
            res = llama_workload(ctx, args.llama_blocks)
        if ctx.rank == 0:
            cpu = None
            if ctx.world == 1 and not args.no_cpu_baseline:
                from oracle import asvd_oracle as O
                torch.set_num_threads(_host_threads())
                per_shape, total = llama_cpu_estimate(O)
                cpu = {"value": 225 / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": reference_kind(),
                       "sample": "one weight of each of the four shapes timed on the host (upstream's svd_lowrank path), multiplied by the shape's count: "
                                 + json.dumps(per_shape)}
            _emit({"metric": METRIC, "value": res["matrices_per_s"], "unit": UNIT, "n_gpus": ctx.world, "steps": 1, "warmup": 0,
                   "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                   "dtype": "f32", "data": "synthetic", "config": {"workload": LLAMA_WORKLOAD}, "run": res,
                   "roofline": None, "cpu_baseline": cpu,
                   "e2e": {"value": res["matrices_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "note": "weights live on the GPU as in upstream's GPU runs; the exchange is inside the timed region"},
                   "gpu_launches": res["gpu_launches"], "clocks": clocks.result})
    else:
        line = svd_workload(ctx)
        extras = None
        if not args.no_extras:
            extras = {}
            try:
                res, launches, K, W = forward_workload(ctx, steps=5)
                if ctx.rank == 0:
                    extras["forward_config4"] = {"workload": FWD_WORKLOAD, "steps": K, "warmup": W, "per_rank": res, "gpu_launches": int(launches),
                                                 "peak": peaks()}
                torch.cuda.empty_cache()
                lres = llama_workload(ctx, args.llama_blocks)
                if ctx.rank == 0:
                    extras["llama7b_config3"] = dict({"workload": LLAMA_WORKLOAD, "scaling": "strong"}, **lres)
            except Exception as e:       # noqa  (the headline line must still be printed)
                if ctx.rank == 0:
                    extras["error"] = repr(e)[:300]
        if ctx.rank == 0:
            line["extras"] = extras
            _emit(line)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
