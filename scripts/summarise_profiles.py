"""Turns the ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python scripts/summarise_profiles.py gpurun_out/<report>.ncu-rep [...]  ->  profiles/<report>.summary.csv
  python scripts/summarise_profiles.py --launches gpurun_out/launches_r01.csv -> profiles/launches_r01.summary.txt"""
import csv, subprocess, sys, os, collections, json

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]

def summarise_report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = os.path.join("profiles", os.path.basename(path).replace(".ncu-rep", ".summary.csv"))
    traffic = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        cols = ["Kernel Name"] + [k for k in KEEP if k in idx]
        w.writerow(cols); w.writerow([""] + [units[idx[k]] for k in cols[1:]])
        for r in rows[2:]:
            w.writerow([r[idx[c]] for c in cols])
            name = r[idx["Kernel Name"]].split("(")[0].split("::")[-1]
            try:
                def val(k):
                    v, u = float(r[idx[k]]), units[idx[k]]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                traffic.setdefault(name, []).append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
            except Exception:
                pass
    print("wrote", out)
    return {k: sum(v) / len(v) for k, v in traffic.items()}

def summarise_launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Unit, Value
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        unit, val = r[-2], float(r[-1].replace(",", ""))
        us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    out = os.path.join("profiles", os.path.basename(path).replace(".csv", ".summary.txt"))
    with open(out, "w") as f:
        f.write(f"# per-kernel launch counts and device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({os.path.basename(path)})\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name[:60]:60s} {n:8d} {us:12.1f} {us / n:10.1f} {us / tot:7.3f}\n")
    print("wrote", out)

if __name__ == "__main__":
    os.makedirs("profiles", exist_ok=True)
    args = sys.argv[1:]
    if args and args[0] == "--launches":
        for p in args[1:]: summarise_launches(p)
    else:
        traffic = {}
        for p in args: traffic.update(summarise_report(p))
        if traffic:
            tp = os.path.join("profiles", "traffic.json")
            old = json.load(open(tp)) if os.path.exists(tp) else {}
            old.update({k.replace("_kernel", ""): v for k, v in traffic.items()})
            json.dump(old, open(tp, "w"), indent=1)
            print("updated", tp, old)
