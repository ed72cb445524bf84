"""Top stalled SASS instructions per kernel of an .ncu-rep (source page), robust to several kernels in one report.
  python scripts/ncu_src_top.py report.ncu-rep [n] [kernel index]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25; which = int(sys.argv[3]) if len(sys.argv) > 3 else -1
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = starts[which]
end = starts[starts.index(start) + 1] - 1 if start != starts[-1] else len(rows)
print(rows[start - 1][1][:100] if start > 0 else "")
hdr = rows[start]; body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
iS, iN, iX = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iN] or 0) for r in body); print("total samples", tot)
for k in sorted(range(len(body)), key=lambda k: -int(body[k][iN] or 0))[:n]:
    r = body[k]; st = sorted(((int(r[i] or 0), hdr[i]) for i in stall), reverse=True)[:2]
    print(f"{k:5d} {int(r[iN]):7d} {100*int(r[iN])/tot:5.1f}% x{r[iX]:>8s} {r[iS].strip()[:80]:80s} {st}")
