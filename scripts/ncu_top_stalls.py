"""Top stalled SASS instructions of an `ncu --page source --csv` dump (stdin), with a little context."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[1]
body = rows[2:]
iS, iN, iX = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iN] or 0) for r in body)
print("total samples", tot)
order = sorted(range(len(body)), key=lambda k: -int(body[k][iN] or 0))[: int(sys.argv[1]) if len(sys.argv) > 1 else 25]
for k in order:
    r = body[k]
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {int(r[iN]):7d} {100*int(r[iN])/tot:5.1f}% x{r[iX]:>8s} {r[iS].strip()[:70]:70s} {st}")
