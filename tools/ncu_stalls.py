"""Dev tool: aggregate warp-stall samples from `ncu -i X.ncu-rep --page source --csv -k regex:NAME`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
si, src = hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stall_cols}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {k:28s} {v:8d} {100 * v / max(tot, 1):5.1f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
    reasons = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[si]):6d} {100 * int(r[si]) / max(tot, 1):5.1f}%  {r[src].strip()[:64]:64s} {reasons}")
