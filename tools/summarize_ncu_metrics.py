"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed` log:
per kernel name -> launches, mean time, achieved HBM GB/s (bytes moved / time), DRAM % of peak, tensor-pipe % (active)."""
import csv
import re
import sys
from collections import defaultdict

UNIT = {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}


def main(path, top=40):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    per = defaultdict(lambda: defaultdict(dict))   # name -> launch id -> metric -> value
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1.0)
        per[name][r["ID"]][r["Metric Name"]] = v
    rows = []
    for name, launches in per.items():
        n = len(launches)
        t = sum(m.get("gpu__time_duration.sum", 0.0) for m in launches.values())
        b = sum(m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0) for m in launches.values())
        tp = sum(m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * m.get("gpu__time_duration.sum", 0.0)
                 for m in launches.values())
        dp = sum(m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0) * m.get("gpu__time_duration.sum", 0.0)
                 for m in launches.values())
        rows.append((t, n, name, b, tp / t if t else 0.0, dp / t if t else 0.0))
    total = sum(r[0] for r in rows)
    print(f"total {total*1e3:.3f} ms over {sum(r[1] for r in rows)} launches (under ncu: cold caches, serialised)")
    print(f"{'time ms':>9} {'share':>6} {'n':>5} {'avg us':>8} {'HBM GB/s':>9} {'DRAM %pk':>8} {'tensor %':>8}  kernel")
    for t, n, name, b, tp, dp in sorted(rows, reverse=True)[:top]:
        print(f"{t*1e3:9.3f} {100*t/total:5.1f}% {n:5d} {t/n*1e6:8.1f} {b/t/1e9:9.0f} {dp:8.1f} {tp:8.1f}  {name[:80]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
