"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=25):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(float); cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        unit = r.get("Metric Unit", "ns")
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v)
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    print(f"total {total/1e3:.3f} ms over {sum(cnt.values())} launches")
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{v/1e3:10.3f} ms {100*v/total:6.2f}%  n={cnt[name]:5d}  avg {v/cnt[name]:9.1f} us  {name[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
