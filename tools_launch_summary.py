"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of total device time)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void ", "", name)
        rows.append((name, r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
    tot = sum(r[3] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, g, b, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    print(f"{len(rows)} launches, {tot/1e3:.2f} ms device time")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{us/1e3:9.3f} ms {100*us/tot:5.1f}% {c:5d}x avg {us/c:8.1f} us  {n[:110]}")
    return rows


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
