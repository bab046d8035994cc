"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel (share of total device time; DRAM traffic per launch when captured).  `--json out.json` also writes the summary."""
import csv
import json
import re
import sys
from collections import defaultdict


def main(path, top=40, json_out=None):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    per_launch = defaultdict(dict)  # launch id -> {metric: value, name, grid, block}
    for r in csv.DictReader(lines):
        d = per_launch[r["ID"]]
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))
        d["name"], d["grid"], d["block"] = name, r["Grid Size"], r["Block Size"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v / 1e3)
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d[m] = v * scale
    rows = [d for d in per_launch.values() if "us" in d]
    tot = sum(d["us"] for d in rows)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for d in rows:
        a = agg[d["name"]]
        a[0] += 1
        a[1] += d["us"]
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    print(f"{len(rows)} launches, {tot/1e3:.2f} ms device time")
    out = []
    for n, (c, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        extra = f"  dram {by/c/1e6:8.2f} MB/launch" if by > 0 else ""
        print(f"{us/1e3:9.3f} ms {100*us/tot:5.1f}% {c:5d}x avg {us/c:8.1f} us  {n[:100]}{extra}")
        out.append({"kernel": n, "launches": c, "ms": round(us / 1e3, 3), "share": round(us / tot, 4), "avg_us": round(us / c, 2),
                    "dram_bytes_per_launch": round(by / c) if by > 0 else None})
    if json_out:
        json.dump({"launches": len(rows), "device_ms": round(tot / 1e3, 3), "kernels": out}, open(json_out, "w"), indent=1)
    return rows


if __name__ == "__main__":
    args = sys.argv[1:]
    jo = None
    if "--json" in args:
        i = args.index("--json")
        jo = args[i + 1]
        del args[i : i + 2]
    main(args[0], int(args[1]) if len(args) > 1 else 40, jo)
