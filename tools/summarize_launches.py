"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0.0, 0])
    for k, t in rows:
        k = re.sub(r"\(.*$", "", k)
        k = re.sub(r"^void ", "", k)
        agg[k][0] += t
        agg[k][1] += 1
    print(f"total {tot / 1e3:.2f} ms over {len(rows)} launches")
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}% {n:5d}x  avg {t / n:8.1f} us  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
