"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], val * scale))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for k, us in rows:
        name = k.split("(")[0][:70]
        tot[name] += us
        cnt[name] += 1
    total = sum(tot.values()) or 1.0
    print(f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for name, us in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{name:72s} {cnt[name]:8d} {us:12.1f} {us / cnt[name]:10.2f} {100 * us / total:6.1f}%")
    print(f"{'TOTAL':72s} {len(rows):8d} {total:12.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
