#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count / total / mean (us), in launch order of
first appearance; with --last N only the last N launches (e.g. one batch).   python tools/ncu_launches.py file.csv [--last N]"""
import csv
import sys


def main():
    path = sys.argv[1]
    last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v
            rows.append((r["Kernel Name"].split("(")[0], us))
    if last:
        rows = rows[-last:]
    order, agg = [], {}
    for k, us in rows:
        if k not in agg:
            agg[k] = [0, 0.0]; order.append(k)
        agg[k][0] += 1; agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    print("%-60s %6s %10s %10s %6s" % ("kernel", "n", "total us", "mean us", "share"))
    for k in order:
        n, t = agg[k]
        print("%-60s %6d %10.1f %10.2f %5.1f%%" % (k[:60], n, t, t / n, 100 * t / max(tot, 1e-9)))
    print("%-60s %6d %10.1f" % ("TOTAL", len(rows), tot))


if __name__ == "__main__":
    main()
