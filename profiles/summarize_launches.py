#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python summarize_launches.py in.csv [title]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# {title}\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:100]}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    print(f"\nTotal {tot / 1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches.")


if __name__ == "__main__":
    main()
