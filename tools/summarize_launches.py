#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: share, us/launch, launches."""
import collections
import csv
import sys


def main(path, steps=None, out=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        n = x["Kernel Name"].split("(")[0][:120]
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    rows = ["| share | us/launch | launches | kernel |", "|---|---|---|---|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        rows.append(f"| {t / tot * 100:.2f}% | {t / c:.1f} | {c} | `{n}` |")
    rows.append("")
    rows.append(f"total {tot / 1e3:.2f} ms" + (f" over {steps} steps = {tot / 1e3 / steps:.2f} ms/step" if steps else ""))
    text = "\n".join(rows)
    if out:
        open(out, "a").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
