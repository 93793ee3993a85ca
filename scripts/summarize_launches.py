"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys


def summarize(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"| share | launches | avg us | kernel |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {v[1] / tot * 100:.2f}% | {v[0]} | {v[1] / v[0]:.1f} | `{k[:100]}` |")
    out.append(f"\ntotal {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
