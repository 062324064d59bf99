#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel count, total, average and share of the listed launches."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (ValueError, KeyError):
            continue
        unit = row.get('Metric Unit', 'ns')
        v *= {'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6}.get(unit, 1.0)
        a = agg[row['Kernel Name'].split('(')[0][:70]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f'| kernel | launches | total ms | avg us | share |')
    print('|---|---:|---:|---:|---:|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / v[0] / 1e3:.1f} '
              f'| {v[1] / tot:.1%} |')


if __name__ == '__main__':
    main(sys.argv[1])
