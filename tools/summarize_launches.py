#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the captured window)."""
import collections
import csv
import re
import sys


def main(path, title=''):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('ptta::', '')
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1000 if unit in ('ns', 'nsecond') else (v * 1000 if unit in ('ms', 'msecond') else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print('# %s' % (title or path))
    print('%d launches, %.1f us total (cold-cache, serialised under ncu: compare shares, not absolutes)\n' % (n, tot))
    print('| kernel | launches | total us | avg us | share |')
    print('|---|---:|---:|---:|---:|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| %s | %d | %.1f | %.1f | %.1f%% |' % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


if __name__ == '__main__':
    main(*sys.argv[1:])
