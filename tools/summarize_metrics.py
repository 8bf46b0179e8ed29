#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` capture by kernel AND grid size:
launches, average duration, average DRAM bytes, DRAM GB/s and its fraction of the measured HBM peak (cold-cache, serialised launches)."""
import collections, csv, json, os, re, sys


def main(path, title='', last=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    peak = 6456.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        i = int(row['ID'])
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('ptta::', '')
        d = per.setdefault(i, {'name': name, 'grid': row['Grid Size'], 'block': row['Block Size']})
        v = float(row['Metric Value'].replace(',', ''))
        m, unit = row['Metric Name'], row['Metric Unit']
        if m.startswith('gpu__time'):
            v = v / 1000 if unit in ('ns', 'nsecond') else (v * 1000 if unit in ('ms', 'msecond') else v)
            d['us'] = v
        else:
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
            d['rd' if 'read' in m else 'wr'] = v
    ids = sorted(per)
    if last:
        ids = ids[-int(last):]
    agg = collections.OrderedDict()
    for i in ids:
        d = per[i]
        a = agg.setdefault((d['name'], d['grid']), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get('us', 0); a[2] += d.get('rd', 0); a[3] += d.get('wr', 0)
    tot = sum(a[1] for a in agg.values())
    print('# %s' % (title or path))
    print('%d launches, %.1f us of kernel time (cold-cache, serialised under ncu); HBM peak %.1f GB/s (MEASURED_PEAKS.json)\n' % (len(ids), tot, peak))
    print('| kernel | grid | launches | avg us | share | DRAM read MB | DRAM write MB | DRAM GB/s | of HBM peak |')
    print('|---|---|---:|---:|---:|---:|---:|---:|---:|')
    for (k, g), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        n = a[0]
        gbs = (a[2] + a[3]) / max(a[1], 1e-9) / 1e3
        print('| %s | %s | %d | %.1f | %.1f%% | %.2f | %.2f | %.0f | %.2f |' % (k, g, n, a[1] / n, 100 * a[1] / tot, a[2] / n / 1e6, a[3] / n / 1e6, gbs, gbs / peak))


if __name__ == '__main__':
    main(*sys.argv[1:])
