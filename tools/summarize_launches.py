#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = d['Kernel Name'].split('(')[0]
        v = float(d['Metric Value'].replace(',', ''))
        u = d['Metric Unit']
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'second': 1e6}.get(u, 1.0)
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print('%-44s %5s %12s %10s %7s' % ('kernel', 'n', 'total_us', 'avg_us', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print('%-44s %5d %12.1f %10.1f %7.3f' % (k[:44], len(v), sum(v), sum(v) / len(v), sum(v) / tot))


if __name__ == '__main__':
    main(sys.argv[1])
