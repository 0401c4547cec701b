#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel launch count, total/avg device time and share of the total."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    d = collections.defaultdict(lambda: [0, 0.0])
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6,
             'usecond': 1e-3, 'msecond': 1.0, 's': 1e3, 'second': 1e3}
    for r in rows[hdr + 2:]:
        if len(r) < len(rows[hdr]):
            continue
        rec = dict(zip(rows[hdr], r))
        name = re.sub(r'\(.*', '', rec['Kernel Name'])[:72]
        try:
            v = float(rec['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        d[name][0] += 1
        d[name][1] += v * scale[rec['Metric Unit']]
    tot = sum(v[1] for v in d.values())
    print("| kernel | launches | total ms | avg ms | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.4f | %.1f%% |"
              % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
