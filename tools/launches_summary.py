"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 2:]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) < len(hdr):
            continue
        v = float(r[ix['Metric Value']])
        unit = r[ix['Metric Unit']]
        v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
        agg[r[ix['Kernel Name']]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print('kernel,launches,avg_us,total_us,share')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print('"%s",%d,%.1f,%.1f,%.3f' % (k[:110], len(v), sum(v) / len(v), sum(v), sum(v) / tot))


if __name__ == '__main__':
    main(sys.argv[1])
