"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list: per kernel the number of
launches, time per launch, share of the step and DRAM bytes per launch.   python profiles/launch_summary.py <launches.csv> [--json]"""
import collections
import csv
import json
import sys


def summarise(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    kn, mn, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > mv:
            per.setdefault((r[idc], r[kn]), {})[r[mn]] = float(r[mv].replace(',', ''))
    agg = collections.OrderedDict()
    for (_, k), m in per.items():
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += m.get('gpu__time_duration.sum', 0); a[2] += m.get('dram__bytes_read.sum', 0); a[3] += m.get('dram__bytes_write.sum', 0)
    tot = sum(a[1] for a in agg.values())
    return [{'kernel': k, 'launches': a[0], 'us_per_launch': a[1] / a[0] / 1e3, 'share': a[1] / tot, 'dram_read_bytes_per_launch': a[2] / a[0],
             'dram_write_bytes_per_launch': a[3] / a[0]} for k, a in agg.items()]


if __name__ == '__main__':
    out = summarise(sys.argv[1])
    if '--json' in sys.argv:
        print(json.dumps(out, indent=1))
    else:
        for o in out:
            print(f"{o['kernel'][:70]:70s} n={o['launches']:3d} {o['us_per_launch']:9.1f} us/launch  {100 * o['share']:5.1f} %  DRAM R/W "
                  f"{o['dram_read_bytes_per_launch'] / 1e6:8.1f} / {o['dram_write_bytes_per_launch'] / 1e6:8.1f} MB")
