"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean and share."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path, errors='ignore')) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hdr]
    ik, im, iv, iu = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[hdr + 1:]:
        if r[im] != 'gpu__time_duration.sum':
            continue
        v = float(r[iv].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 'usecond': v, 'nsecond': v / 1e3, 'msecond': v * 1e3}.get(r[iu], v)
        name = r[ik].split('(')[0].replace('void ', '').replace('gptq::<unnamed>::', '')
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f'{"kernel":60s} {"launches":>8s} {"total us":>10s} {"mean us":>8s} {"share":>7s}')
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f'{k[:60]:60s} {cnt[k]:8d} {tot[k]:10.1f} {tot[k] / cnt[k]:8.2f} {tot[k] / total:7.1%}')
    print(f'{"TOTAL":60s} {sum(cnt.values()):8d} {total:10.1f}')


if __name__ == '__main__':
    main(sys.argv[1])
