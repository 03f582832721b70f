"""Summarise an `ncu --page source --csv` export: executed warp-instructions per SASS opcode.
usage: python tools/sass_mix.py src.csv [top_lines]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
    hdr = rows[hi]
    ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot = 0
    byop, samp, data = collections.Counter(), collections.Counter(), []
    for r in rows[hi + 1:]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        s = r[ia].strip()
        n = int(r[ie])
        op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]
        byop[op] += n
        tot += n
        samp[op] += int(r[isamp])
        data.append((n, s, int(r[isamp])))
    print('total warp-instructions', tot)
    for op, n in byop.most_common(32):
        print(f"{op:10s} {n / 1e6:9.2f}M {100 * n / tot:5.1f}%  samples {samp[op]}")
    if top:
        print('--- hottest lines by samples')
        for n, s, sm in sorted(data, key=lambda t: -t[2])[:top]:
            print(f"{sm:6d} {n / 1e6:8.2f}M  {s}")


if __name__ == '__main__':
    main()
