"""Per-SASS-instruction executed counts of one kernel from an ncu report's source page.
   ncu -i rep.ncu-rep --page source --csv --kernel-name NAME > k.csv ; python tools/sass_hot.py k.csv [segment]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ia, isrc, ist = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
data = []
for r in rows[hi + 1:]:
    try:
        data.append((r[isrc].strip(), int(r[ia]), int(r[ist])))
    except (ValueError, IndexError):
        break            # next kernel instance
tot = sum(d[1] for d in data)
print('total warp-inst', tot, 'sass lines', len(data), 'samples', sum(d[2] for d in data))
for s in range(0, len(data), seg):
    c = sum(d[1] for d in data[s:s + seg])
    st = sum(d[2] for d in data[s:s + seg])
    ops = {}
    for d in data[s:s + seg]:
        op = d[0].split()[0] if not d[0].startswith('@') else d[0].split()[1]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + d[1]
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    print(f"{s:5d} {c:12d} {100 * c / tot:5.1f}%  stall {st:6d}  " + " ".join(f"{k}:{100 * v / max(c, 1):.0f}%" for k, v in top))
