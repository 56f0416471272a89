"""Aggregate an ncu source-page CSV (--print-source cuda,sass) by CUDA source line: share of executed
instructions and of stall samples.  usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass
--kernel-name regex:K [--launch-skip N --launch-count 1] > f.csv ; python tools/ncu_lines.py f.csv [warps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[hi]
iI = hdr.index('Instructions Executed')
iS = hdr.index('Warp Stall Sampling (All Samples)')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > iI and r[0].strip().isdigit():
        ln = int(r[0])
        try:
            old = agg.get(ln, (0, 0, ''))
            agg[ln] = (old[0] + int(r[iI]), old[1] + int(r[iS]), r[1][:110])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
warps = float(sys.argv[2]) if len(sys.argv) > 2 else 0
print("total warp instructions", tot, "stall samples", tots)
for ln, (i, s, src) in sorted(agg.items()):
    if i > tot * 0.008 or s > tots * 0.012:
        print("%4d %5.1f%% %s %5.1f%%  %s" % (ln, 100.0 * i / tot, ("%6.0f" % (i / warps)) if warps else "", 100.0 * s / tots, src))
