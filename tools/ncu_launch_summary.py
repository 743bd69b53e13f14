"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/ncu_launch_summary.py <csv>"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
if not rows:
    sys.exit('no launches in %s (ncu skipped past the end of the run?)' % sys.argv[1])
hdr = rows[0]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(',', ''))
    v = v / 1e3 if r[iu] in ('ns', 'nsecond') else v * (1e3 if r[iu] in ('ms', 'msecond') else 1.0)
    name = r[ik].split('(')[0]
    t = tot.setdefault(name, [0, 0.0, 0.0])
    t[0] += 1
    t[1] += v
    t[2] = max(t[2], v)
allus = sum(t[1] for t in tot.values())
print('%-60s %6s %12s %10s %10s %7s' % ('kernel', 'n', 'total us', 'avg us', 'max us', 'share'))
for name, (n, us, mx) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-60s %6d %12.1f %10.1f %10.1f %6.1f%%' % (name[:60], n, us, us / n, mx, 100 * us / allus))
