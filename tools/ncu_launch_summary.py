"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
python tools/ncu_launch_summary.py <csv> [steps in the list]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
if not rows:
    sys.exit('no launches in %s (ncu skipped past the end of the run?)' % sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = rows[0]
ik, im, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
SCALE = {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'second': 1e6,
         'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tot = OrderedDict()   # kernel -> [launches, us, max us, dram bytes]
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(',', '')) * SCALE.get(r[iu], 1.0)
    name = r[ik].split('(')[0]
    t = tot.setdefault(name, [0, 0.0, 0.0, 0.0])
    if r[im] == 'gpu__time_duration.sum':
        t[0] += 1
        t[1] += v
        t[2] = max(t[2], v)
    elif r[im].startswith('dram__bytes'):
        t[3] += v
allus = sum(t[1] for t in tot.values())
print('%-60s %6s %12s %10s %10s %7s %12s' % ('kernel', 'n', 'total us', 'avg us', 'max us', 'share', 'DRAM MB'))
for name, (n, us, mx, by) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-60s %6d %12.1f %10.1f %10.1f %6.1f%% %12.1f' % (name[:60], n, us, us / max(n, 1), mx, 100 * us / allus, by / 1e6))
if steps:
    print('per step (%d steps): %.1f us in kernels (serialised, cold cache), %.1f MB of DRAM traffic' %
          (steps, allus / steps, sum(t[3] for t in tot.values()) / 1e6 / steps))
