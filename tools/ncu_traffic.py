"""Per-launch DRAM traffic of dg_step_kernel from `ncu --set full` reports -> profiles/ncu_traffic.json (read by bench.py).
Usage: python tools/ncu_traffic.py <config>:<n_envs>=<report.ncu-rep> ..."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
out_path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
table = json.load(open(out_path)) if os.path.isfile(out_path) else {}
UNITS = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
for spec in sys.argv[1:]:
    key, rep = spec.split('=')
    rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(name)
        tot += float(vals[i]) * UNITS[units[i]]
    table[key] = tot
    print(key, tot)
json.dump(table, open(out_path, 'w'), indent=1, sort_keys=True)
