"""per-kernel mean duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hdr]
kn, mv, mu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
d = defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > mv:
        v = float(r[mv].replace(',', ''))
        v = v / 1e3 if r[mu] == 'ns' else v * 1e3 if r[mu] == 'ms' else v
        d[r[kn][:70]].append(v)
for k, v in d.items():
    print("%-72s n=%4d  mean %9.1f us  total %10.1f us" % (k, len(v), sum(v) / len(v), sum(v)))
