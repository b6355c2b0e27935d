"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.   python scripts/launch_summary.py file.csv [top]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
t = defaultdict(float); n = defaultdict(int)
for r in rows[1:]:
    t[r[ki][:70]] += float(r[vi].replace(",", "")); n[r[ki][:70]] += 1
for k, v in sorted(t.items(), key=lambda x: -x[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print("%10.3f ms %4d  %s" % (v / 1e6, n[k], k))
