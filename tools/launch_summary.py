#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and share per kernel."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[iu], 1e-6)
    name = r[ik].split("(")[0].replace("void ", "").replace("phos::", "")
    tot[name] += v
    cnt[name] += 1
all_ms = sum(tot.values())
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k:40s} {cnt[k]:6d} launches {tot[k]:10.3f} ms {100 * tot[k] / all_ms:5.1f} %  avg {1e3 * tot[k] / cnt[k]:9.1f} us")
print(f"{'total':40s} {sum(cnt.values()):6d} launches {all_ms:10.3f} ms")
