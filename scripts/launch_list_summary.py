"""Summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`): launches, total and share of
the summed kernel time per kernel name.  Usage: python scripts/launch_list_summary.py X.csv[.gz]"""
import collections
import csv
import gzip
import io
import re
import sys

path = sys.argv[1]
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(io.StringIO("".join(l for l in f if l.startswith('"')))))
hdr = rows[0]
iname, imetric, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
UNIT = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
acc = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= ival or r[imetric] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[iname]).strip()
    name = re.sub(r"^void |dimb::|\(anonymous namespace\)::|<unnamed>::", "", name)
    us = float(r[ival].replace(",", "")) * UNIT.get(r[iunit], 1.0)
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in acc.values())
print(f"{sum(v[0] for v in acc.values())} launches, {tot / 1e3:.2f} ms of kernel time (profiler-serialised, cold caches)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {us / 1e3:10.3f} {us / tot:7.1%}")
