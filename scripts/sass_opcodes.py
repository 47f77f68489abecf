"""Histogram of the SASS opcodes of libdimb200.so per kernel family (cuobjdump -sass): the evidence that the tensor-core / TMA /
TMEM paths are what is compiled in.  python scripts/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "dyadic-interaction-modeling_b200", "libdimb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAPF", "LDTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "HMMA", "LDSM", "LDGSTS", "SYNCS", "UCGABAR", "REDG", "ATOMG",
         "BAR", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "DADD", "CCTL", "MEMBAR", "ERRBAR", "BPT")
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        d = d.replace("(anonymous namespace)::", "").replace("dimb::", "")
        d = re.sub(r"^void ", "", d)
        cur = re.sub(r"\(.*", "", d)
        hist.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        hist[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                hist[cur][w] += 1
                break
print(f"# SASS opcode histogram of {os.path.basename(so)} (sm_100a), one line per kernel: total instructions, then the watched opcode families")
print("# UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, HMMA/LDSM = mma.sync / ldmatrix, LDGSTS = cp.async")
for k, c in sorted(hist.items(), key=lambda kv: -kv[1]["_total"]):
    rest = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{k[:78]:78s} total={c['_total']:6d}  {rest}")
