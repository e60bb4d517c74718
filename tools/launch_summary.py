"""Summarise an ncu launch list (gpu__time_duration.sum csv): time and share per kernel name.
python tools/launch_summary.py launches.csv [skip_first_n] [stop_at_n]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
stop = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr_i]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
n = 0
for r in rows[hdr_i + 1:]:
    if len(r) <= vi:
        continue
    n += 1
    if n <= skip:
        continue
    if n > stop:
        n -= 1
        break
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(u, 1)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("pio::<unnamed>::", "").strip()
    agg[name][0] += 1
    agg[name][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"{n - skip} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised: compare SHARES)")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * t / tot:6.2f}%  {t / 1e6:9.3f} ms  {c:6d} x {t / c / 1e3:9.1f} us  {name}")
