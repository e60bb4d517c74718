"""List the hottest SASS instructions of an ncu report: python tools/ncu_hot.py report.ncu-rep [kernel-index] [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
k = kernels[which]
h = k["hdr"]
si = h.index("# Samples")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data = []
for idx, r in enumerate(k["rows"]):
    try:
        data.append((float(r[si]), idx, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print(k["name"][:120], "total samples", tot, "instructions", len(data))
for s, idx, r in sorted(data, reverse=True)[:top]:
    st = sorted(((float(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{s:8.0f} {100 * s / tot:5.1f}%  #{idx:5d} {r[1].strip()[:70]:70s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")
