"""Top stalled SASS lines of one kernel in an ncu report: python scripts/ncu_top.py rep.ncu-rep <1-based kernel id> [n]"""
import csv
import io
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1])
h = rows[1]
si, so = h.index("# Samples"), h.index("Source")
stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [(int(r[si]) if r[si].isdigit() else 0, i, r) for i, r in enumerate(rows[2:])]
tot = sum(d[0] for d in data)
print("total samples", tot)
for cnt, i, r in sorted(data, key=lambda d: -d[0])[:n]:
    top = sorted(((int(r[k]) if r[k].isdigit() else 0, h[k]) for k in stalls), reverse=True)[:2]
    ctx = " | ".join(data[j][2][so].strip()[:38] for j in range(max(0, i - 3), i))
    print("%5.1f%% L%-4d %-44s %s   <- %s" % (100 * cnt / tot, i, r[so].strip()[:44], ",".join("%s:%d" % (b[6:], a) for a, b in top), ctx))
