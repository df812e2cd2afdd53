"""Where do the executed instructions go?  python scripts/ncu_inst.py rep <1-based kernel id> [n]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1])
h = rows[1]
ie, so = h.index("Instructions Executed"), h.index("Source")
data = [(int(r[ie]) if r[ie].isdigit() else 0, i, r[so].strip()) for i, r in enumerate(rows[2:])]
tot = sum(d[0] for d in data)
print("total warp-instructions", tot)
# contiguous regions of equal count = basic blocks; report the heaviest blocks
blocks = []
cur = None
for cnt, i, src in data:
    if cur and cur[0] == cnt:
        cur[2] = i; cur[3] += cnt
    else:
        cur = [cnt, i, i, cnt, src]; blocks.append(cur)
for b in sorted(blocks, key=lambda b: -b[3])[:n]:
    ops = " | ".join(d[2][:28] for d in data[b[1]:min(b[2] + 1, b[1] + 4)])
    print("%5.1f%%  lines %4d-%-4d x%-9d  %s" % (100 * b[3] / tot, b[1], b[2], b[0], ops))
