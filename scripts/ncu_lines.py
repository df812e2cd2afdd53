"""Warp-stall samples and executed instructions per CUDA source line: python scripts/ncu_lines.py rep.ncu-rep <1-based kernel id> [n]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::" + kid],
                     capture_output=True, text=True).stdout
cur, hdr, rows = None, None, []
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif r[0].isdigit() and hdr:
        si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
        rows.append((int(r[si]) if r[si].isdigit() else 0, int(r[ie]) if r[ie].isdigit() else 0, cur, int(r[0]), r[1].strip()))
ts, ti = sum(r[0] for r in rows), sum(r[1] for r in rows)
print(fn[:60], "samples", ts, "warp-instructions", ti)
for s, i, f, ln, src in sorted(rows, key=lambda r: -r[0])[:n]:
    print("%5.1f%% smp %5.1f%% inst  %-22s %4d  %s" % (100.0 * s / ts, 100.0 * i / ti, f, ln, src[:110]))
