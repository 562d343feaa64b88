"""Aggregate warp-stall samples of an ncu report by CUDA source line:  python scripts/ncu_lines.py x.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = {}
fname = None
rd = csv.reader(io.StringIO(txt))
hdr = None
for r in rd:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= si: continue
    try: n = int(r[si]); ins = int(r[ie])
    except ValueError: continue
    if r[0] in ("", "-"): continue       # sass rows are nested; cuda rows carry the line number
    key = (fname, r[0], r[1].strip()[:100])
    a = agg.setdefault(key, [0, 0]); a[0] += n; a[1] += ins
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print("total samples", tot, "instructions", toti)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6d %5.1f%%  inst %5.1f%%  %s:%s  %s" % (a[0], 100 * a[0] / tot, 100 * a[1] / toti, k[0], k[1], k[2]))
