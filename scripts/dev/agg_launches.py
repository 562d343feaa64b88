"""aggregate the LAST train step of an ncu launch list (gpu__time_duration.sum) per kernel name"""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
seq = [(re.sub(r"\(.*", "", r[4]).replace("void ", "")[:70], float(r[-1]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[-2], 1e-3)) for r in rows]
ad = [i for i, (n, t) in enumerate(seq) if n.startswith("adam")]
step = seq[ad[-2] + 1: ad[-1] + 1] if len(ad) >= 2 else seq
agg = collections.OrderedDict()
for n, t in step:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print("%s: one train step, %d launches, %.1f us of kernel time (ncu: serialised, cold caches)" % (sys.argv[2], len(step), tot))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-70s n=%4d %9.1f us %5.1f%%" % (n, a[0], a[1], 100 * a[1] / tot))
