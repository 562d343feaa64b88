"""Turn an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
`bench.py --steps 1 --warmup 3` into profiles/r01_traffic.json (+ a per-kernel table on stdout):
    python scripts/ncu_traffic.py gpurun_out/launches.csv <per_gpu_batch> [out.json]"""
import collections, csv, json, re, sys
src, B = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/r01_traffic.json"
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
by = collections.OrderedDict()
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
for r in rows:
    d = by.setdefault(int(r[0]), {"name": re.sub(r"\(.*", "", r[4]).replace("void ", "")})
    d[r[-3]] = float(r[-1]) * U.get(r[-2], 1)
L = list(by.values())
ad = [i for i, d in enumerate(L) if d["name"].startswith("adam")]
if len(ad) >= 2:
    step = L[ad[-2] + 1: ad[-1] + 1]                               # the last complete step
elif ad:                                                           # trimmed list: the step opens with the Omega draw (randn)
    st = [i for i, d in enumerate(L[:ad[-1]]) if "distribution_elementwise" in d["name"]]
    step = L[(st[-1] if st else 0): ad[-1] + 1]
else:
    step = L
agg = collections.OrderedDict()
for d in step:
    a = agg.setdefault(d["name"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("one train step: %d launches, %.1f us of kernel time (ncu: serialised, cold caches)" % (len(step), tot))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%4d %9.1f us %5.1f%%  avg %7.1f us  dram/launch %8.1f MB" % (k[:62], a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2] / a[0] / 1e6))
g = [a for k, a in agg.items() if "gemm_tc_kernel" in k]
n, b, t = sum(a[0] for a in g), sum(a[2] for a in g), sum(a[1] for a in g)
json.dump({"per_gpu_batch": B, "gemm_launches_per_step": n, "gemm_mean_dram_bytes_per_launch": round(b / max(n, 1)),
           "gemm_share_of_kernel_time": round(t / tot, 4), "source": "ncu launch list " + src.split("/")[-1] +
           " (dram__bytes_read.sum + dram__bytes_write.sum, mean over the GEMM launches of one step)"}, open(out, "w"), indent=1)
print("wrote", out)
