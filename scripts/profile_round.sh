#!/bin/bash
# Round profile pass (run on the GPU box through gpurun): launch list of the bench command with per-launch DRAM
# bytes, and one `ncu --set full` capture of each hot kernel.  Outputs land in gpurun_out/ and are summarised into
# profiles/ by scripts/ncu_traffic.py / scripts/ncu_summary.py.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r01}
PARTS=${PARTS:-launches gemm favor attn}       # which passes to run
mkdir -p gpurun_out
if [[ " $PARTS " == *" launches "* ]]; then
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -c 1400 --csv --log-file gpurun_out/${TAG}_launches_all.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode \
  > gpurun_out/${TAG}_ncu_bench.log 2>&1
# keep the last two complete steps (adam kernel = end of a step) to stay small
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/${TAG}_launches_all.csv")))
hdr = [r for r in rows if r and r[0] == "ID"]
body = [r for r in rows if len(r) > 10 and r[0].isdigit()]
adam = sorted({int(r[0]) for r in body if "adam_kernel" in r[4]})
lo, hi = (adam[-3] + 1, adam[-1]) if len(adam) >= 3 else (0, 10 ** 9)
w = csv.writer(open("gpurun_out/${TAG}_launches_bench.csv", "w"))
w.writerow(hdr[0])
for r in body:
    if lo <= int(r[0]) <= hi: w.writerow(r)
PY
fi
if [[ " $PARTS " == *" gemm "* ]]; then
# ncu --set full of every GEMM shape of a layer (12 launches): the raw metric page goes to CSV on the box and the
# (large) report is dropped -- gpurun brings back at most 64 MiB; one single-kernel report with source is kept
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc -c 12 -o gpurun_out/${TAG}_gemm_all python scripts/gemm_shapes.py > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_gemm_all.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm_shapes_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_gemm_all.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 1 -s 2 -o gpurun_out/${TAG}_gemm_ffn1 python scripts/gemm_shapes.py > /dev/null 2>&1
fi
if [[ " $PARTS " == *" favor "* ]]; then
B=74 timeout 600 ncu --set full --clock-control none --import-source on -k regex:favor_fwd_tc_kernel -c 1 -s 3 -o gpurun_out/${TAG}_favor_fwd python scripts/favor_perf.py > /dev/null 2>&1
B=74 timeout 600 ncu --set full --clock-control none --import-source on -k regex:favor_bwd_tc_kernel -c 1 -s 3 -o gpurun_out/${TAG}_favor_bwd python scripts/favor_perf.py > /dev/null 2>&1
fi
if [[ " $PARTS " == *" attn "* ]]; then
# GPT-2 causal attention (tcgen05 + TMA) at B=16 T=2048: the p = 0.1 launches (the training configuration)
AB=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc_kernel -c 1 -s 16 -o gpurun_out/${TAG}_attn_fwd python scripts/attn_perf.py > /dev/null 2>&1
AB=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc_kernel -c 1 -s 16 -o gpurun_out/${TAG}_attn_bwd python scripts/attn_perf.py > /dev/null 2>&1
fi
rm -f gpurun_out/${TAG}_launches_all.csv
du -sh gpurun_out
ls -la gpurun_out/${TAG}_*
