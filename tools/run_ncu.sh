#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, one --set full capture each of stage 1 (64 MiB bench
# document), validate_utf8 and minify (256 MiB), summarised with the kernel-source hash bench.py checks
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
R=${1:-r2}
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; tail -c 300 gpurun_out/ncu_bench.log; python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/${R}_launches.csv")))
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[rows.index(hdr) + 1:]:
    if len(r) > vi:
        a = agg.setdefault(r[ki][:90], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
with open("gpurun_out/${R}_launches_summary.txt", "w") as f:
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"{n:5d} launches  {t/1e3:10.1f} us total  {t/n/1e3:8.2f} us mean  {k}"
        print(line); f.write(line + "\n")
PY
echo "== stage 1 full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan4_kernel -s 4 -c 1 -o gpurun_out/${R}_scan4_full -f python bench.py --steps 3 --warmup 1 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
echo "== utf8 + minify full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"utf8v2_kernel|scan4_minify_kernel" -s 6 -c 3 -o gpurun_out/${R}_utf8_minify_full -f python bench.py --config utf8_minify_256m --steps 3 > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
python tools/ncu_summary.py gpurun_out/${R}_scan4_full.ncu-rep gpurun_out/${R}_scan4_kernel_ncu_full.json
python tools/ncu_summary.py gpurun_out/${R}_utf8_minify_full.ncu-rep gpurun_out/${R}_utf8_minify_ncu_full.json
ncu -i gpurun_out/${R}_scan4_full.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${R}_scan4_source.csv 2>/dev/null; python tools/ncu_lines.py gpurun_out/${R}_scan4_source.csv 40 > gpurun_out/${R}_scan4_kernel_hot_lines.txt; head -25 gpurun_out/${R}_scan4_kernel_hot_lines.txt
ncu -i gpurun_out/${R}_utf8_minify_full.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${R}_um_source.csv 2>/dev/null; python tools/ncu_lines.py gpurun_out/${R}_um_source.csv 40 > gpurun_out/${R}_utf8_minify_hot_lines.txt; head -30 gpurun_out/${R}_utf8_minify_hot_lines.txt
rm -f gpurun_out/${R}_scan4_source.csv gpurun_out/${R}_um_source.csv
ls -la gpurun_out | head -40
