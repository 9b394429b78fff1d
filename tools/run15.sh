#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
R=r2m
echo "== minify full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"scan4_minify_kernel" -s 2 -c 1 -o gpurun_out/${R}_minify_full -f python bench.py --config utf8_minify_256m --steps 3 > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
python tools/ncu_summary.py gpurun_out/${R}_minify_full.ncu-rep gpurun_out/${R}_minify_ncu_full.json
ncu -i gpurun_out/${R}_minify_full.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${R}_m_source.csv 2>/dev/null; python tools/ncu_lines.py gpurun_out/${R}_m_source.csv 60 > gpurun_out/${R}_minify_hot_lines.txt; head -70 gpurun_out/${R}_minify_hot_lines.txt
rm -f gpurun_out/${R}_m_source.csv
echo "== stage1 probes"; for mb in 64 256; do PROBE_TAG=init_${mb}m PROBE_BYTES=$((mb<<20)) timeout 300 python tools/probe2.py 2>&1 | tail -1 | cut -c1-260; done
echo "== e2e stats"; timeout 300 python tools/e2e_stats.py 2>&1 | tail -8
