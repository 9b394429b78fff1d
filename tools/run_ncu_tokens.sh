#!/bin/bash
# one ncu --set full capture of the stage-2-lite kernels (one launch each) on the 64 MiB bench document
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
R=${1:-r2}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"token_scan_kernel|string_write_kernel" -s 2 -c 2 -o gpurun_out/${R}_tokens_full -f python bench.py --config tokens_64m --steps 3 > gpurun_out/ncu_tokens.log 2>&1; tail -2 gpurun_out/ncu_tokens.log
python tools/ncu_summary.py gpurun_out/${R}_tokens_full.ncu-rep gpurun_out/${R}_tokens_ncu_full.json
ncu -i gpurun_out/${R}_tokens_full.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${R}_tok_source.csv 2>/dev/null; python tools/ncu_lines.py gpurun_out/${R}_tok_source.csv 45 > gpurun_out/${R}_tokens_hot_lines.txt; head -50 gpurun_out/${R}_tokens_hot_lines.txt
rm -f gpurun_out/${R}_tok_source.csv gpurun_out/${R}_tokens_full.ncu-rep
