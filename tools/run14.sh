#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "emit_warp or sizes or sharded or epilogue" 2>&1 | tail -3
echo "== stage1 default lib (>= 384 MiB: emit-warp kernel)"; for mb in 64 256 512 1024; do PROBE_TAG=auto_${mb}m PROBE_BYTES=$((mb<<20)) timeout 300 python tools/probe2.py 2>&1 | tail -1 | cut -c1-260; done
echo "== timeline default 64 MiB"; timeout 300 python tools/timeline.py > gpurun_out/r2_timeline_default.txt 2>&1; head -12 gpurun_out/r2_timeline_default.txt
echo "== trace default 64 MiB"; SJB200_LIB=$PWD/tools/variants/lib_trace.so timeout 300 python tools/trace4.py > gpurun_out/r2_trace4_default.txt 2>&1; head -24 gpurun_out/r2_trace4_default.txt
echo "== ndjson_1g"; timeout 600 python bench.py --config ndjson_1g --steps 10 2>&1 | tail -1 | cut -c1-1500
