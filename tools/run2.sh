#!/bin/bash
# round-2 GPU run #2: where the time goes inside scan4 (trace build, size sweep, fixed cost of a launch), the new host-pointer pipeline
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl gpurun_out/e2e_probe.jsonl
echo "== host path parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_pointer_paths or golden_stage1 or capacity" 2>&1 | tail -4
echo "== dropin"; timeout 600 python -m pytest tests/test_dropin.py -x -q -m gpu 2>&1 | tail -3
echo "== trace 64 MiB"; SJB200_LIB=tools/variants/lib_trace.so timeout 200 python tools/trace4.py 2>&1 | tee gpurun_out/trace4_64m.txt | tail -24
echo "== trace 256 MiB"; PROBE_BYTES=268435456 SJB200_LIB=tools/variants/lib_trace.so timeout 200 python tools/trace4.py 2>&1 | tee gpurun_out/trace4_256m.txt | tail -24
echo "== default build, 64 MiB, parity"; PROBE_TAG=default timeout 300 python tools/probe2.py 2>&1 | tail -1
echo "== size sweep (default build)"
for mb in 1 4 16 32 64 128 256; do PROBE_PARITY=0 PROBE_TAG=size_${mb}m PROBE_BYTES=$((mb<<20)) timeout 200 python tools/probe2.py 2>&1 | tail -1; done
echo "== fixed cost: 64 KiB document on a full grid"
PROBE_PARITY=0 PROBE_TAG=tiny_fullgrid PROBE_FORCE_GRID=148 PROBE_BYTES=65536 timeout 200 python tools/probe2.py 2>&1 | tail -1
PROBE_PARITY=0 PROBE_TAG=tiny_grid1 PROBE_BYTES=65536 timeout 200 python tools/probe2.py 2>&1 | tail -1
echo "== e2e through the plug-in"; timeout 600 python tools/e2e_probe.py 2>&1 | grep -v Warning
ls gpurun_out
