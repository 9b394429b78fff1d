#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== probe default" ; PROBE_KERNEL=4 PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -1
echo "== timeline default" ; PROBE_KERNEL=4 timeout 300 python tools/timeline.py > gpurun_out/timeline_k4.txt 2>&1 ; tail -30 gpurun_out/timeline_k4.txt
