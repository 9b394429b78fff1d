#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== timeline default" ; PROBE_KERNEL=4 timeout 300 python tools/timeline.py > gpurun_out/timeline_k4.txt 2>&1 ; tail -34 gpurun_out/timeline_k4.txt | cut -c1-400
