#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== probe deferred" ; PROBE_KERNEL=4 PROBE_DEFERRED=1 PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -1
echo "== probe pipelined" ; PROBE_KERNEL=4 PROBE_DEFERRED=0 PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -1
for v in w16 w16k5; do for d in 1 0; do echo "== probe $v deferred=$d" ; SJB200_LIB=$PWD/tools/variants/lib_$v.so PROBE_KERNEL=4 PROBE_DEFERRED=$d PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -1 ; done; done
echo "== timeline deferred" ; PROBE_KERNEL=4 PROBE_DEFERRED=1 timeout 300 python tools/timeline.py > gpurun_out/timeline_k4.txt 2>&1 ; grep -v "^   #\|^gate" gpurun_out/timeline_k4.txt | tail -16 | cut -c1-400
echo "== 16 MiB" ; PROBE_KERNEL=4 PROBE_BYTES=$((16<<20)) PROBE_KINDS=stage1 timeout 600 python tools/gpu_probe.py 2>&1 | tail -1
echo "== 256 MiB deferred=2" ; PROBE_KERNEL=4 PROBE_DEFERRED=2 PROBE_BYTES=$((256<<20)) PROBE_KINDS=stage1 timeout 600 python tools/gpu_probe.py 2>&1 | tail -1
echo "== 256 MiB pipelined" ; PROBE_KERNEL=4 PROBE_DEFERRED=0 PROBE_BYTES=$((256<<20)) PROBE_KINDS=stage1 timeout 600 python tools/gpu_probe.py 2>&1 | tail -1
echo "== 1 GiB w16" ; SJB200_LIB=$PWD/tools/variants/lib_w16.so PROBE_KERNEL=4 PROBE_BYTES=$((1<<30)) PROBE_KINDS=stage1 timeout 600 python tools/gpu_probe.py 2>&1 | tail -1
echo "== 1 GiB default" ; PROBE_KERNEL=4 PROBE_BYTES=$((1<<30)) PROBE_KINDS=stage1 timeout 600 python tools/gpu_probe.py 2>&1 | tail -1
echo "== timeline w16 deferred" ; SJB200_LIB=$PWD/tools/variants/lib_w16.so PROBE_KERNEL=4 PROBE_DEFERRED=1 timeout 300 python tools/timeline.py 2>&1 | grep -v "^   #\|^gate" | tail -14 | cut -c1-400
