#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for v in park4 park4k5 k5; do echo "== probe $v" ; SJB200_LIB=$PWD/tools/variants/lib_$v.so PROBE_KERNEL=4 PROBE_KINDS=stage1 timeout 200 python tools/gpu_probe.py 2>&1 | tail -1 ; done
echo "== probe default" ; PROBE_KERNEL=4 PROBE_KINDS=stage1 timeout 200 python tools/gpu_probe.py 2>&1 | tail -1
echo "== 256 MiB park4" ; SJB200_LIB=$PWD/tools/variants/lib_park4.so PROBE_KERNEL=4 PROBE_BYTES=$((256<<20)) PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -1
echo "== 256 MiB default" ; PROBE_KERNEL=4 PROBE_BYTES=$((256<<20)) PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -1
