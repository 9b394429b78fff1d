#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== probe default" ; PROBE_KERNEL=4 PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -1
for v in nosleep park3 park5; do echo "== probe $v" ; SJB200_LIB=$PWD/tools/variants/lib_$v.so PROBE_KERNEL=4 PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -1 ; done
echo "== timeline default" ; PROBE_KERNEL=4 timeout 300 python tools/timeline.py > gpurun_out/timeline_k4.txt 2>&1 ; tail -30 gpurun_out/timeline_k4.txt
echo "== 1 GiB" ; PROBE_KERNEL=4 PROBE_BYTES=$((1<<30)) PROBE_KINDS=stage1 timeout 600 python tools/gpu_probe.py 2>&1 | tail -1
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan4_kernel -s 2 -c 1 -o gpurun_out/scan4_full -f python bench.py --steps 3 --warmup 1 > gpurun_out/ncu_full.log 2>&1 ; tail -1 gpurun_out/ncu_full.log
