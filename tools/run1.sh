#!/bin/bash
# round-2 GPU run #1: parity of the new default build, A/B of emit / stagger / chain variants, ablations, timeline, minify on scan4
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl gpurun_out/probe_mu.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== parity (default build)"; SJB200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config2 or fuzz_small or golden or scan4_experimental or kernel_variants or sharded" 2>&1 | tail -4
echo "== 64 MiB variants"
for v in emit0 emit1 emit2 emit1_park4 emit1_stag emit0_stag emit1_counter; do SJB200_LIB=tools/variants/lib_$v.so timeout 200 python tools/probe2.py 2>&1 | tail -1; done
for v in diag_noutf8 diag_noemit; do PROBE_PARITY=0 SJB200_LIB=tools/variants/lib_$v.so timeout 200 python tools/probe2.py 2>&1 | tail -1; done
echo "== 1 GiB"
for v in emit0 emit1 emit1_stag; do PROBE_BYTES=1073741824 SJB200_LIB=tools/variants/lib_$v.so timeout 300 python tools/probe2.py 2>&1 | tail -1; done
echo "== timeline emit1"; SJB200_LIB=tools/variants/lib_emit1.so PROBE_KERNEL=4 timeout 200 python tools/timeline.py > gpurun_out/timeline_emit1.txt 2>&1; grep -v "^   #\|^gate" gpurun_out/timeline_emit1.txt | head -30
echo "== timeline emit1_stag"; SJB200_LIB=tools/variants/lib_emit1_stag.so PROBE_KERNEL=4 timeout 200 python tools/timeline.py > gpurun_out/timeline_emit1_stag.txt 2>&1; grep -v "^   #\|^gate" gpurun_out/timeline_emit1_stag.txt | head -30
echo "== minify / utf8 256 MiB"; timeout 400 python tools/probe_mu.py 2>&1 | tail -1
ls gpurun_out
