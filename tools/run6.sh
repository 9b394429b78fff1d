#!/bin/bash
# round-2 GPU run #6: chain latency (no look-back sleep, rotating ticket duty), emit warps again, device filters, e2e ramp
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl gpurun_out/e2e_probe.jsonl gpurun_out/probe_mu.jsonl
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "epilogue or config2 or fuzz_small or golden or kernel_variants or host_pointer or scan4_experimental" 2>&1 | tail -4
echo "== 64 MiB variants"
for v in base emitw0 lbsleep100 emitw0_lbsleep100 nolb emitw6; do SJB200_LIB=tools/variants/lib_$v.so timeout 200 python tools/probe2.py 2>&1 | tail -1; done
echo "== 1 GiB"
for v in base emitw0; do PROBE_BYTES=1073741824 SJB200_LIB=tools/variants/lib_$v.so timeout 300 python tools/probe2.py 2>&1 | tail -1; done
echo "== trace base 64 MiB"; SJB200_LIB=tools/variants/lib_trace.so timeout 200 python tools/trace4.py 2>&1 | tee gpurun_out/trace4_64m.txt | tail -22
echo "== timeline base"; SJB200_LIB=tools/variants/lib_base.so PROBE_KERNEL=4 timeout 200 python tools/timeline.py > gpurun_out/timeline_base.txt 2>&1; grep -v "^   #\|^gate\|^cta\|^   warp" gpurun_out/timeline_base.txt | head -16
echo "== timeline emitw0"; SJB200_LIB=tools/variants/lib_emitw0.so PROBE_KERNEL=4 timeout 200 python tools/timeline.py > gpurun_out/timeline_emitw0.txt 2>&1; grep -v "^   #\|^gate\|^cta\|^   warp" gpurun_out/timeline_emitw0.txt | head -16
echo "== minify / utf8 256 MiB"; timeout 400 python tools/probe_mu.py 2>&1 | tail -1
echo "== e2e through the plug-in"; timeout 600 python tools/e2e_probe.py 2>&1 | grep -v Warning
ls gpurun_out
