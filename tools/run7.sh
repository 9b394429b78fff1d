#!/bin/bash
# round-2 GPU run #7: emit warps + parked masks in an L2-resident ring (deep decoupling from the chain)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl gpurun_out/probe_mu.jsonl
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "epilogue or config2 or fuzz_small or golden or kernel_variants or scan4_experimental or sharded" 2>&1 | tail -4
echo "== 64 MiB variants"
for v in base emitw8 gpark12 emitw8_gpark16 emitw0 gpark0; do SJB200_LIB=tools/variants/lib_$v.so timeout 200 python tools/probe2.py 2>&1 | tail -1; done
echo "== 1 GiB"
for v in base emitw8 emitw0; do PROBE_BYTES=1073741824 SJB200_LIB=tools/variants/lib_$v.so timeout 300 python tools/probe2.py 2>&1 | tail -1; done
echo "== 16/32/128/256 MiB base"
for mb in 16 32 128 256; do PROBE_PARITY=0 PROBE_TAG=base_${mb}m PROBE_BYTES=$((mb<<20)) SJB200_LIB=tools/variants/lib_base.so timeout 200 python tools/probe2.py 2>&1 | tail -1; done
echo "== trace base 64 MiB"; SJB200_LIB=tools/variants/lib_trace.so timeout 200 python tools/trace4.py 2>&1 | tee gpurun_out/trace4_64m.txt | tail -22
echo "== timeline base"; SJB200_LIB=tools/variants/lib_base.so PROBE_KERNEL=4 timeout 200 python tools/timeline.py > gpurun_out/timeline_base.txt 2>&1; grep -v "^   #\|^gate\|^cta\|^   warp" gpurun_out/timeline_base.txt | head -16
echo "== minify / utf8 256 MiB"; timeout 400 python tools/probe_mu.py 2>&1 | tail -1
ls gpurun_out
