#!/bin/bash
# round-2 GPU run #3: ticket prefetch + load reorder (A/B), stagger again, trace, launch / PCIe microbench, utf8v2, fused exchange test
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl gpurun_out/probe_mu.jsonl
echo "== microbench3"; timeout 300 ./tools/microbench3 2>&1 | tee gpurun_out/microbench3.txt
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config2 or fuzz_small or golden or sharded or kernel_variants or host_pointer" 2>&1 | tail -4
echo "== 64 MiB variants"
for v in base park4 stag stag_emit2; do SJB200_LIB=tools/variants/lib_$v.so timeout 200 python tools/probe2.py 2>&1 | tail -1; done
echo "== 1 GiB"
for v in base stag; do PROBE_BYTES=1073741824 SJB200_LIB=tools/variants/lib_$v.so timeout 300 python tools/probe2.py 2>&1 | tail -1; done
echo "== trace base 64 MiB"; SJB200_LIB=tools/variants/lib_trace.so timeout 200 python tools/trace4.py 2>&1 | tee gpurun_out/trace4_64m.txt | tail -22
echo "== trace stag 64 MiB"; SJB200_LIB=tools/variants/lib_trace_stag.so timeout 200 python tools/trace4.py 2>&1 | tee gpurun_out/trace4_stag_64m.txt | tail -22
echo "== minify / utf8 256 MiB (utf8v2 default)"; timeout 400 python tools/probe_mu.py 2>&1 | tail -1
ls gpurun_out
