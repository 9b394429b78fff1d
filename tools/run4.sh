#!/bin/bash
# round-2 GPU run #4: early look-back, UTF-8 per block, IMAD shifts; e2e with the non-spinning pool; new bench.py
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl gpurun_out/e2e_probe.jsonl
echo "== cpu quota"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config2 or fuzz_small or golden or sharded or kernel_variants or host_pointer" 2>&1 | tail -4
echo "== 64 MiB variants"
for v in base nolb park4 imad imad_park4 utf8unit; do SJB200_LIB=tools/variants/lib_$v.so timeout 200 python tools/probe2.py 2>&1 | tail -1; done
echo "== 1 GiB"
for v in base park4; do PROBE_BYTES=1073741824 SJB200_LIB=tools/variants/lib_$v.so timeout 300 python tools/probe2.py 2>&1 | tail -1; done
echo "== trace base 64 MiB"; SJB200_LIB=tools/variants/lib_trace.so timeout 200 python tools/trace4.py 2>&1 | tee gpurun_out/trace4_64m.txt | tail -22
echo "== timeline base"; SJB200_LIB=tools/variants/lib_base.so PROBE_KERNEL=4 timeout 200 python tools/timeline.py > gpurun_out/timeline_base.txt 2>&1; grep -v "^   #\|^gate\|^cta\|^   warp" gpurun_out/timeline_base.txt | head -16
echo "== e2e through the plug-in"; timeout 600 python tools/e2e_probe.py 2>&1 | grep -v Warning
echo "== bench.py"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench.py reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -c 1200
ls gpurun_out
