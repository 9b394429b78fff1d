#!/bin/bash
# round-2 GPU run #5: emit warps, full GPU test-suite, e2e with streaming-store staging
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl gpurun_out/e2e_probe.jsonl
echo "== tests (all)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== 64 MiB variants"
for v in base emitw0 emitw3 emitw7_nolb warps8 warps8_emitw3; do SJB200_LIB=tools/variants/lib_$v.so timeout 200 python tools/probe2.py 2>&1 | tail -1; done
echo "== 1 GiB"
for v in base emitw0 warps8_emitw3; do PROBE_BYTES=1073741824 SJB200_LIB=tools/variants/lib_$v.so timeout 300 python tools/probe2.py 2>&1 | tail -1; done
echo "== trace base 64 MiB"; SJB200_LIB=tools/variants/lib_trace.so timeout 200 python tools/trace4.py 2>&1 | tee gpurun_out/trace4_64m.txt | tail -22
echo "== timeline base"; SJB200_LIB=tools/variants/lib_base.so PROBE_KERNEL=4 timeout 200 python tools/timeline.py > gpurun_out/timeline_base.txt 2>&1; grep -v "^   #\|^gate\|^cta\|^   warp" gpurun_out/timeline_base.txt | head -16
echo "== e2e through the plug-in"; timeout 600 python tools/e2e_probe.py 2>&1 | grep -v Warning
echo "== bench.py"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','parity','e2e','gpu_launches')}); print(d['roofline']); print(d['cpu_baseline'])"; tail -3 gpurun_out/bench.err
echo "== bench.py --check"; timeout 600 python bench.py --check 2>&1 | tail -2
echo "== configs"; for c in jsonexamples utf8_minify_256m; do timeout 900 python bench.py --config $c --steps 10 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -c 1500 gpurun_out/bench_$c.json; tail -2 gpurun_out/bench_$c.err; done
ls gpurun_out
