#!/bin/bash
# one gpurun call at the end of a round: parity tests, bench, launch list + full ncu capture of the stage-1 kernel, timelines,
# the full-size configurations.  Everything lands in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== tests" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 1800 gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
echo "== probes" ; for d in 1 0; do PROBE_KERNEL=4 PROBE_DEFERRED=$d timeout 300 python tools/gpu_probe.py 2>&1 | tail -1; done; PROBE_KERNEL=3 PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -1
echo "== timelines" ; PROBE_KERNEL=4 PROBE_DEFERRED=1 timeout 300 python tools/timeline.py > gpurun_out/timeline_deferred.txt 2>&1 ; PROBE_KERNEL=4 PROBE_DEFERRED=0 timeout 300 python tools/timeline.py > gpurun_out/timeline_pipelined.txt 2>&1 ; grep -v "^   #\|^gate\|^cta" gpurun_out/timeline_deferred.txt | tail -12
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1 ; tail -c 300 gpurun_out/ncu_bench.log
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan4 -s 3 -c 1 -o gpurun_out/scan4_full -f python bench.py --steps 3 --warmup 1 > gpurun_out/ncu_full.log 2>&1 ; tail -1 gpurun_out/ncu_full.log
echo "== full-size configurations" ; timeout 900 python tools/big_probe.py 2>&1 | tail -8
ls -la gpurun_out | head -30
