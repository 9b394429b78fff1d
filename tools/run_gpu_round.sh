#!/bin/bash
# one gpurun call at the end of a round: parity tests, bench, launch list + full ncu capture of the stage-1 kernel, timeline,
# A/B probes.  Everything lands in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench" ; timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 1900 gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
echo "== ncu launches" ; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1 ; tail -c 200 gpurun_out/ncu_bench.log
echo "== ncu full" ; timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan4 -s 3 -c 1 -o gpurun_out/scan4_full -f python bench.py --steps 3 --warmup 1 > gpurun_out/ncu_full.log 2>&1 ; tail -1 gpurun_out/ncu_full.log
echo "== timeline" ; PROBE_KERNEL=4 timeout 200 python tools/timeline.py > gpurun_out/timeline_scan4.txt 2>&1 ; grep -v "^   #\|^gate\|^cta" gpurun_out/timeline_scan4.txt | tail -12
echo "== probes" ; PROBE_KERNEL=4 timeout 200 python tools/gpu_probe.py 2>&1 | tail -1; PROBE_KERNEL=3 PROBE_KINDS=stage1 timeout 200 python tools/gpu_probe.py 2>&1 | tail -1
echo "== smoke" ; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ls gpurun_out | head -30
