#!/bin/bash
# one gpurun call: parity tests, bench, A/B of the two stage-1 kernels, timeline, micro-probes, ncu.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== probe k4" ; PROBE_KERNEL=4 PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -2
echo "== probe k3" ; PROBE_KERNEL=3 PROBE_KINDS=stage1 timeout 300 python tools/gpu_probe.py 2>&1 | tail -2
echo "== timeline k4" ; PROBE_KERNEL=4 timeout 300 python tools/timeline.py > gpurun_out/timeline_k4.txt 2>&1 ; tail -22 gpurun_out/timeline_k4.txt
echo "== tests" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 1500 gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
echo "== microbench2" ; timeout 120 tools/microbench2 > gpurun_out/microbench2.txt 2>&1 ; cat gpurun_out/microbench2.txt
echo "== ldmatrix" ; timeout 60 tools/ldmatrix_probe > gpurun_out/ldmatrix.txt 2>&1 ; head -34 gpurun_out/ldmatrix.txt
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 > gpurun_out/ncu_bench.log 2>&1 ; tail -2 gpurun_out/ncu_bench.log
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan4_kernel -s 2 -c 1 -o gpurun_out/scan4_full -f python bench.py --steps 3 --warmup 1 > gpurun_out/ncu_full.log 2>&1 ; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | head -30
