#!/bin/bash
# round-2 GPU run #8: after the clean-up -- all GPU tests, probes, bench, e2e timing breakdown
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl gpurun_out/probe_mu.jsonl gpurun_out/e2e_probe.jsonl
echo "== tests (all)"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== stage1"; for mb in 64 1024; do PROBE_TAG=default_${mb}m PROBE_BYTES=$((mb<<20)) timeout 300 python tools/probe2.py 2>&1 | tail -1; done
echo "== minify / utf8 256 MiB"; timeout 400 python tools/probe_mu.py 2>&1 | tail -1
echo "== e2e stats"; timeout 300 python tools/e2e_stats.py 2>&1 | tail -5
echo "== bench.py"; timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','parity','e2e','gpu_launches')}); print(d['roofline']); print(d['cpu_baseline'])"; tail -3 gpurun_out/bench.err
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls gpurun_out | head -50
