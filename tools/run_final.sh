#!/bin/bash
# end-of-round evidence: the GPU test suite, smoke(), both bench arms, the other configs, the ncu captures
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 600 gpurun_out/r2_bench_reference.json
echo "== bench"; timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 2500 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
: > gpurun_out/r2_configs.jsonl
for cfg in jsonexamples ndjson_1g utf8_minify_256m; do echo "== config $cfg"; timeout 900 python bench.py --config $cfg --steps 10 2>gpurun_out/cfg_$cfg.err | tail -1 >> gpurun_out/r2_configs.jsonl; tail -1 gpurun_out/r2_configs.jsonl | cut -c1-700; done
bash tools/run_ncu.sh r2 2>&1 | tail -70
