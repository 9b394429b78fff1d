#!/bin/bash
# final evidence of a round in one gpurun call, most valuable first: ncu launch list + full captures (tools/run_ncu.sh),
# both bench arms, then the other single-GPU configs.  Every step has its own timeout; results land in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
R=${1:-r2}
T0=$(date +%s)
bash tools/run_ncu.sh $R 2>&1 | tail -60
echo "== t=$(( $(date +%s) - T0 )) s: bench"; timeout 300 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 2600 gpurun_out/${R}_bench.json; tail -3 gpurun_out/${R}_bench.err
echo "== t=$(( $(date +%s) - T0 )) s: bench --impl reference"; timeout 300 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err; tail -c 700 gpurun_out/${R}_bench_reference.json
: > gpurun_out/${R}_configs.jsonl
for cfg in ${CONFIGS:-utf8_minify_256m jsonexamples}; do
  echo "== t=$(( $(date +%s) - T0 )) s: config $cfg"
  timeout 300 python bench.py --config $cfg --steps 10 2> gpurun_out/cfg_$cfg.err | tail -1 >> gpurun_out/${R}_configs.jsonl
  tail -1 gpurun_out/${R}_configs.jsonl | cut -c1-900
done
echo "== done t=$(( $(date +%s) - T0 )) s"
