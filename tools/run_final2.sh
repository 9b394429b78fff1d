#!/bin/bash
# last GPU call of the round: regression of what changed since tools/run_evidence.sh (C ABI additions, stage-2-lite),
# smoke(), the default bench line, the stage-2-lite bench line / launch list / ncu --set full summary
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T0=$(date +%s)
echo "== tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tokens or golden or capacity_and_empty or unaligned" 2>&1 | tail -6
echo "== t=$(( $(date +%s) - T0 )) s: smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== t=$(( $(date +%s) - T0 )) s: bench"; timeout 200 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 1500 gpurun_out/r2_bench.json; tail -2 gpurun_out/r2_bench.err
: > gpurun_out/r2_tokens.jsonl
for st in 1 0; do
  echo "== t=$(( $(date +%s) - T0 )) s: bench tokens_64m stage=$st"
  SJB200_TOK_STAGE=$st timeout 200 python bench.py --config tokens_64m --steps 10 2> gpurun_out/f4_bench_$st.err | tail -1 >> gpurun_out/r2_tokens.jsonl
  tail -1 gpurun_out/r2_tokens.jsonl | cut -c1-420; tail -2 gpurun_out/f4_bench_$st.err
done
echo "== t=$(( $(date +%s) - T0 )) s: launch list"; timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"token_scan|tile_scan|string_write" -c 12 --csv --log-file gpurun_out/r2_tokens_launches.csv python bench.py --config tokens_64m --steps 3 > gpurun_out/f4_ncu.log 2>&1; grep -E "token_scan|tile_scan|string_write" gpurun_out/r2_tokens_launches.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -3
echo "== t=$(( $(date +%s) - T0 )) s: ncu full"; bash tools/run_ncu_tokens.sh r2 2>&1 | tail -25 | cut -c1-200
echo "== done t=$(( $(date +%s) - T0 )) s"
