#!/bin/bash
# stage-2-lite on the GPU: parity tests, memcheck of the small cases, bench line with and without staging
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T0=$(date +%s)
echo "== tests"; timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k tokens 2>&1 | tail -15
echo "== t=$(( $(date +%s) - T0 )) s: memcheck (golden)"; timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k tokens_device_matches_golden > gpurun_out/f4_memcheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/f4_memcheck.log | head -8
: > gpurun_out/r2_tokens.jsonl
for st in 1 0; do
  echo "== t=$(( $(date +%s) - T0 )) s: bench tokens_64m stage=$st"
  SJB200_TOK_STAGE=$st timeout 300 python bench.py --config tokens_64m --steps 10 2> gpurun_out/f4_bench_$st.err | tail -1 >> gpurun_out/r2_tokens.jsonl
  tail -1 gpurun_out/r2_tokens.jsonl | cut -c1-1500; tail -2 gpurun_out/f4_bench_$st.err
done
echo "== t=$(( $(date +%s) - T0 )) s: launch list"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"token_scan|tile_scan|string_write" -c 12 --csv --log-file gpurun_out/r2_tokens_launches.csv python bench.py --config tokens_64m --steps 3 > gpurun_out/f4_ncu.log 2>&1; grep -E "token_scan|tile_scan|string_write" gpurun_out/r2_tokens_launches.csv | awk -F'","' '{print $5, $NF}' | tail -6
echo "== done t=$(( $(date +%s) - T0 )) s"
