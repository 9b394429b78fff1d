#!/bin/bash
# A/B of the compacted stage-2-lite kernels (tok_stage = 3) against the staged ones (1): parity tests over all variants,
# memcheck of the golden cases on the compacted ones, one bench line each
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T0=$(date +%s)
echo "== tests"; timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "tokens" 2>&1 | tail -8
echo "== t=$(( $(date +%s) - T0 )) s: memcheck (golden, tok_stage=3)"; SJB200_TEST_TOK_STAGE=3 timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k tokens_device_matches_golden > gpurun_out/f4_memcheck3.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/f4_memcheck3.log | head -5
: > gpurun_out/r2_tokens_ab.jsonl
for st in 3 1; do
  echo "== t=$(( $(date +%s) - T0 )) s: bench tokens_64m tok_stage=$st"
  SJB200_TOK_STAGE=$st timeout 120 python bench.py --config tokens_64m --steps 10 2> gpurun_out/f4_bench_$st.err | tail -1 >> gpurun_out/r2_tokens_ab.jsonl
  tail -1 gpurun_out/r2_tokens_ab.jsonl | cut -c1-330; tail -2 gpurun_out/f4_bench_$st.err
done
echo "== t=$(( $(date +%s) - T0 )) s: launch list (3)"; SJB200_TOK_STAGE=3 timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"token_scan|tile_scan|string_write" -c 6 --csv --log-file gpurun_out/r2_tokens3_launches.csv python bench.py --config tokens_64m --steps 3 > gpurun_out/f4_ncu3.log 2>&1; grep -E "token_scan|tile_scan|string_write" gpurun_out/r2_tokens3_launches.csv | awk -F'","' '{print substr($5,1,42), $NF}' | tail -3
echo "== done t=$(( $(date +%s) - T0 )) s"
