#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe_mu.jsonl
for v in default u8w8s2 u9w9s2 u4c4s3; do
  if [ $v = default ]; then unset SJB200_LIB; else export SJB200_LIB=$PWD/tools/variants/lib_$v.so; fi
  echo "== $v minify / utf8 256 MiB"; PROBE_TAG=$v timeout 400 python tools/probe_mu.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['tag'], 'utf8', d['utf8'], 'ascii', d['utf8_on_json'], 'minify', d['minify_k4'])"
  PROBE_TAG=${v}_1g PROBE_BYTES=$((1<<30)) timeout 400 python tools/probe_mu.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['tag'], 'utf8', d['utf8'], 'ascii', d['utf8_on_json'], 'minify', d['minify_k4'])"
done
unset SJB200_LIB
echo "== tests quick"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "utf8 or minify" 2>&1 | tail -3
