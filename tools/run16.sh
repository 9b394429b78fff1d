#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/probe2.jsonl gpurun_out/probe_mu.jsonl
for v in default poll4 poll16; do
  if [ $v = default ]; then unset SJB200_LIB; else export SJB200_LIB=$PWD/tools/variants/lib_$v.so; fi
  echo "== $v stage1"; for mb in 64 256 1024; do PROBE_TAG=${v}_${mb}m PROBE_BYTES=$((mb<<20)) timeout 300 python tools/probe2.py 2>&1 | tail -1 | cut -c1-260; done
  echo "== $v minify / utf8 256 MiB"; PROBE_TAG=$v timeout 400 python tools/probe_mu.py 2>&1 | tail -1 | cut -c1-330
done
