#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== minify / utf8 256 MiB"; PROBE_TAG=seamzero timeout 400 python tools/probe_mu.py 2>&1 | tail -1 | cut -c1-330
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "minify or golden or fuzz_multi" 2>&1 | tail -3
