#!/bin/bash
# tuning aid: builds libsjb200 variants with different -D flags into tools/variants/ (select one with SJB200_LIB=...)
set -e
cd "$(dirname "$0")/.."
SRCS="simdjson_b200/csrc/sjb200_kernels.cu simdjson_b200/csrc/sjb200_capi.cu simdjson_b200/csrc/sjb200_finish.cpp"
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden -shared"
mkdir -p tools/variants
build() { name=$1; shift; nvcc $FLAGS "$@" -o tools/variants/lib_$name.so $SRCS & }
# next to measure: helping look-back with one CTA per SM, two chain warps, lag 3
build help64 -DSJB200_SCAN4_HELP=64
build help1 -DSJB200_SCAN4_HELP=1
build chain2 -DSJB200_SCAN4_CHAIN=2
build counter -DSJB200_SCAN4_COUNTER=1
build counterhelp -DSJB200_SCAN4_COUNTER=1 -DSJB200_SCAN4_HELP=64
build park4 -DSJB200_SCAN4_PARK=4
wait
ls -la tools/variants
