#!/bin/bash
# builds experimental variants of the library: scripts/build_variants.sh NAME "-DFLAG=..." [NAME2 "..."] ...
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O3 -Xptxas=-v $flags \
    -ccbin /usr/bin/g++ -Iinclude -shared -o stereo_matchin_b200/libasw_b200_$name.so stereo_matchin_b200/csrc/asw_api.cu stereo_matchin_b200/csrc/asw_multi.cu > /tmp/build_$name.log 2>&1 &
done
wait
grep -l "error" /tmp/build_*.log
