#!/bin/bash
# builds the stand-alone tap-loop probe (stereo_matchin_b200/csrc/ubench/taploop.cu -> scripts/taploop.bin)
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas=-v -ccbin /usr/bin/g++ \
  -o scripts/taploop.bin stereo_matchin_b200/csrc/ubench/taploop.cu > /tmp/build_taploop.log 2>&1 || { tail -20 /tmp/build_taploop.log; exit 1; }
grep -E "spill|registers" /tmp/build_taploop.log | paste - - | awk '{print $0}' | sort | uniq -c | sort -rn | head -40
