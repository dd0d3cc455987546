#!/bin/bash
# usage: tools/hash_bench.sh   -- builds tools/hash_bench.cpp with several flag sets and runs each (host-only)
cd "$(dirname "$0")/.."
grep -m1 "model name" /proc/cpuinfo
for flags in "-O3" "-O2" "-O3 -fno-tree-vectorize -fno-tree-slp-vectorize" "-O3 -funroll-loops" "-O3 -fschedule-insns -fsched-pressure" "-O3 -mbmi2 -madx" "-O3 -march=native" "-O3 -fno-split-wide-types" "-O1"; do
    g++ $flags -std=c++17 tools/hash_bench.cpp gkr_b200/csrc/transcript.cpp -o /tmp/hash_bench || exit 1
    echo -n "[$flags] "; /tmp/hash_bench
done
