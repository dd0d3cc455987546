#!/bin/bash
# usage: tools/hash_bench.sh   -- builds tools/hash_bench.cpp with several flag sets / variants and runs each
cd "$(dirname "$0")/.."
grep -m1 "model name" /proc/cpuinfo
for flags in "-O3" "-O3 -mbmi2 -madx" "-O3 -march=native"; do
  for var in ""; do
    g++ $flags $var -std=c++17 tools/hash_bench.cpp gkr_b200/csrc/transcript.cpp -o /tmp/hash_bench || exit 1
    echo -n "[$flags $var] "; /tmp/hash_bench
  done
done
