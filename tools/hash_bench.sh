#!/bin/bash
# usage: tools/hash_bench.sh   -- builds tools/hash_bench.cpp with several flag sets / variants and runs each (host-only).
# The variants are compile-time switches of gkr_b200/csrc/transcript.cpp and host_field.hpp; all give the same hashes.
cd "$(dirname "$0")/.."
grep -m1 "model name" /proc/cpuinfo
for flags in "-O3" "-O3 -DGKR_HOST_SQR=2" "-O3 -DGKR_HOST_SQR=1" "-O3 -DGKR_HASH_CHAIN=1" "-O3 -fschedule-insns -fsched-pressure" \
             "-O3 -mbmi2 -madx" "-O3 -march=native" "-O2"; do
    g++ $flags -std=c++17 tools/hash_bench.cpp gkr_b200/csrc/transcript.cpp -o /tmp/hash_bench || exit 1
    for rep in 1 2; do echo -n "[$flags] "; /tmp/hash_bench; done
done
