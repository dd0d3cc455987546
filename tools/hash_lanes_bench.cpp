// Host check + micro-benchmark of the AVX-512 IFMA lane hash (gkr_b200/csrc/mimc7_lanes.cpp) against the scalar
// transcript hash.  Build/run:
//   g++ -O3 -std=c++17 -mavx512f -mavx512ifma tools/hash_lanes_bench.cpp gkr_b200/csrc/mimc7_lanes.cpp \
//       gkr_b200/csrc/transcript.cpp -o build/hash_lanes_bench && build/hash_lanes_bench
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>

#include "../gkr_b200/csrc/transcript.hpp"
using namespace gkr;

static uint64_t rng_state = 88172645463325252ULL;
static uint64_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}
static HFr rand_fr() {
    HFr a{{rnd(), rnd(), rnd(), rnd() >> 3}};
    while (hf::geq_p(a.l)) hf::sub_p(a.l);
    return a;
}

int main() {
    if (!mimc7_lanes_available()) {
        printf("no avx512ifma on this CPU\n");
        return 0;
    }
    int bad = 0;
    for (int trial = 0; trial < 200; ++trial) {
        const int lanes = 1 + (int)(rnd() % 16);
        HFr store[16][4];
        const HFr *ptr[16];
        uint32_t n[16];
        for (int l = 0; l < lanes; ++l) {
            n[l] = (uint32_t)(rnd() % 5);        // 0..4 elements (0 = empty message -> key 0)
            for (uint32_t e = 0; e < n[l]; ++e) {
                store[l][e] = rand_fr();
                if (trial % 7 == 0 && e == 0) store[l][e] = hfr_zero();
                if (trial % 11 == 0 && e == 1) store[l][e] = HFr{{hf::P[0] - 1, hf::P[1], hf::P[2], hf::P[3]}};
            }
            ptr[l] = store[l];
        }
        HFr out[16];
        mimc7_multi_hash_lanes(ptr, n, out, lanes);
        for (int l = 0; l < lanes; ++l) {
            const HFr want = mimc7_multi_hash(store[l], n[l], hfr_zero());
            if (!hfr_eq(want, out[l])) ++bad;
        }
    }
    printf("lane hash vs scalar: %s (%d mismatches)\n", bad ? "FAIL" : "ok", bad);
    for (int lanes : {1, 8, 16}) {
        HFr store[16][3];
        const HFr *ptr[16];
        uint32_t n[16];
        for (int l = 0; l < 16; ++l) {
            for (int e = 0; e < 3; ++e) store[l][e] = rand_fr();
            ptr[l] = store[l];
            n[l] = 3;
        }
        HFr out[16];
        double best = 1e30;
        for (int rep = 0; rep < 50; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < 50; ++i) {
                mimc7_multi_hash_lanes(ptr, n, out, lanes);
                for (int l = 0; l < lanes; ++l) store[l][0] = out[l];
            }
            auto t1 = std::chrono::steady_clock::now();
            const double us = std::chrono::duration<double>(t1 - t0).count() * 1e6 / 50;
            if (us < best) best = us;
        }
        printf("{\"lanes\": %d, \"us_per_call\": %.2f, \"us_per_hash\": %.3f}\n", lanes, best, best / lanes);
    }
    HFr msg[3] = {rand_fr(), rand_fr(), rand_fr()};
    double best = 1e30;
    for (int rep = 0; rep < 50; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 50; ++i) msg[0] = mimc7_multi_hash(msg, 3, hfr_zero());
        auto t1 = std::chrono::steady_clock::now();
        const double us = std::chrono::duration<double>(t1 - t0).count() * 1e6 / 50;
        if (us < best) best = us;
    }
    printf("{\"scalar_us_per_hash\": %.3f}\n", best);
    return bad != 0;
}
