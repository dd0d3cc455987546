// Host micro-benchmark for the serial transcript hash (the floor of the proving time): times the library's
// mimc7_multi_hash on 3-element messages chained through their result, plus experimental variants selected at
// compile time.  Build/run: tools/hash_bench.sh (tries several compiler flag sets).
#include <chrono>
#include <cstdio>

#include "../gkr_b200/csrc/transcript.hpp"
using namespace gkr;

int main() {
    HFr msg[3] = {hfr_from_u64(123456789), hfr_from_u64(987654321), hfr_from_u64(555)};
    HFr acc = hfr_zero();
    const int N = 400;
    double best = 1e30;
    for (int rep = 0; rep < 400; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < N; ++i) {
            msg[0] = hfr_add(msg[0], acc);
            acc = mimc7_multi_hash(msg, 3, hfr_zero());
        }
        auto t1 = std::chrono::steady_clock::now();
        const double us = std::chrono::duration<double>(t1 - t0).count() * 1e6 / N;
        if (us < best) best = us;
    }
    uint8_t out[32];
    hfr_to_canonical(out, acc);
    printf("%.3f us per 3-element multi_hash  (check %02x%02x%02x%02x)\n", best, out[0], out[1], out[2], out[3]);
    return 0;
}
