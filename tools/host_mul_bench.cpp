// Host micro-benchmark behind the transcript's floor: latency of ONE dependent Montgomery product chain against the
// throughput of 2 / 3 / 4 independent chains (are we bound by the multiplier's latency or by its port?), plus the
// additions and one MiMC round.  Build/run: g++ -O3 -std=c++17 tools/host_mul_bench.cpp -o /tmp/hmb && /tmp/hmb
#include <chrono>
#include <cstdio>

#include "../gkr_b200/csrc/host_field.hpp"
using namespace gkr;

template <class F>
static double best_ns(F f, int inner) {
    double best = 1e30;
    for (int rep = 0; rep < 200; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        f();
        auto t1 = std::chrono::steady_clock::now();
        const double ns = std::chrono::duration<double>(t1 - t0).count() * 1e9 / inner;
        if (ns < best) best = ns;
    }
    return best;
}
int main() {
    const int N = 20000;
    HFr c = hfr_from_u64(0x123456789abcdefULL), x0 = hfr_from_u64(77), x1 = hfr_from_u64(78), x2 = hfr_from_u64(79), x3 = hfr_from_u64(80);
    volatile uint64_t sink = 0;
    const double l1 = best_ns([&] { HFr a = x0; for (int i = 0; i < N; ++i) a = hfr_mul(a, c); sink += a.l[0]; }, N);
    const double l2 = best_ns([&] { HFr a = x0, b = x1; for (int i = 0; i < N; ++i) { a = hfr_mul(a, c); b = hfr_mul(b, c); } sink += a.l[0] + b.l[0]; }, 2 * N);
    const double l3 = best_ns([&] { HFr a = x0, b = x1, d = x2; for (int i = 0; i < N; ++i) { a = hfr_mul(a, c); b = hfr_mul(b, c); d = hfr_mul(d, c); } sink += a.l[0] + b.l[0] + d.l[0]; }, 3 * N);
    const double l4 = best_ns([&] { HFr a = x0, b = x1, d = x2, e = x3; for (int i = 0; i < N; ++i) { a = hfr_mul(a, c); b = hfr_mul(b, c); d = hfr_mul(d, c); e = hfr_mul(e, c); } sink += a.l[0] + b.l[0] + d.l[0] + e.l[0]; }, 4 * N);
    const double sq = best_ns([&] { HFr a = x0; for (int i = 0; i < N; ++i) a = hfr_mul(a, a); sink += a.l[0]; }, N);
    const double ad = best_ns([&] { HFr a = x0; for (int i = 0; i < N; ++i) a = hfr_add(a, c); sink += a.l[0]; }, N);
    const double rd = best_ns([&] {
        HFr h = x0;
        for (int i = 0; i < N; ++i) {
            const HFr t = hfr_add(hfr_add(h, c), x1);
            const HFr t2 = hfr_mul(t, t), t3 = hfr_mul(t2, t), t4 = hfr_mul(t2, t2);
            h = hfr_mul(t3, t4);
        }
        sink += h.l[0]; }, N);
    printf("dependent product %.2f ns | per product with 2 / 3 / 4 independent chains %.2f / %.2f / %.2f ns | dependent square %.2f ns | "
           "dependent add %.2f ns | one MiMC round %.2f ns (= 3 products deep + 2 adds)\n", l1, l2, l3, l4, sq, ad, rd);
    return 0;
}
