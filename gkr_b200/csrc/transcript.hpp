// MiMC7-91 Fiat-Shamir transcript on the host (see transcript.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

#include "host_field.hpp"

namespace gkr {
void keccak256(const uint8_t *data, size_t len, uint8_t out[32]);
HFr mimc7_hash(const HFr &x, const HFr &key);
bool mimc7_round_constant(unsigned i, HFr *out);      // Montgomery form; false if i >= 91
// r = key; for a in msg: r = r + a + hash(a, r)     (mimc-rs `multi_hash`)
HFr mimc7_multi_hash(const HFr *msg, size_t n, const HFr &key);
// `lanes` independent multi_hash(msg[l][0..n[l]), key 0) evaluations in AVX-512 IFMA lanes (mimc7_lanes.cpp);
// only callable when mimc7_lanes_available()
bool mimc7_lanes_available();
void mimc7_multi_hash_lanes(const HFr *const *msg, const uint32_t *n, HFr *out, int lanes);
}  // namespace gkr
