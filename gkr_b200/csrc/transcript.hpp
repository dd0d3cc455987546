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
}  // namespace gkr
