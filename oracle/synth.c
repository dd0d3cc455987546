/*
 * ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/fr.h header).
 *
 * Deterministic synthetic workloads of SURVEY.md 8(d) / BASELINE.json configs 2-4 (the reference
 * ships no generator: its only inputs are circom artefacts, rust/src/aggregator.rs:391-408).
 * Counter-based splitmix64 so that CPU, numpy and CUDA produce identical streams in any order:
 *   mix(z): z=(z^(z>>30))*0xBF58476D1CE4E5B9; z=(z^(z>>27))*0x94D049BB133111EB; z^(z>>31)
 *   word(seed,stream,idx,j) = mix(mix(mix(seed + G*(stream+1)) + G*(idx+1)) + G*(j+1)), G=0x9E3779B97F4A7C15
 * gate g of layer i : stream 0x1000+i; type = word(.,g,0)&1; left = word(.,g,1) mod 2^k_in; right = word(.,g,2) mod 2^k_in
 * field element idx of stream s : limbs word(.,idx,0..3) little-endian, top two bits cleared, minus p if >= p
 */
#include "oracle.h"

#define GOLD 0x9E3779B97F4A7C15ULL
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint64_t orc_synth_word(uint64_t seed, uint64_t stream, uint64_t idx, uint64_t j) {
    uint64_t h = mix64(seed + GOLD * (stream + 1));
    h = mix64(h + GOLD * (idx + 1));
    return mix64(h + GOLD * (j + 1));
}
void orc_synth_gates(uint64_t seed, uint32_t layer, uint32_t k_in, uint32_t n_gates,
                     uint8_t *type, uint32_t *left, uint32_t *right) {
    uint64_t stream = 0x1000 + layer, mask = ((uint64_t)1 << k_in) - 1;
#pragma omp parallel for
    for (uint32_t g = 0; g < n_gates; ++g) {
        type[g] = (uint8_t)(orc_synth_word(seed, stream, g, 0) & 1);
        left[g] = (uint32_t)(orc_synth_word(seed, stream, g, 1) & mask);
        right[g] = (uint32_t)(orc_synth_word(seed, stream, g, 2) & mask);
    }
}
void orc_synth_values(uint64_t seed, uint64_t stream, uint64_t first, uint64_t n, uint8_t *out) {
#pragma omp parallel for
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t v[4];
        for (int j = 0; j < 4; ++j) v[j] = orc_synth_word(seed, stream, first + i, (uint64_t)j);
        v[3] &= 0x3FFFFFFFFFFFFFFFULL;
        if (fr_geq_p(v)) fr_sub_p(v);
        memcpy(out + 32 * i, v, 32);
    }
}
