/*
 * ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/fr.h header).  C entry points of the CPU oracle,
 * loaded with ctypes by oracle/oracle.py.  All field elements cross this interface as 32-byte
 * little-endian canonical values (`Fr::to_repr()`, rust/src/gkr/sumcheck.rs:14-21).
 */
#ifndef GKR_ORACLE_H
#define GKR_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "fr.h"

#ifdef __cplusplus
extern "C" {
#endif

/* mimc7.c */
void orc_keccak256(const uint8_t *in, size_t len, uint8_t out[32]);
int orc_mimc7_constant(uint32_t i, uint8_t out[32]);
int orc_mimc7_hash(const uint8_t x[32], const uint8_t k[32], uint8_t out[32]);
int orc_mimc7_multi_hash(const uint8_t *arr, size_t n, const uint8_t key[32], uint8_t out[32]);
fr_t mimc7_hash(fr_t x, fr_t k);
fr_t mimc7_multi_hash(const fr_t *arr, size_t n, fr_t key);

/* synth.c : deterministic synthetic circuits / tables */
uint64_t orc_synth_word(uint64_t seed, uint64_t stream, uint64_t idx, uint64_t j);
void orc_synth_gates(uint64_t seed, uint32_t layer, uint32_t k_in, uint32_t n_gates,
                     uint8_t *type, uint32_t *left, uint32_t *right);
void orc_synth_values(uint64_t seed, uint64_t stream, uint64_t first, uint64_t n, uint8_t *out);

/* gkr_dense.c : building blocks (each mirrors one device kernel) */
int orc_fr_binop(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n);
int orc_eq_table(const uint8_t *z, uint32_t k, uint8_t *out);
int orc_mobius(const uint8_t *vals, uint32_t k, uint8_t *coef, uint32_t *dep_mask, uint32_t *max_deg);
int orc_layer_eval(uint32_t n_gates, const uint8_t *type, const uint32_t *left, const uint32_t *right,
                   const uint8_t *in_vals, uint32_t k_in, uint8_t *out_vals, uint32_t k_out);
int orc_line_restrict(const uint8_t *vals, uint32_t k, const uint8_t *b, const uint8_t *c,
                      uint8_t *coef_ascending /* k+1 */);

/* gkr_dense.c : dense GKR prover, SURVEY.md Appendix B == rust/src/gkr/prover.rs:6-96 */
typedef struct {
    uint32_t k_out, k_in, n_gates;
    const uint8_t *type;   /* 0 = add, 1 = mult */
    const uint32_t *left, *right;
} orc_layer_t;

/* values[i] : 2^{k_i} canonical elements of layer i, i = 0..n_layers (k_{n_layers} = input_k).
 * Outputs (caller-allocated, R = sum_i 2*k_{i+1} rounds):
 *   msgs    R*3*32 bytes, round-major, descending coefficients left-aligned; msg_len[R] in {2,3}
 *   chal    R*32   (sumcheck_r)
 *   q       sum_i (k_{i+1}+1)*32, descending, left aligned per layer slot; q_len[n_layers]
 *   z       sum_{i=0..n_layers} k_i * 32
 *   rstar   n_layers*32
 *   d_coef  2^{k_0}*32 Moebius coefficients of layer 0; in_coef 2^{input_k}*32 of the last layer
 */
int orc_gkr_prove(uint32_t n_layers, const orc_layer_t *layers, const uint8_t *const *values,
                  uint8_t *msgs, uint8_t *msg_len, uint8_t *chal, uint8_t *q, uint32_t *q_len,
                  uint8_t *z, uint8_t *rstar, uint8_t *d_coef, uint8_t *in_coef);

/* generic sumcheck of a product of n_tables multilinear tables (rust/src/gkr/sumcheck.rs:158-214
 * semantics on dense tables): msgs rounds*(n_tables+1)*32 descending left-aligned, msg_len per round
 * (leading zeros stripped in rounds 1..v-1, SURVEY.md Appendix B rule 4), chal rounds*32. */
int orc_sumcheck_prod(uint32_t n_tables, uint32_t n_vars, const uint8_t *const *tables,
                      uint8_t *msgs, uint8_t *msg_len, uint8_t *chal, uint8_t *final_vals);
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
