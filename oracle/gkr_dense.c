/*
 * ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/fr.h header).  "L1" of the oracle ladder.
 *
 * Dense-table CPU restatement of the reference GKR prover.  The reference works on sparse term
 * lists (rust/src/gkr/poly.rs) and is O(4^k) per round; this file computes the SAME messages,
 * challenges, q, z, r on dense tables following SURVEY.md Appendix B.  Equality with the literal
 * term-list restatement (oracle/l0_reference.py, "L0") is asserted in tests/test_oracle_ladder.py.
 * Reference lines followed:
 *   prover loop, z_0 = 0, r* rule ............ rust/src/gkr/prover.rs:6-96
 *   round messages / hash placement .......... rust/src/gkr/sumcheck.rs:36-156
 *   coefficient order & static lengths ....... rust/src/gkr/poly.rs:388-467
 *   q_i = W restricted to the line b*->c* .... rust/src/gkr/poly.rs:469-500
 *   z_{i+1} = l(b*,c*,r*) .................... rust/src/gkr/poly.rs:538-551
 *   MSB-first wiring / gate indexing ......... rust/src/convert.rs:704-777
 *   forward evaluation, MLE (Moebius) ........ rust/src/convert.rs:787-849, rust/src/gkr/poly.rs:502-536
 *   generic product sumcheck ................. rust/src/gkr/sumcheck.rs:158-214
 * PARITY: equal to the literal restatement (oracle/l0_reference.py) and, through tests/golden/refpy_vectors.json, to the
 * reference's own Python prover run here (tests/test_golden_refpy.py).  UNPINNED against the Rust binary: the reference
 * holds no golden vectors and cannot be built here (no Rust).
 */
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle.h"

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static fr_t *load_table(const uint8_t *bytes, size_t n, int *err) {
    fr_t *t = (fr_t *)malloc(n * sizeof(fr_t) + 32);
    int bad = 0;
#pragma omp parallel for reduction(| : bad)
    for (size_t i = 0; i < n; ++i) bad |= fr_from_bytes(&t[i], bytes + 32 * i) != 0;
    if (bad) *err = 1;
    return t;
}
static void store_table(uint8_t *bytes, const fr_t *t, size_t n) {
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) fr_to_bytes(bytes + 32 * i, t[i]);
}

int orc_fr_binop(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        fr_t x, y, r;
        if (fr_from_bytes(&x, a + 32 * i) || fr_from_bytes(&y, b + 32 * i)) return -1;
        r = op == 0 ? fr_add(x, y) : op == 1 ? fr_sub(x, y) : fr_mul(x, y);
        fr_to_bytes(out + 32 * i, r);
    }
    return 0;
}

/* eq(z, idx) for all idx, MSB-first: bit_1 of idx is its most significant bit
 * (== partial_eval_binary_form of the chi-form term of gate idx, rust/src/gkr/poly.rs:43-62) */
static void eq_table(const fr_t *z, uint32_t k, fr_t *out) {
    out[0] = fr_one();
    size_t size = 1;
    for (uint32_t j = 0; j < k; ++j) {
        for (size_t i = size; i-- > 0;) {
            fr_t hi = fr_mul(out[i], z[j]);
            fr_t lo = fr_sub(out[i], hi);
            out[2 * i] = lo;
            out[2 * i + 1] = hi;
        }
        size *= 2;
    }
}
int orc_eq_table(const uint8_t *zb, uint32_t k, uint8_t *out) {
    int err = 0;
    fr_t *z = load_table(zb, k, &err);
    fr_t *t = (fr_t *)malloc(((size_t)1 << k) * sizeof(fr_t));
    eq_table(z, k, t);
    store_table(out, t, (size_t)1 << k);
    free(z); free(t);
    return err ? -1 : 0;
}

/* values -> monomial coefficients (== get_multi_ext, rust/src/gkr/poly.rs:502-536, as a table:
 * coef[S] multiplies prod_{j in S} x_j, variable j <-> index bit k-j).  Also the OR of the supports
 * of non-zero coefficients (which variables W depends on) and the max total degree. */
static void mobius(fr_t *c, uint32_t k, uint32_t *dep_mask, uint32_t *max_deg) {
    size_t n = (size_t)1 << k;
    for (uint32_t s = 0; s < k; ++s) {
        size_t bit = (size_t)1 << s;
#pragma omp parallel for if (n >= 4096)
        for (size_t i = 0; i < n; ++i)
            if (i & bit) c[i] = fr_sub(c[i], c[i ^ bit]);
    }
    uint32_t mask = 0, deg = 0;
#pragma omp parallel for reduction(| : mask) reduction(max : deg) if (n >= 4096)
    for (size_t i = 0; i < n; ++i)
        if (!fr_is_zero(&c[i])) {
            mask |= (uint32_t)i;
            uint32_t pc = (uint32_t)__builtin_popcountll(i);
            if (pc > deg) deg = pc;
        }
    *dep_mask = mask;
    *max_deg = deg;
}
int orc_mobius(const uint8_t *vals, uint32_t k, uint8_t *coef, uint32_t *dep_mask, uint32_t *max_deg) {
    int err = 0;
    fr_t *c = load_table(vals, (size_t)1 << k, &err);
    mobius(c, k, dep_mask, max_deg);
    store_table(coef, c, (size_t)1 << k);
    free(c);
    return err ? -1 : 0;
}

/* forward evaluation of one layer (rust/src/convert.rs:812-830); absent gates (>= n_gates) are 0 */
int orc_layer_eval(uint32_t n_gates, const uint8_t *type, const uint32_t *left, const uint32_t *right,
                   const uint8_t *in_vals, uint32_t k_in, uint8_t *out_vals, uint32_t k_out) {
    int err = 0;
    size_t n_in = (size_t)1 << k_in, n_out = (size_t)1 << k_out;
    if (n_gates > n_out) return -1;
    fr_t *in = load_table(in_vals, n_in, &err);
    fr_t *out = (fr_t *)calloc(n_out, sizeof(fr_t));
#pragma omp parallel for
    for (uint32_t g = 0; g < n_gates; ++g) {
        if (left[g] >= n_in || right[g] >= n_in) { err = 1; continue; }
        out[g] = type[g] ? fr_mul(in[left[g]], in[right[g]]) : fr_add(in[left[g]], in[right[g]]);
    }
    store_table(out_vals, out, n_out);
    free(in); free(out);
    return err ? -1 : 0;
}

/* coefficients (ascending, k+1 of them) of t -> W(b + t(c-b)): fold variable by variable keeping a
 * polynomial in t per entry.  Equals reduce_multiple_polynomial (rust/src/gkr/poly.rs:469-500) up to
 * the static length rule, which the caller applies. */
static void line_restrict(const fr_t *w, uint32_t k, const fr_t *b, const fr_t *c, fr_t *coef) {
    size_t n = (size_t)1 << k;
    fr_t *cur = (fr_t *)malloc(n * sizeof(fr_t));
    memcpy(cur, w, n * sizeof(fr_t));
    uint32_t deg = 0;                      /* entries of cur are polynomials with deg+1 coefficients */
    size_t cnt = n;                        /* number of entries; entry e coefficient d at cur[d*cnt + e] */
    for (uint32_t j = 0; j < k; ++j) {
        size_t half = cnt / 2;
        fr_t g = fr_sub(c[j], b[j]);
        fr_t *nxt = (fr_t *)malloc(half * (deg + 2) * sizeof(fr_t));
#pragma omp parallel for if (half >= 1024)
        for (size_t e = 0; e < half; ++e) {
            fr_t carry = fr_zero();        /* g * d_{dd-1} */
            for (uint32_t dd = 0; dd <= deg; ++dd) {
                fr_t lo = cur[dd * cnt + e], hi = cur[dd * cnt + e + half];
                fr_t d = fr_sub(hi, lo);
                /* lo(t) + (b + g t) * d(t) */
                fr_t v = fr_add(fr_add(lo, fr_mul(b[j], d)), carry);
                nxt[dd * half + e] = v;
                carry = fr_mul(g, d);
            }
            nxt[(deg + 1) * half + e] = carry;
        }
        free(cur);
        cur = nxt; cnt = half; deg += 1;
    }
    for (uint32_t dd = 0; dd <= k; ++dd) coef[dd] = cur[dd];
    free(cur);
}
int orc_line_restrict(const uint8_t *vals, uint32_t k, const uint8_t *bb, const uint8_t *cb, uint8_t *out) {
    int err = 0;
    fr_t *w = load_table(vals, (size_t)1 << k, &err);
    fr_t *b = load_table(bb, k, &err), *c = load_table(cb, k, &err);
    fr_t *coef = (fr_t *)malloc((k + 1) * sizeof(fr_t));
    line_restrict(w, k, b, c, coef);
    store_table(out, coef, k + 1);
    free(w); free(b); free(c); free(coef);
    return err ? -1 : 0;
}

/* one round of the GKR-specialised sumcheck on dense tables (Appendix B):
 *   g(X) = sum_i (H_lo + X dH)(W_lo + X dW) + (A_lo + X dA),  lo = T[i], hi = T[i + n/2] */
static void round_coeffs(const fr_t *H, const fr_t *W, const fr_t *A, size_t n, fr_t out[3]) {
    size_t half = n / 2;
    fr_t c0 = fr_zero(), c1 = fr_zero(), c2 = fr_zero();
#pragma omp parallel if (half >= 2048)
    {
        fr_t p0 = fr_zero(), p1 = fr_zero(), p2 = fr_zero();
#pragma omp for nowait
        for (size_t i = 0; i < half; ++i) {
            fr_t dH = fr_sub(H[i + half], H[i]), dW = fr_sub(W[i + half], W[i]), dA = fr_sub(A[i + half], A[i]);
            p2 = fr_add(p2, fr_mul(dH, dW));
            p1 = fr_add(p1, fr_add(fr_add(fr_mul(H[i], dW), fr_mul(dH, W[i])), dA));
            p0 = fr_add(p0, fr_add(fr_mul(H[i], W[i]), A[i]));
        }
#pragma omp critical
        { c0 = fr_add(c0, p0); c1 = fr_add(c1, p1); c2 = fr_add(c2, p2); }
    }
    out[0] = c2; out[1] = c1; out[2] = c0;
}
static void fold_table(fr_t *T, size_t n, fr_t r) {
    size_t half = n / 2;
#pragma omp parallel for if (half >= 2048)
    for (size_t i = 0; i < half; ++i) T[i] = fr_add(T[i], fr_mul(r, fr_sub(T[i + half], T[i])));
}

int orc_gkr_prove(uint32_t n_layers, const orc_layer_t *layers, const uint8_t *const *values,
                  uint8_t *msgs, uint8_t *msg_len, uint8_t *chal, uint8_t *q, uint32_t *q_len,
                  uint8_t *z_out, uint8_t *rstar_out, uint8_t *d_coef, uint8_t *in_coef) {
    int err = 0;
    if (n_layers == 0) return -1;
    for (uint32_t i = 0; i < n_layers; ++i) {
        if (layers[i].k_in == 0 || layers[i].k_in > 30 || layers[i].k_out > 30) return -2;  /* v-1 underflow, sumcheck.rs:49 */
        if (layers[i].n_gates == 0 || layers[i].n_gates > ((uint32_t)1 << layers[i].k_out)) return -3;
        if (i + 1 < n_layers && layers[i + 1].k_out != layers[i].k_in) return -4;
    }
    /* z_0 = zeros (prover.rs:16-21) */
    uint32_t kz = layers[0].k_out;
    fr_t *z = (fr_t *)calloc(kz + 1, sizeof(fr_t));
    size_t z_off = 0, round_off = 0, q_off = 0;
    store_table(z_out, z, kz); z_off += kz;

    /* d and input_func: monomial coefficient tables of layer 0 and of the input layer (prover.rs:88,93) */
    {
        uint32_t dm, dg;
        fr_t *c = load_table(values[0], (size_t)1 << layers[0].k_out, &err);
        mobius(c, layers[0].k_out, &dm, &dg);
        store_table(d_coef, c, (size_t)1 << layers[0].k_out); free(c);
        uint32_t kin = layers[n_layers - 1].k_in;
        c = load_table(values[n_layers], (size_t)1 << kin, &err);
        mobius(c, kin, &dm, &dg);
        store_table(in_coef, c, (size_t)1 << kin); free(c);
    }

    for (uint32_t li = 0; li < n_layers; ++li) {
        const orc_layer_t *L = &layers[li];
        uint32_t k = L->k_in;
        size_t N = (size_t)1 << k, NO = (size_t)1 << L->k_out;
        fr_t *Wfull = load_table(values[li + 1], N, &err);
        for (uint32_t g = 0; g < L->n_gates; ++g)
            if (L->left[g] >= N || L->right[g] >= N || L->type[g] > 1) { free(Wfull); free(z); return -5; }

        /* static shape of W_{i+1}: which variables it depends on, its max total degree */
        uint32_t dep_mask, max_deg;
        {
            fr_t *c = (fr_t *)malloc(N * sizeof(fr_t));
            memcpy(c, Wfull, N * sizeof(fr_t));
            mobius(c, k, &dep_mask, &max_deg);
            free(c);
        }
        fr_t *eqz = (fr_t *)malloc(NO * sizeof(fr_t));
        eq_table(z, L->k_out, eqz);

        fr_t *H = (fr_t *)calloc(N, sizeof(fr_t)), *A = (fr_t *)calloc(N, sizeof(fr_t));
        fr_t *W = (fr_t *)malloc(N * sizeof(fr_t));
        fr_t *rs = (fr_t *)malloc(2 * k * sizeof(fr_t));
        fr_t last_hash = fr_zero();

        for (int phase = 0; phase < 2; ++phase) {
            memcpy(W, Wfull, N * sizeof(fr_t));
            memset(H, 0, N * sizeof(fr_t)); memset(A, 0, N * sizeof(fr_t));
            if (phase == 0) {
                /* H[b] = sum_{add: l=b} eqz[g] + sum_{mul: l=b} eqz[g] W[r_g];  A[b] = sum_{add: l=b} eqz[g] W[r_g]
                 * (products on all threads, the scatter of the sums in gate order on one: modular sums are exact, so the
                 * order does not matter for the result -- only the multiplications are worth spreading) */
                fr_t *ew = (fr_t *)malloc((size_t)L->n_gates * sizeof(fr_t));
#pragma omp parallel for if (L->n_gates >= 4096)
                for (uint32_t g = 0; g < L->n_gates; ++g) ew[g] = fr_mul(eqz[g], Wfull[L->right[g]]);
                for (uint32_t g = 0; g < L->n_gates; ++g) {
                    uint32_t l = L->left[g];
                    if (L->type[g] == 0) { H[l] = fr_add(H[l], eqz[g]); A[l] = fr_add(A[l], ew[g]); }
                    else H[l] = fr_add(H[l], ew[g]);
                }
                free(ew);
            } else {
                /* u = b* ; W(u) = fully folded phase-1 W;  equ = eq(u, .) */
                fr_t *equ = (fr_t *)malloc(N * sizeof(fr_t));
                eq_table(rs, k, equ);
                fr_t wu = Wfull[0];
                {   /* W(u): fold a private copy */
                    fr_t *t = (fr_t *)malloc(N * sizeof(fr_t));
                    memcpy(t, Wfull, N * sizeof(fr_t));
                    size_t n = N;
                    for (uint32_t j = 0; j < k; ++j) { fold_table(t, n, rs[j]); n /= 2; }
                    wu = t[0]; free(t);
                }
                fr_t *e = (fr_t *)malloc((size_t)L->n_gates * sizeof(fr_t));
                fr_t *we = (fr_t *)malloc((size_t)L->n_gates * sizeof(fr_t));
#pragma omp parallel for if (L->n_gates >= 4096)
                for (uint32_t g = 0; g < L->n_gates; ++g) {
                    e[g] = fr_mul(eqz[g], equ[L->left[g]]);
                    we[g] = fr_mul(wu, e[g]);
                }
                for (uint32_t g = 0; g < L->n_gates; ++g) {
                    uint32_t r = L->right[g];
                    if (L->type[g] == 0) { H[r] = fr_add(H[r], e[g]); A[r] = fr_add(A[r], we[g]); }
                    else H[r] = fr_add(H[r], we[g]);
                }
                free(e); free(we);
                free(equ);
            }
            size_t n = N;
            for (uint32_t j = 0; j < k; ++j) {
                fr_t c[3];
                round_coeffs(H, W, A, n, c);
                int dep = (dep_mask >> (k - 1 - j)) & 1;          /* variable j+1 <-> index bit k-1-j */
                size_t ridx = round_off + (size_t)phase * k + j;
                uint32_t len = dep ? 3 : 2;
                const fr_t *m = dep ? c : c + 1;
                memset(msgs + ridx * 96, 0, 96);
                store_table(msgs + ridx * 96, m, len);
                msg_len[ridx] = (uint8_t)len;
                fr_t r = mimc7_multi_hash(m, len, fr_zero());
                rs[phase * k + j] = r;
                last_hash = r;
                fr_to_bytes(chal + ridx * 32, r);
                fold_table(H, n, r); fold_table(W, n, r); fold_table(A, n, r);
                n /= 2;
            }
        }
        /* q_i: W on the line b* -> c*, descending, static length 1 + max_deg (poly.rs:469-500) */
        {
            fr_t *coef = (fr_t *)malloc((k + 1) * sizeof(fr_t));
            line_restrict(Wfull, k, rs, rs + k, coef);
            uint32_t len = max_deg + 1;
            for (uint32_t d = len; d <= k; ++d)
                if (!fr_is_zero(&coef[d])) err = 1;       /* cannot happen: deg_t <= max total degree */
            memset(q + q_off * 32, 0, (size_t)(k + 1) * 32);
            for (uint32_t d = 0; d < len; ++d) fr_to_bytes(q + (q_off + d) * 32, coef[len - 1 - d]);
            q_len[li] = len;
            free(coef);
        }
        /* r* = hash of the last message (prover.rs:74-78) == last sumcheck challenge; z_{i+1} (poly.rs:538-551) */
        fr_to_bytes(rstar_out + (size_t)li * 32, last_hash);
        free(z);
        z = (fr_t *)malloc((k + 1) * sizeof(fr_t));
        for (uint32_t j = 0; j < k; ++j) z[j] = fr_add(rs[j], fr_mul(fr_sub(rs[k + j], rs[j]), last_hash));
        store_table(z_out + z_off * 32, z, k); z_off += k;
        round_off += 2 * (size_t)k; q_off += k + 1;
        free(Wfull); free(eqz); free(H); free(A); free(W); free(rs);
    }
    free(z);
    return err ? -1 : 0;
}

/* sumcheck of prod_t T_t(x) over v variables, MSB-first binding, messages as descending
 * coefficients of degree <= n_tables (generic prove_sumcheck, sumcheck.rs:158-214, on dense tables) */
int orc_sumcheck_prod(uint32_t n_tables, uint32_t n_vars, const uint8_t *const *tables,
                      uint8_t *msgs, uint8_t *msg_len, uint8_t *chal, uint8_t *final_vals) {
    if (n_tables < 1 || n_tables > 4 || n_vars < 2 || n_vars > 34) return -2;
    int err = 0;
    size_t N = (size_t)1 << n_vars;
    uint32_t D = n_tables, W = D + 1;
    fr_t *T[4];
    uint32_t depm[4];
    for (uint32_t t = 0; t < D; ++t) {
        T[t] = load_table(tables[t], N, &err);
        /* dependence on the LAST variable only matters (static length of the final round) */
        depm[t] = 0;
        for (size_t i = 0; i < N; i += 2)
            if (!fr_eq(&T[t][i], &T[t][i + 1])) { depm[t] = 1; break; }
    }
    int all_nonzero = 1;
    for (uint32_t t = 0; t < D; ++t) {
        int nz = 0;
        for (size_t i = 0; i < N && !nz; ++i) nz = !fr_is_zero(&T[t][i]);
        all_nonzero &= nz;
    }
    size_t n = N;
    for (uint32_t j = 0; j < n_vars; ++j) {
        size_t half = n / 2;
        fr_t acc[5];
        for (uint32_t d = 0; d < W; ++d) acc[d] = fr_zero();
#pragma omp parallel if (half >= 2048)
        {
            fr_t part[5];
            for (uint32_t d = 0; d < W; ++d) part[d] = fr_zero();
#pragma omp for nowait
            for (size_t i = 0; i < half; ++i) {
                /* product polynomial of the D linear factors lo + X d, ascending coefficients */
                fr_t poly[5];
                poly[0] = fr_one();
                uint32_t deg = 0;
                for (uint32_t t = 0; t < D; ++t) {
                    fr_t lo = T[t][i], dd = fr_sub(T[t][i + half], lo);
                    poly[deg + 1] = fr_mul(poly[deg], dd);
                    for (uint32_t e = deg; e >= 1; --e)
                        poly[e] = fr_add(fr_mul(poly[e], lo), fr_mul(poly[e - 1], dd));
                    poly[0] = fr_mul(poly[0], lo);
                    deg++;
                }
                for (uint32_t d = 0; d < W; ++d) part[d] = fr_add(part[d], poly[d]);
            }
#pragma omp critical
            for (uint32_t d = 0; d < W; ++d) acc[d] = fr_add(acc[d], part[d]);
        }
        /* descending; rounds 1..v-1: zero-sum terms dropped by add_poly => leading zeros stripped
         * (poly.rs:324-327); final round: static length (sumcheck.rs:206-207) */
        uint32_t len;
        if (j + 1 < n_vars) {
            len = W;
            while (len > 1 && fr_is_zero(&acc[len - 1])) --len;
        } else {
            len = 1;
            if (all_nonzero) for (uint32_t t = 0; t < D; ++t) len += depm[t];
        }
        fr_t m[5];
        for (uint32_t d = 0; d < len; ++d) m[d] = acc[len - 1 - d];
        memset(msgs + (size_t)j * W * 32, 0, (size_t)W * 32);
        store_table(msgs + (size_t)j * W * 32, m, len);
        msg_len[j] = (uint8_t)len;
        fr_t r = mimc7_multi_hash(m, len, fr_zero());
        fr_to_bytes(chal + (size_t)j * 32, r);
        for (uint32_t t = 0; t < D; ++t) fold_table(T[t], n, r);
        n = half;
    }
    for (uint32_t t = 0; t < D; ++t) {
        if (final_vals) fr_to_bytes(final_vals + 32 * t, T[t][0]);
        free(T[t]);
    }
    return err ? -1 : 0;
}
