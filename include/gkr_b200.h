/*
 * gkr_b200 -- C ABI of the B200-native GKR prover hot path.
 *
 * Drop-in boundary for the reference's prover seam
 *     prover::prove(&GKRCircuit<S>, &Input<S>) -> Proof<S>         rust/src/gkr/prover.rs:6-9
 * called per sub-circuit from rayon workers at rust/src/aggregator.rs:352-355 and :413-416.
 * The reference has no FFI; its seam passes sparse term lists (rust/src/gkr.rs:21-56) whose
 * construction is a 3^k blow-up (get_multi_ext, rust/src/gkr/poly.rs:502-536).  The ABI therefore
 * binds one step earlier, where the reference still holds dense data:
 *     IntermediateLayer{node_types, operand_index}                  rust/src/convert.rs:103-106, consumed :704-777
 *     w_values (dense layer values)                                 rust/src/convert.rs:793-831
 * and returns every field of Proof<S> (rust/src/gkr.rs:8-19).  INTEGRATION.md shows the Rust
 * (cc + bindgen) side a maintainer would add.
 *
 * Conventions
 *  - gkr_fr is the canonical value, 32 bytes little-endian == Fr::to_repr() (sumcheck.rs:14-21,
 *    file_utils.rs:20-24).  Montgomery form never crosses the boundary.  Inputs >= p are an error.
 *  - Gate g of layer i is output index g; indices are MSB-first bit strings (convert.rs:721-728):
 *    variable 1 of a layer is the most significant index bit.
 *  - Every function returns 0 on success or a negative gkr_status; gkr_last_error() gives a
 *    thread-local message.  Nothing unwinds or aborts across the boundary.
 *  - There is NO CPU fallback: without a CUDA device every entry point that computes fails with
 *    GKR_ERR_CUDA.
 *  - A gkr_ctx is bound to one device + one stream and must not be used from two threads at once;
 *    distinct contexts are fully concurrent (the reference proves sub-circuits concurrently,
 *    aggregator.rs:353,414).  A gkr_circuit is immutable after creation.
 */
#ifndef GKR_B200_H
#define GKR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t l[8]; } gkr_fr;

typedef enum {
    GKR_OK = 0,
    GKR_ERR_INVALID = -1,     /* bad argument / unsupported shape (e.g. k_in = 0: sumcheck.rs:49 underflows) */
    GKR_ERR_CUDA = -2,        /* CUDA runtime error or no device */
    GKR_ERR_OOM = -3,
    GKR_ERR_RANGE = -4,       /* a field element >= p was supplied */
    GKR_ERR_TRANSCRIPT = -5,  /* user transcript callback failed */
    GKR_ERR_COMM = -6,        /* multi-GPU exchange failed */
    GKR_ERR_INTERNAL = -7
} gkr_status;

typedef struct gkr_ctx gkr_ctx;
typedef struct gkr_circuit gkr_circuit;
typedef struct gkr_witness gkr_witness;

/* ---- context ------------------------------------------------------------------------------- */
int gkr_ctx_create(int device, gkr_ctx **out);
void gkr_ctx_destroy(gkr_ctx *ctx);
/* options: "paranoid" = 1 makes every round also accumulate g(1) on the device and checks
 * g_j(0) + g_j(1) == g_{j-1}(r_{j-1}) on the host (default 0: g(1) is derived from the running claim);
 * "lookahead" = 0 disables the look-ahead rounds (default 1: while the host hashes a round message the device
 * already computes the next message as a quadratic in the pending challenge, so a round costs max(hash, device));
 * "prelaunch" = 0 disables launching the small-table rounds of a phase ahead of their challenges (default 1:
 * those kernels wait up to ~30 s for each challenge in a mapped command block, so a transcript callback must
 * not block for longer than that.  The library itself makes no implicitly synchronising CUDA call (cudaFree ...)
 * while such kernels wait -- all its device memory is recycled through per-context pools -- but a host that calls
 * cudaFree / cudaDeviceSynchronize on the same device from other threads during a proof can stall behind them;
 * such hosts should set "prelaunch" = 0) */
int gkr_ctx_set_option(gkr_ctx *ctx, const char *name, int value);
/* the cudaStream_t every kernel of this context is launched on (for CUDA-event timing by the caller) */
void *gkr_ctx_stream(gkr_ctx *ctx);
/* block until all work enqueued by this context has finished */
int gkr_ctx_sync(gkr_ctx *ctx);
const char *gkr_last_error(void);
const char *gkr_version(void);

/* ---- Fiat-Shamir ------------------------------------------------------------------------------
 * challenge(user, msg, n, r_out) must return 0 and write the challenge for the round message
 * msg[0..n) (descending coefficients).  NULL transcript => built-in MiMC7, 91 rounds, key 0:
 * r = multi_hash(msg, 0)   (mimc-rs; sumcheck.rs:83-85,128-130,151-153; prover.rs:74-78).
 * Invoked once per round, synchronously, on the calling thread. */
typedef struct {
    void *user;
    int (*challenge)(void *user, const gkr_fr *msg, uint32_t n, gkr_fr *r_out);
} gkr_transcript;

/* built-in transcript, exposed for hosts/tests: out = multi_hash(msg[0..n), key) ; hash(x, key) */
int gkr_mimc7_multi_hash(const gkr_fr *msg, uint32_t n, const gkr_fr *key, gkr_fr *out);
int gkr_mimc7_hash(const gkr_fr *x, const gkr_fr *key, gkr_fr *out);
/* round constant c_i of MiMC7-91 (c_0 = 0, c_i = keccak256^{i+1}("mimc") mod p), i < 91: what circomlib's MiMC7 template
   (rust/t.circom:9) bakes into its constraints; used to build that constraint system without circom */
int gkr_mimc7_round_constant(uint32_t i, gkr_fr *out);

/* ---- circuit (replaces the add_i/mult_i/wire emission of convert.rs:704-777) ------------------- */
typedef struct {
    uint32_t k_out;           /* layer i has 2^k_out output slots (Layer.k, gkr.rs:36) */
    uint32_t k_in;            /* k of layer i+1 (or input_k for the last layer); must be >= 1 */
    uint32_t n_gates;         /* 1..2^k_out; slots >= n_gates are absent gates with value 0 */
    const uint8_t *type;      /* 0 = Add, 1 = Mult (NodeType, convert.rs:815-826) */
    const uint32_t *left;     /* operand_index.0, < 2^k_in */
    const uint32_t *right;    /* operand_index.1, < 2^k_in */
} gkr_layer_desc;

/* layers[0] is the output layer; layers[i].k_in == layers[i+1].k_out.  Copies everything it needs
 * (the caller may free its arrays on return) and builds the by-left / by-right CSR on the device. */
int gkr_circuit_create(gkr_ctx *ctx, uint32_t n_layers, const gkr_layer_desc *layers, gkr_circuit **out);
/* returns the device arrays to the creating context's pool: destroy circuits BEFORE their gkr_ctx */
void gkr_circuit_destroy(gkr_circuit *c);

/* ---- witness (replaces calculate_input, convert.rs:787-849) -------------------------------------- */
/* all layer tables supplied by the host: layer_values[i] has 2^{k_i} elements, i = 0..n_layers */
int gkr_witness_create(gkr_ctx *ctx, const gkr_circuit *c, const gkr_fr *const *layer_values, gkr_witness **out);
/* only the input layer supplied (2^{input_k} elements); the layers are evaluated on the device */
int gkr_witness_eval(gkr_ctx *ctx, const gkr_circuit *c, const gkr_fr *input_values, gkr_witness **out);
/* copy layer i (canonical values) back to the host */
int gkr_witness_layer(gkr_ctx *ctx, const gkr_witness *w, uint32_t layer, gkr_fr *out);
/* returns the tables to the creating context's pool: destroy witnesses BEFORE their gkr_ctx */
void gkr_witness_destroy(gkr_witness *w);

/* ---- proof == Proof<S> (gkr.rs:8-19), flat ------------------------------------------------------- */
typedef struct {
    uint32_t n_layers;         /* circuit.depth() */
    uint32_t depth;            /* Proof.depth = n_layers + 1 (prover.rs:92) */
    const uint32_t *k;         /* [n_layers+1]  Proof.k */
    uint64_t n_rounds;         /* sum_i 2*k_{i+1} */
    const uint64_t *round_off; /* [n_layers+1] prefix: rounds of layer i are round_off[i]..round_off[i+1] */
    const uint8_t *msg_len;    /* [n_rounds] 2 or 3 */
    const gkr_fr *msgs;        /* [n_rounds][3] sumcheck_proofs: descending coefficients, left aligned */
    const gkr_fr *chal;        /* [n_rounds]    sumcheck_r */
    const uint64_t *q_off;     /* [n_layers+1] prefix with stride k_{i+1}+1 */
    const uint32_t *q_len;     /* [n_layers] */
    const gkr_fr *q;           /* q[q_off[i] .. +q_len[i]) descending */
    const uint64_t *z_off;     /* [n_layers+2] prefix: z_i has k_i elements */
    const gkr_fr *z;
    const gkr_fr *r;           /* [n_layers]  r*_i */
    uint64_t d_len;            /* 2^{k_0} */
    const gkr_fr *d_coef;      /* Proof.d as a dense monomial table: entry S multiplies prod_{j in S} x_j,
                                  variable j <-> index bit k-j; zero entries are absent terms */
    uint64_t input_len;        /* 2^{input_k} */
    const gkr_fr *input_coef;  /* Proof.input_func, same encoding */
} gkr_proof;

int gkr_prove(gkr_ctx *ctx, const gkr_circuit *c, const gkr_witness *w, const gkr_transcript *t, gkr_proof **out);
void gkr_proof_free(gkr_proof *p);

/* ---- batches of independent proofs (the reference proves the sub-circuits of one input under rayon par_iter,
 * rust/src/aggregator.rs:352-355 and :413-416; one OS thread and one serial transcript per proof) -------------------
 * A gkr_batch owns n_threads worker threads (0 = one per allowed CPU, at most 16), each pinned to a CPU of its own and
 * advancing `lanes` proofs (0 = default: 4..16 depending on n_threads; at most 32) in lockstep, so that one round message of
 * every proof is hashed per SIMD call (MiMC7 multi_hash, 8 or 16 lanes) and the device work of one proof overlaps
 * the host work of the others.  Every proof is bit-identical to gkr_prove's.
 * gkr_batch_load uploads the circuits and evaluates the witnesses on the device (the job descriptions are not
 * referenced after it returns; a new load replaces the previous one); gkr_batch_prove proves every loaded job and
 * may be called repeatedly.  proofs_out[n_jobs] receives the proofs in job order (free each with gkr_proof_free), or
 * is NULL to discard them; seconds_out (optional) = wall clock of the proving alone, from the moment all workers
 * stand ready to the last proof (what aggregator.rs:406-418 times).  The calling thread blocks meanwhile. */
typedef struct gkr_batch gkr_batch;
typedef struct {
    uint32_t n_layers;
    const gkr_layer_desc *layers;
    const gkr_fr *input_values;       /* 2^(k_in of the last layer) values of the input layer */
} gkr_job;
int gkr_batch_create(int device, int n_threads, int lanes, gkr_batch **out);
int gkr_batch_load(gkr_batch *b, const gkr_job *jobs, size_t n_jobs);
int gkr_batch_prove(gkr_batch *b, gkr_proof **proofs_out, double *seconds_out);
int gkr_batch_set_option(gkr_batch *b, const char *name, int value);   /* gkr_ctx_set_option on every context of the batch */
int gkr_batch_threads(const gkr_batch *b);
int gkr_batch_lanes(const gkr_batch *b);
int gkr_batch_simd_hash(void);        /* 1 when this CPU runs the AVX-512 IFMA lane hash */
void gkr_batch_destroy(gkr_batch *b);
/* create + load + prove + destroy */
int gkr_prove_many(int device, const gkr_job *jobs, size_t n_jobs, int n_threads, int lanes, gkr_proof **proofs_out);
/* `count` independent multi_hash(msg_i, key 0) evaluations: message i = msgs[i * stride .. + n[i]) (n[i] <= stride).
 * Uses the lane hash when the CPU has it, else the scalar one; the results are identical (host only). */
int gkr_mimc7_multi_hash_many(const gkr_fr *msgs, const uint32_t *n, uint32_t stride, uint32_t count, gkr_fr *out);

/* ---- verifier (complete check of the reference protocol; the reference's own verifiers are partial:
 * gkr-verifier-circuits/circom/circom/verifier.circom:39-71 never evaluates add_i/mult_i nor the hashes,
 * python/gkr.py:202-231 never ties the last sumcheck claim to q) -----------------------------------------
 * Checks, per layer: g_j(0)+g_j(1) == claim, r_j == challenge(g_j), claim <- g_j(r_j);
 * final claim == add_i(z,b*,c*)(q(0)+q(1)) + mult_i(z,b*,c*) q(0) q(1)  (predicates evaluated on the device);
 * r*_i == challenge(last message); z_{i+1} == b* + r*(c*-b*); claim <- q_i(r*_i);
 * and at the end W_depth(z_depth) == claim on the supplied input layer (device MLE evaluation).
 * The starting claim is W_0(0..0) = d_coef[0].  `proof` may come from anywhere: only its arrays are read.
 * *accepted = 1/0; on rejection gkr_last_error() names the first failing check.  transcript NULL => MiMC7. */
int gkr_verify(gkr_ctx *ctx, const gkr_circuit *c, const gkr_proof *proof, const gkr_fr *input_values,
               const gkr_transcript *t, int *accepted);

/* ---- standalone sumcheck of a product of 3 multilinear tables (prove_sumcheck, sumcheck.rs:158-214) -
 * tables: n_tables (= 3) pointers to 2^n_vars canonical elements in HOST memory (on_device = 0) or
 * Montgomery-form DEVICE buffers created by gkr_dev_table_* (on_device = 1).
 * msgs: [n_vars][4] descending left-aligned, msg_len[n_vars], chal[n_vars], final_vals[n_tables]. */
int gkr_sumcheck_prod(gkr_ctx *ctx, uint32_t n_tables, uint32_t n_vars, const void *const *tables, int on_device,
                      const gkr_transcript *t, gkr_fr *msgs, uint8_t *msg_len, gkr_fr *chal, gkr_fr *final_vals);

/* ---- multi-GPU: one process (and one gkr_ctx) per GPU, NCCL over NVLink -----------------------------
 * Rank 0 calls gkr_comm_unique_id and ships the bytes to the other ranks (MPI, torch.distributed, a file);
 * every rank then calls gkr_comm_init (collective).  n_ranks must be a power of two. */
#define GKR_COMM_ID_BYTES 256
int gkr_comm_unique_id(uint8_t out[GKR_COMM_ID_BYTES]);
int gkr_comm_init(gkr_ctx *ctx, int n_ranks, int rank, const uint8_t id[GKR_COMM_ID_BYTES]);
void gkr_comm_destroy(gkr_ctx *ctx);
/* The same communicator without NCCL: the per-round exchange of this library never goes through NCCL anyway (a block of
 * pinned host memory shared by the ranks), so all a rank needs is the name of the POSIX shared-memory object -- the same
 * string "/..." (< 64 bytes, unique per job) on every rank, handed over by whatever starts the ranks.  Rank 0 creates
 * the object, the others wait for it (60 s), all meet inside the block, rank 0 unlinks the name.  Ranks may share a
 * device (a multi-process job on a single-GPU box), which an NCCL communicator refuses.  Collective. */
int gkr_comm_init_shared(gkr_ctx *ctx, int n_ranks, int rank, const char *shm_name);
/* The same communicator with every rank inside ONE process: creates n_ranks contexts (rank r on device_ids[r]; ranks may
 * share a device) whose mailboxes address each other directly.  Each context is then driven by its own host thread;
 * the sharded entry points are collective over the group.  No NCCL involved.  Destroy every context with
 * gkr_ctx_destroy (in any order, after all ranks have finished). */
int gkr_comm_create(int n_ranks, const int *device_ids, gkr_ctx **ctxs_out);
/* Table-sharded product sumcheck over n_vars GLOBAL variables.  The tables are split on the variables
 * bound LAST, i.e. the low log2(n_ranks) index bits: rank p holds entries idx = i * n_ranks + p as
 * local_tables[t][i], i < 2^(n_vars - log2 n_ranks), Montgomery-form device buffers (gkr_dev_table_*).
 * Per round each rank reduces its shard, the partial sums are all-gathered (NCCL) and added on every rank,
 * every rank derives the same challenge and folds locally; the last log2(n_ranks) rounds run on the
 * gathered single entries.  Every rank returns the complete, identical proof. */
int gkr_sumcheck_prod_sharded(gkr_ctx *ctx, uint32_t n_tables, uint32_t n_vars, const void *const *local_tables,
                              const gkr_transcript *t, gkr_fr *msgs, uint8_t *msg_len, gkr_fr *chal, gkr_fr *final_vals);

/* device-resident tables for benchmarks at sizes that should not cross PCIe (BASELINE.json config 4) */
int gkr_dev_table_synth(gkr_ctx *ctx, uint64_t seed, uint64_t stream, uint64_t n, void **dev_table_out);
/* element i = element (first + i * stride) of the synthetic stream (shards of a larger table) */
int gkr_dev_table_synth_strided(gkr_ctx *ctx, uint64_t seed, uint64_t stream, uint64_t first, uint64_t stride, uint64_t n,
                                void **dev_table_out);
int gkr_dev_table_upload(gkr_ctx *ctx, const gkr_fr *host_values, uint64_t n, void **dev_table_out);
int gkr_dev_table_download(gkr_ctx *ctx, const void *dev_table, uint64_t n, gkr_fr *host_out);
/* MLE evaluation of a device table of 2^n_vars entries at `point` (point[0] pairs with the most significant index bit, the
 * order the sumcheck rounds bind the variables): eq table from the point, one dot product.  This is the verifier's final
 * check of gkr_sumcheck_prod: final_vals[i] == eval(table i, challenges). */
int gkr_dev_table_eval(gkr_ctx *ctx, const void *dev, uint32_t n_vars, const gkr_fr *point, gkr_fr *out);
void gkr_dev_table_free(gkr_ctx *ctx, void *dev_table);

/* ---- building blocks, exposed so that each kernel can be checked against the oracle --------------- */
int gkr_fr_binop(gkr_ctx *ctx, int op /*0 add,1 sub,2 mul*/, const gkr_fr *a, const gkr_fr *b, gkr_fr *out, uint64_t n);
int gkr_eq_table(gkr_ctx *ctx, const gkr_fr *z, uint32_t k, gkr_fr *out);
int gkr_mobius(gkr_ctx *ctx, const gkr_fr *values, uint32_t k, gkr_fr *coef_out, uint32_t *dep_mask, uint32_t *max_deg);
int gkr_line_restrict(gkr_ctx *ctx, const gkr_fr *values, uint32_t k, const gkr_fr *b, const gkr_fr *c,
                      gkr_fr *coef_ascending /* k+1 */);

/* ---- front end (host only): pre-generated circom artefacts -> sub-circuits at the dense boundary ------------------
 * Replaces R1csFile::read / WtnsFile::read (rust/src/aggregator.rs:399-404), convert_constraints_to_nodes
 * (rust/src/convert.rs:360-632), compile (:154-358) and the input-layer evaluation (:793-811): the bytes of a .r1cs
 * and a .wtns file (iden3 formats, BN254) become one (layers, input values) pair per sub-circuit, to be passed to
 * gkr_circuit_create / gkr_witness_eval.  Constraints with an empty linear combination are an error (the reference does
 * not terminate on them, merge_nodes :108-139).  The returned arrays live until gkr_frontend_destroy. */
typedef struct gkr_frontend gkr_frontend;
int gkr_frontend_compile(const uint8_t *r1cs, size_t r1cs_len, const uint8_t *wtns, size_t wtns_len, gkr_frontend **out);
uint32_t gkr_frontend_n_circuits(const gkr_frontend *fe);
uint32_t gkr_frontend_n_public(const gkr_frontend *fe);       /* n_pub_in + n_pub_out of the r1cs header (make_output, :653-667) */
int gkr_frontend_circuit(const gkr_frontend *fe, uint32_t i, uint32_t *n_layers, const gkr_layer_desc **layers,
                         uint32_t *input_k, const gkr_fr **input_values /* 2^input_k canonical values */);
/* the same with the .sym text (`#s,#w,#c,main.name` per line): also builds the reference's `Output` (convert.rs:634-667,
 * parse_sym :851-871): output i < n_outputs is wire i + 1, its witness value and the name after `main.`.  The name pointer
 * lives until gkr_frontend_destroy.  A malformed .sym line is GKR_ERR_INVALID (the reference panics). */
int gkr_frontend_compile_sym(const uint8_t *r1cs, size_t r1cs_len, const uint8_t *wtns, size_t wtns_len, const char *sym,
                             size_t sym_len, gkr_frontend **out);
uint32_t gkr_frontend_n_outputs(const gkr_frontend *fe);
int gkr_frontend_output(const gkr_frontend *fe, uint32_t i, uint32_t *wire, gkr_fr *value, const char **name);
void gkr_frontend_destroy(gkr_frontend *fe);

/* ---- instrumentation ------------------------------------------------------------------------------ */
typedef struct {
    uint64_t kernel_launches;  /* kernels launched by this context since creation / last reset */
    uint64_t h2d_bytes, d2h_bytes;
    double transcript_seconds; /* host time spent in the challenge callback / MiMC7 */
    double wait_seconds;       /* host time spent waiting for round results */
} gkr_stats;
int gkr_ctx_stats(gkr_ctx *ctx, gkr_stats *out, int reset);
/* per-kernel-class device timing (CUDA events around every launch; slows the prover down).
 * classes: 0 gkr_round (no fold), 1 gkr_round fused fold, 2 prod3 round, 3 prod3 fused, 4 wiring,
 * 5 eq, 6 mobius/alt-sum, 7 line, 8 other, 9 gkr_round launches below 2^16 pairs (latency-bound tail),
 * 10 prod3 launches below 2^16 pairs; classes 0-3 count only launches of at least 2^16 pairs */
#define GKR_N_KERNEL_CLASSES 11
typedef struct {
    uint64_t launches[GKR_N_KERNEL_CLASSES];
    double ms[GKR_N_KERNEL_CLASSES];
    double algo_bytes[GKR_N_KERNEL_CLASSES];   /* algorithmic bytes moved (reads + writes of table entries) */
} gkr_profile;
int gkr_ctx_profile(gkr_ctx *ctx, int enable, gkr_profile *out);
/* integer-pipe ceiling of this device: Montgomery products per second in a register-resident loop
 * (ilp = 1, 2 or 4 independent chains per thread; add 16 to time fr_mul_const -- the constant-multiplier fold
 * product -- or 32 to time wide_mac, the unreduced 512-bit multiply-accumulate; blocks_per_sm CTAs of 256 threads) */
int gkr_bench_field_mul(gkr_ctx *ctx, int ilp, int blocks_per_sm, int iters, double *mul_per_second);

/* device self-test of the arithmetic identities the round kernels rely on: failures2[0] = threads whose lazy 512-bit
 * accumulation diverged from the sum of Montgomery products within `iters` products (pseudo-random and maximal
 * operands, checked after every product), failures2[1] = threads whose FP64-pipe fold differed from the integer-pipe
 * fold for the challenge r (NULL: a fixed one).  Both must be 0. */
int gkr_selftest(gkr_ctx *ctx, uint32_t iters, const gkr_fr *r, uint32_t *failures2);

/* diagnostic: the 11 x 11 constants the FP64-pipe fold uses for the challenge r (row-major: out[11 i + j] = balanced
 * base-2^24 digit j of the centred representative of r * 2^(24 i) mod p).  Host only; needs no device. */
int gkr_fold_f64_constants(const gkr_fr *r, double *out121);

#ifdef __cplusplus
}
#endif
#endif
