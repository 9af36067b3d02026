/* gkrb200 -- C ABI of the B200-native GKR prover for batched MiMC hashing.
 *
 * Drop-in boundary for the prover hot path of Consensys/gkr-mimc.  Each entry point names the reference
 * (Go) interface it replaces (file:line relative to the reference root).  The reference-side cgo binding
 * is shown in INTEGRATION.md.
 *
 * Data layout at the boundary == Go's []fr.Element: contiguous elements of 4 x uint64 little-endian limbs,
 * Montgomery form (value * 2^256 mod q), canonical (< q).  BN254 scalar field.
 * No pointer passed in is retained after a call returns (cgo rule).  Outputs go to caller-allocated memory.
 *
 * Every function returns 0 on success or a negative gkrb200_status; gkrb200_last_error() gives a
 * thread-local message.  The reference panics on the same conditions (sumcheck/prover.go:54,:114,
 * gkr/prover.go:84, poly/pool.go:71); the Go shim panics on non-zero.
 * There is NO CPU fallback: without a CUDA device gkrb200_init fails with GKRB200_ERR_CUDA.
 */
#ifndef GKRB200_H
#define GKRB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gkrb200_ctx gkrb200_ctx;

typedef enum {
    GKRB200_OK = 0,
    GKRB200_ERR_ARG = -1,    /* bad size / null pointer / non power of two / inconsistent claims */
    GKRB200_ERR_CUDA = -2,   /* CUDA runtime error or no device */
    GKRB200_ERR_OOM = -3,    /* batch larger than the context was sized for (poly/pool.go:70-72 analogue) */
    GKRB200_ERR_COMM = -4,   /* NCCL / multi-GPU error */
    GKRB200_ERR_STATE = -5,  /* call order (e.g. prove before assign) */
    GKRB200_ERR_VERIFY = -6  /* gkrb200_gkr_verify_mimc: the proof was rejected (message says where) */
} gkrb200_status;

/* circuit.Gate implementations that can cross the ABI (circuit/gates.go:9-21):
 * gates.IdentityGate (circuit/gates/copy.go:9-32) and *gates.CipherGate{Ark} (circuit/gates/cipher.go:11-70). */
#define GKRB200_GATE_IDENTITY 0
#define GKRB200_GATE_CIPHER 1

#define GKRB200_MIMC_LAYERS 94 /* examples/mimc.go:13 */

/* flags for gkrb200_gkr_prove_mimc */
#define GKRB200_PROOF_MONTGOMERY 0u /* words as Go holds them in gkr.Proof (gkr/prover.go:14-18) */
#define GKRB200_PROOF_REGULAR 1u    /* ToBigIntRegular words, as GkrProofToVec writes (prover/gadget/hints.go:236-271) */

const char *gkrb200_last_error(void);
const char *gkrb200_version(void);

/* ---- context ------------------------------------------------------------------------------------------
 * Replaces poly/pool.go:13-126 (MakeLarge/DumpLarge arena of 2^24-entry arrays) and the goroutine pool of
 * sumcheck/worker.go:8-26: one device arena sized for batches up to 2^max_bn, one stream, pinned result slots.
 * `stream` may be NULL (the library creates one) or a cudaStream_t owned by the caller.               */
int gkrb200_init(gkrb200_ctx **ctx, int device, int max_bn, void *stream);
/* Same for a context that will join a communicator of `world` ranks (gkrb200_comm_init): the arena is sized for this rank's
 * 1/world shard from the start (93 x 2^max_bn / world entries) instead of being shrunk at comm_init -- what lets many contexts
 * (proofs in flight) be created side by side on one GPU.                                                                   */
int gkrb200_init_shard(gkrb200_ctx **ctx, int device, int max_bn, void *stream, int world);
void gkrb200_free(gkrb200_ctx *ctx);

/* Multi-GPU: one context per process/GPU.  Rank r of `world` (a power of two <= 8) owns the entries
 * {i : i mod world == r} of every table.  nccl_unique_id: the 128 bytes of an
 * ncclUniqueId produced by gkrb200_comm_unique_id on rank 0 and broadcast by the caller.               */
int gkrb200_comm_unique_id(uint8_t id_out[128]);
int gkrb200_comm_init(gkrb200_ctx *ctx, int rank, int world, const uint8_t nccl_unique_id[128]);
/* Which rank runs the Fiat-Shamir transcript of this communicator's proofs (default 0; set identically on every rank).
 * "the transcript, interpolation and the layer loop ... run once, broadcast".                   */
int gkrb200_comm_set_leader(gkrb200_ctx *ctx, int leader_rank);
/* how the per-round partial sums travel between the ranks: 0 = exchange window (host-shared mapped memory written by the
 * round kernels themselves), 1 = NCCL all-gather, -1 = single GPU (see GKRB200_OPT_EXCHANGE)                       */
int gkrb200_comm_exchange_mode(gkrb200_ctx *ctx);

/* ---- circuit.Circuit.Assign  (circuit/assignment.go:12-32, circuit/circuit.go:48-64,
 *      circuit/gates/cipher.go:25-42) for examples.MimcCircuit() (examples/mimc.go:10-37) -------------------
 * key = inputs[0] (layer 0), msg = inputs[1] (layer 1), n = 2^bn entries each (host memory, untouched).
 * Computes all 94 layers on the device and keeps them in the context (the Assignment).  If out93 != NULL
 * it receives layer 93 (n entries): a[93][x] == MimcKeyedPermutation(msg[x], key[x]) (hash/mimc.go:31-39).
 * Multi-GPU: every rank passes the FULL key/msg; each keeps its strided shard; out93 is the shard
 * (n/world entries, entry j = global index j*world + rank).                                              */
int gkrb200_mimc_assign(gkrb200_ctx *ctx, const uint64_t *key, const uint64_t *msg, size_t n, uint64_t *out93);

/* Same, with key/msg already resident on the device in Go layout (no host<->device copies).            */
int gkrb200_mimc_assign_device(gkrb200_ctx *ctx, const void *d_key, const void *d_msg, size_t n);

/* Assignment[layer] -> host (what Go code reads as a[layer]); local shard when multi-GPU.              */
int gkrb200_assign_layer_to_host(gkrb200_ctx *ctx, int layer, uint64_t *dst, size_t n);

/* ---- the production caller's view: GkrProverHint.Call / HashHint.Call
 *      (prover/gadget/hints.go:135-145,197-233) hand over big.Int values in REGULAR form and want regular words back.
 * flags: GKRB200_IO_INPUT_REGULAR  key/msg are regular-form words (reduced mod q): SetBigInt on the device (hints.go:202-205)
 *        GKRB200_IO_OUTPUT_REGULAR out is written in regular form (ToBigIntRegular)
 *        GKRB200_IO_OUTPUT_HASH    out[x] = a[93][x] + 2*key[x] + msg[x], the hash the gadget exposes
 *                                  (hash.MimcUpdateInplace hash/mimc.go:24-28, prover/gadget/gadget_api.go:28): one batched
 *                                  call replaces N sequential HashHint.Call's and feeds the same assignment to the prover. */
#define GKRB200_IO_INPUT_REGULAR 1u
#define GKRB200_IO_OUTPUT_REGULAR 2u
#define GKRB200_IO_OUTPUT_HASH 4u
int gkrb200_mimc_assign_ex(gkrb200_ctx *ctx, const uint64_t *key, const uint64_t *msg, size_t n, uint64_t *out, uint32_t flags);
/* batched fr.Element.SetBigInt (to_montgomery != 0) / ToBigIntRegular on the device; in == out allowed */
int gkrb200_convert(gkrb200_ctx *ctx, const uint64_t *in, size_t n, uint64_t *out, int to_montgomery);

/* ---- verifier side helpers (gkr/verifier.go:15-132), for the hint's self-check (hints.go:225-229) and for tests ----
 * MultiLin.Evaluate (poly/multilin.go:59-66) of a host table / of a layer of the assignment held by ctx (no copy);
 * multi-GPU: collective, the point is over all bn variables.                                                      */
int gkrb200_mle_evaluate(gkrb200_ctx *ctx, const uint64_t *table, size_t n, const uint64_t *point, uint64_t *out);
int gkrb200_assign_layer_evaluate(gkrb200_ctx *ctx, int layer, const uint64_t *point, int bn, uint64_t *out);
/* gkr.Verify(c, proof, inputs, outputs, qPrime) for the MiMC circuit against the assignment held by ctx (inputs =
 * layers 0 and 1, outputs = layer 93, evaluated on the device).  proof_vec as produced by gkrb200_gkr_prove_mimc with the
 * same flags.  Returns 0 when the proof is accepted, GKRB200_ERR_VERIFY otherwise.                               */
int gkrb200_gkr_verify_mimc(gkrb200_ctx *ctx, const uint64_t *proof_vec, int bn, const uint64_t *qprime, uint32_t flags);
/* Same with the CALLER's inputs and outputs (host tables of 2^bn entries: key = inputs[0], msg = inputs[1], outputs), exactly the
 * arguments of gkr.Verify(c, proof, inputs, outputs, qPrime) (gkr/verifier.go:15): their multilinear extensions are evaluated on
 * the device from these bytes, so the check does not depend on the assignment the prover computed (the hint's self-check,
 * prover/gadget/hints.go:225-229, passes the solver's outputs).  Not available on a sharded context (tables must fit one GPU).    */
int gkrb200_gkr_verify_mimc_io(gkrb200_ctx *ctx, const uint64_t *proof_vec, int bn, const uint64_t *qprime, uint32_t flags,
                               const uint64_t *key, const uint64_t *msg, const uint64_t *outputs);

/* ---- gkr.Prove(c, a, qPrime)  (gkr/prover.go:21-91) for the MiMC circuit --------------------------------
 * Uses the assignment held by ctx.  proof_vec_out receives 1006*bn+183 elements in the order of
 * GkrProofToVec (prover/gadget/hints.go:236-271): all SumcheckProofs[l][k][j], all Claims[l][j], all
 * QPrimes[l][j][k].  Unlike the reference the assignment is NOT consumed (it can be proven again).
 * Multi-GPU: collective call; every rank receives the identical vector.                                 */
int gkrb200_gkr_prove_mimc(gkrb200_ctx *ctx, const uint64_t *qprime, int bn, uint64_t *proof_vec_out, uint32_t flags);
size_t gkrb200_proof_vec_len(int bn); /* prover/gadget/hints.go:76-116 NbOutputs */

/* ---- sumcheck.Prove(X, qPrimes, claims, gate)  (sumcheck/prover.go:46-90) -----------------------------
 * X0/X1: input tables of 2^bn entries (X1 ignored for the identity gate); qprimes: n_q*bn elements;
 * claims: n_claims elements (n_claims == n_q unless n_q == 1).  ark: CipherGate.Ark or NULL.
 * proof_out: bn*(deg+2) coefficients low->high (9 per round cipher, 3 identity); challenges_out: bn;
 * final_claims_out: [Eq(r), X0(r), X1(r)] (1 + arity).  Host tables are not modified.                   */
int gkrb200_sumcheck_prove(gkrb200_ctx *ctx, const uint64_t *X0, const uint64_t *X1, int bn, const uint64_t *qprimes, size_t n_q,
                           const uint64_t *claims, size_t n_claims, int gate_kind, const uint64_t *ark, uint64_t *proof_out,
                           uint64_t *challenges_out, uint64_t *final_claims_out);

/* Same with the tables already resident on the device (Go layout); they are not modified.                */
int gkrb200_sumcheck_prove_device(gkrb200_ctx *ctx, const void *d_X0, const void *d_X1, int bn, const uint64_t *qprimes, size_t n_q,
                                  const uint64_t *claims, size_t n_claims, int gate_kind, const uint64_t *ark, uint64_t *proof_out,
                                  uint64_t *challenges_out, uint64_t *final_claims_out);

/* ---- circuit.Circuit / BuildCircuit / IsInputLayer (circuit/circuit.go:11-91) -------------------------------
 * The library proves examples.MimcCircuit() only.  The shim passes the description of the Circuit it was given: per layer the
 * number of inputs n_in[l] (0 = input layer), the flattened In lists in_flat (sum of n_in entries), the gate kind
 * gate_kinds[l] (-1 for an input layer, else GKRB200_GATE_*) and arks[l] (4 words per layer; read for cipher layers only).
 * Returns 0 when it IS the MiMC circuit (examples/mimc.go:10-37), GKRB200_ERR_ARG (message says what differs) otherwise.     */
int gkrb200_check_mimc_circuit(int n_layers, const int *n_in, const int *in_flat, const int *gate_kinds, const uint64_t *arks);

/* ---- building blocks (each is one device kernel; used by tests, benches and other callers) ----------- */
/* poly.FoldedEqTable / ChunkOfEqTable (poly/eq.go:41-89) and the multi-claim combination of
 * sumcheck/prover.go:121-141: out[x] = sum_j mult_j * eq(q_j, x); multipliers NULL => single table, seed 1 */
int gkrb200_eq_table(gkrb200_ctx *ctx, const uint64_t *qprimes, size_t n_q, int bn, const uint64_t *multipliers, uint64_t *out);
/* MultiLin.Fold (poly/multilin.go:19-36): out[i] = t[i] + r*(t[i+n/2]-t[i]), n/2 outputs                */
int gkrb200_fold(gkrb200_ctx *ctx, const uint64_t *table, size_t n, const uint64_t *r, uint64_t *out);
/* one round of getPartialPolyChunk over the whole table (sumcheck/algo.go:54-205): evals at t=0..deg+1  */
int gkrb200_round_eval(gkrb200_ctx *ctx, const uint64_t *eq, const uint64_t *X0, const uint64_t *X1, size_t n, int gate_kind,
                       const uint64_t *ark, uint64_t *evals_out);
/* element-wise device field ops for arithmetic parity tests: op 0 = mul, 1 = add, 2 = sub, 3 = x^7      */
int gkrb200_fr_batch(gkrb200_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);

/* ---- host-side transcript pieces (serial; run on the CPU by design) -------------------- */
/* common.GetChallenge / hash.MimcHash (common/challenge.go:10, hash/mimc.go:11-18)                       */
int gkrb200_mimc_hash(const uint64_t *in, size_t n, uint64_t *out);
/* poly.InterpolateOnRange (poly/lagrange.go:96-111)                                                      */
int gkrb200_interpolate(const uint64_t *evals, size_t n, uint64_t *coeffs_out);
/* Montgomery <-> regular (fr.Element.SetBigInt / ToBigIntRegular) in place-capable                        */
int gkrb200_to_montgomery(const uint64_t *in, size_t n, uint64_t *out);
int gkrb200_from_montgomery(const uint64_t *in, size_t n, uint64_t *out);

/* test hook for the constant-multiplier fold: the 8 residues K_i = r * 2^(32i+64) * 2^-256 mod q the fold
 * kernels receive for a challenge r (Montgomery image of r in, 8 x 8 32-bit limbs out, low limb first); host-only.            */
int gkrb200_const_mul_table(const uint64_t *r, uint32_t *out64);
/* hash.Arks[round], round 0..90 (hash/ark.go:13-337): the constant of the cipher gate of layer round+3 (examples/mimc.go:29) */
int gkrb200_mimc_ark(int round, uint64_t *out);
/* poly.EvalUnivariate (poly/lagrange.go:31-39): coefficients low -> high, n >= 1                            */
int gkrb200_eval_univariate(const uint64_t *coeffs, size_t n, const uint64_t *x, uint64_t *out);
/* poly.EvalEq (poly/eq.go:19-32): prod_i (1 - q_i - h_i + 2 q_i h_i)                                      */
int gkrb200_eval_eq(const uint64_t *q, const uint64_t *h, size_t n, uint64_t *out);
/* scalar fr.Element methods for host code above the ABI (gate.Eval in the verifier and in tests):
 * op 0 = Mul, 1 = Add, 2 = Sub, 3 = x^7 (b ignored), 4 = Inverse (b ignored, 0 -> 0)                     */
int gkrb200_fr_scalar(int op, const uint64_t *a, const uint64_t *b, uint64_t *out);
/* sumcheck.Verify(claims, proof) (sumcheck/verifier.go:28-65): proof = bn rounds of n_coeffs_per_round coefficients.
 * Writes the bn challenges, the final claim P_last(r_last) and (if not NULL) the recombination challenge
 * GetChallenge(claims).  Returns 0, or GKRB200_ERR_VERIFY when a round fails P(0)+P(1) == expected.       */
int gkrb200_sumcheck_verify(const uint64_t *claims, size_t n_claims, const uint64_t *proof, int bn, int n_coeffs_per_round,
                            uint64_t *challenges_out, uint64_t *final_claim_out, uint64_t *recomb_out);

/* ---- instrumentation ---------------------------------------------------------------------------------- */
typedef struct {
    uint64_t launches_total;      /* kernels launched since the last reset                               */
    uint64_t launches[8];         /* per class: 0 assign, 1 eq, 2 round(eval/fold), 3 fold, 4 multi-eq, 5 staging, 6 misc */
    double kernel_ms[8];          /* per class device time (only when profiling is on)                   */
    double transcript_ms;         /* host time spent in interpolation + MiMC challenges                  */
    double wait_ms;               /* host time spent waiting for device results                          */
    double comm_ms;               /* host time spent in the multi-GPU exchange                           */
    uint64_t rounds;              /* sumcheck rounds run                                                 */
    uint64_t fr_mul_assign;       /* algorithmic field multiplications issued to the device, per class   */
    uint64_t fr_mul_round;
    uint64_t bytes_round;         /* algorithmic bytes moved by the round kernels                        */
    uint64_t h2d_bytes;           /* bytes copied host->device / device->host by the library             */
    uint64_t d2h_bytes;
    double follow_wait_ms;        /* multi-GPU follower ranks: host time asleep waiting for the leader's layer headers */
} gkrb200_stats;
int gkrb200_stats_reset(gkrb200_ctx *ctx);
int gkrb200_stats_get(gkrb200_ctx *ctx, gkrb200_stats *out);
/* when on, every kernel launch is bracketed by CUDA events on the context's stream (adds ~us per launch) */
int gkrb200_set_profiling(gkrb200_ctx *ctx, int on);

/* Tunables / test hooks.  GKRB200_OPT_GENERIC_CIPHER (value 0/1): run single-claim cipher sumchecks through the
 * generic evaluate-at-9-points kernel (the direct restatement of sumcheck/algo.go:54-205) instead of the factored
 * coefficient-sum kernel; both must give identical bytes (tests/test_gpu_parity.py).
 * GKRB200_OPT_PAR8_MAX_PAIRS: rounds with at most this many pairs spread one pair over 8 lanes.            */
#define GKRB200_OPT_GENERIC_CIPHER 1
#define GKRB200_OPT_PAR8_MAX_PAIRS 2
/* GKRB200_OPT_HOST_TAIL_LEN (power of two, 1..32, default 32): once the tables of a sumcheck have at most this many
 * entries (over all ranks) the remaining rounds run on the host -- a device round trip costs more than they do.   */
#define GKRB200_OPT_HOST_TAIL_LEN 3
/* GKRB200_OPT_CF_BLOCKS_PER_SM (>= 1; 0 = as many as fit): cap on the resident blocks per SM the one-wave grid of the
 * factored round kernel is sized for (occupancy experiments).                               */
#define GKRB200_OPT_CF_BLOCKS_PER_SM 4
/* GKRB200_OPT_EXCHANGE (multi-GPU, set identically on every rank): 0 (default) = exchange window -- every rank's round
 * kernel publishes its partial sums straight into a host-shared mapped segment that all ranks' host threads read (no
 * collective, no extra launch); 1 = NCCL all-gather of the partials + a publishing kernel (A/B reference; also what
 * gkrb200_comm_init falls back to, on all ranks together, when /dev/shm cannot be mapped).                          */
#define GKRB200_OPT_EXCHANGE 5
/* GKRB200_OPT_INLINE_MIN_PAIRS (default 0): one-thread-per-pair rounds of the factored cipher kernel with at least this many
 * pairs run the build with the multiplier inlined at every call site, smaller ones the out-of-line build (A/B of the
 * call-ABI cost against instruction-cache footprint).                                            */
#define GKRB200_OPT_INLINE_MIN_PAIRS 6
/* GKRB200_OPT_TRANSCRIPT (multi-GPU, set identically on every rank): 0 (default) = ONE transcript per proof: the communicator's
 * leader rank (gkrb200_comm_set_leader) hashes, the challenges reach every GPU through the exchange window and are consumed
 * by device-side waits, follower ranks only enqueue kernels and sleep; 1 = every rank runs the transcript itself, in lockstep
 * (round-1 behaviour, A/B reference; implied when the exchange is NCCL).                                                   */
#define GKRB200_OPT_TRANSCRIPT 7
/* GKRB200_OPT_CONST_FOLD (default 1): fold with the constant-multiplier product (80 instead of 136 wide multiply-adds; the host
 * derives a small table from the challenge at launch) wherever the challenge is known to the host; 0 = plain Montgomery
 * products (A/B reference: both give identical bytes).                                                                     */
#define GKRB200_OPT_CONST_FOLD 8
int gkrb200_set_option(gkrb200_ctx *ctx, int option, long value);

/* integer-pipe microbenchmarks for the roofline denominator: returns achieved rate.
 * kind 0: IMAD.WIDE.U32 issue rate (result in 1e9 wide-MACs/s); kind 1: dependent fr_mul chains (1e9 Fr-mul/s) */
int gkrb200_microbench(gkrb200_ctx *ctx, int kind, int iters, double *rate_out, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif
