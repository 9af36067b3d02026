/* gkrb200_ec -- C ABI of the B200-native Groth16-side operations (libgkrb200ec.so), SURVEY.md section 8(f4).
 *
 * The Groth16 side of Consensys/gkr-mimc's prover spends its time in bn254.G1Affine.MultiExp and in the FFTs of computeH
 * (gnark-crypto, reference go.mod:7):
 *   prover/gadget/hints.go:182-183   InitialRandomnessHint: the 3N GKR inputs/outputs against pubKGkr / privKGkrSigma
 *   prover/gadget/prove.go:76,91     KrsNotGkr / KrsPrivNotGkr
 *   prover/gadget/prove.go:189,202,221   Bs1, Ar, Krs2 (the G1 multi-exponentiations of ComputeGroth16Proof)
 *   prover/gadget/prove.go:277       Bs: the G2 multi-exponentiation
 *   prover/gadget/prove.go:310-366   computeH: seven FFTs over the constraint domain
 * This library is those operations on the device, plus the rest of InitialRandomnessHint.Call (hints.go:147-192).  Each entry
 * point names the Go interface it replaces; the cgo binding is in INTEGRATION.md section 7.
 *
 * Data at the boundary is Go memory, unchanged:
 *   []bn254.G1Affine  = 8 x uint64 per point: X then Y, fp.Element = 4 little-endian limbs, Montgomery form (v * 2^256 mod p),
 *                       canonical; the point at infinity is all zero
 *   []fr.Element      = 4 x uint64 per scalar; MultiExp takes them in REGULAR form (hints.go:171 calls FromMont first),
 *                       GKRB200EC_SCALARS_MONTGOMERY lets the device do that conversion instead
 * No pointer passed in is retained after a call returns, except by gkrb200ec_g1_set_bases, which COPIES the points to the device.
 * Returns 0 or a negative gkrb200ec_status; gkrb200ec_last_error() gives a thread-local message (the Go shim panics / returns it).
 * There is NO CPU fallback: without a CUDA device gkrb200ec_init fails with GKRB200EC_ERR_CUDA.
 */
#ifndef GKRB200_EC_H
#define GKRB200_EC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gkrb200ec_ctx gkrb200ec_ctx;

typedef enum {
    GKRB200EC_OK = 0,
    GKRB200EC_ERR_ARG = -1,   /* null pointer, unknown base slot, more scalars than bases, scalar >= q in regular form */
    GKRB200EC_ERR_CUDA = -2,  /* CUDA runtime error, no device, or a result that fails the library's self-check (not on the curve) */
    GKRB200EC_ERR_OOM = -3    /* device memory */
} gkrb200ec_status;

#define GKRB200EC_SCALARS_REGULAR 0    /* what G1Affine.MultiExp takes in the pinned gnark-crypto (non-Montgomery fr.Element) */
#define GKRB200EC_SCALARS_MONTGOMERY 1 /* fr.Element as Go holds it; FromMont (hints.go:171) happens on the device */
#define GKRB200EC_MAX_SLOTS 16
#define GKRB200EC_MAX_POINTS (1u << 26)

const char *gkrb200ec_version(void);
const char *gkrb200ec_last_error(void);

/* One context = one device, one stream (NULL: the library creates one; else a cudaStream_t owned by the caller), one grow-only
 * device workspace, up to GKRB200EC_MAX_SLOTS resident base arrays.                                                            */
int gkrb200ec_init(gkrb200ec_ctx **ctx, int device, void *stream);
void gkrb200ec_free(gkrb200ec_ctx *ctx);

/* Proving-key points stay on the device (prover/gadget/setup.go:32 `privKNotGkr, pubKGkr, privKGkrSigma []bn254.G1Affine`):
 * upload once, multiply many times.  Replaces the contents of `slot`.  n == 0 empties it.                                      */
int gkrb200ec_g1_set_bases(gkrb200ec_ctx *ctx, int slot, const uint64_t *points, size_t n);

/* G1Affine.MultiExp(points = bases of `slot`[0..n), scalars, ecc.MultiExpConfig{}):  out = sum_i scalars[i] * points[i],
 * affine, Montgomery (the G1Affine memory image, 8 words).  n may be smaller than the slot.                                   */
int gkrb200ec_g1_multiexp(gkrb200ec_ctx *ctx, int slot, const uint64_t *scalars, size_t n, int scalar_form, uint64_t *out);
/* the same with the scalars already in device memory (e.g. written there by the GKR prover's I/O kernels)                      */
int gkrb200ec_g1_multiexp_device(gkrb200ec_ctx *ctx, int slot, const void *d_scalars, size_t n, int scalar_form, uint64_t *out);
/* one-shot form with host points: exactly the arguments of G1Affine.MultiExp                                                  */
int gkrb200ec_g1_multiexp_points(gkrb200ec_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int scalar_form,
                                 uint64_t *out);

/* InitialRandomnessHint.Call (prover/gadget/hints.go:162-192):
 *     KrsGkr = MultiExp(pubKGkr, scalarsPub); KrsGkrPriv = MultiExp(privKGkrSigma, scalarsPriv); KrsGkr += KrsGkrPriv;
 *     initialRandomness = DeriveRandomnessFromPoint(KrsGkr)
 * krs_gkr_priv_out: 8 words (G1Affine, kept in gadget.Proof, hints.go:186); initial_randomness_out: 4 words in REGULAR form
 * (what ToBigIntRegular(oups[0]) hands the solver, hints.go:189).                                                              */
int gkrb200ec_initial_randomness(gkrb200ec_ctx *ctx, int slot_pub, const uint64_t *scalars_pub, size_t n_pub, int slot_priv,
                                 const uint64_t *scalars_priv, size_t n_priv, int scalar_form, uint64_t *krs_gkr_priv_out,
                                 uint64_t *initial_randomness_out);

/* G1Affine.Add on the device (hints.go:184); a, b, out: 8 words each                                                          */
int gkrb200ec_g1_add(gkrb200ec_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out);

/* ---- G2: Bs.MultiExp(pk.G2.B, wireValuesB, ...) (prover/gadget/prove.go:277) -------------------------------------------------
 * []bn254.G2Affine = 16 x uint64 per point: X.A0, X.A1, Y.A0, Y.A1 (fptower.E2 = {A0, A1 fp.Element}, value A0 + A1 u, u^2 = -1),
 * Montgomery form, infinity all zero.  Same kernels as G1 over the quadratic extension; slots are shared with G1 (a slot holds
 * one kind of point, and using it with the other kind's entry points is GKRB200EC_ERR_ARG).                                     */
int gkrb200ec_g2_set_bases(gkrb200ec_ctx *ctx, int slot, const uint64_t *points, size_t n);
int gkrb200ec_g2_multiexp(gkrb200ec_ctx *ctx, int slot, const uint64_t *scalars, size_t n, int scalar_form, uint64_t *out);
int gkrb200ec_g2_multiexp_device(gkrb200ec_ctx *ctx, int slot, const void *d_scalars, size_t n, int scalar_form, uint64_t *out);
int gkrb200ec_g2_multiexp_points(gkrb200ec_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int scalar_form,
                                 uint64_t *out);
int gkrb200ec_g2_add(gkrb200ec_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out);

/* Host-side pieces of DeriveRandomnessFromPoint (hints.go:147-159); no device involved.
 * gkrb200ec_g1_raw_bytes: G1Affine.RawBytes (X || Y big-endian, regular form; 0x40 then zeros for infinity)
 * gkrb200ec_keccak256:    sha3.NewLegacyKeccak256 (golang.org/x/crypto)
 * gkrb200ec_derive_randomness_from_point: fr.SetBytes(keccak(RawBytes(g1))) in regular form (4 words)                          */
int gkrb200ec_g1_raw_bytes(const uint64_t *g1, uint8_t out[64]);
int gkrb200ec_keccak256(const uint8_t *data, size_t len, uint8_t out[32]);
int gkrb200ec_derive_randomness_from_point(const uint64_t *g1, uint64_t *randomness_out);

/* ---- the FFT half of ComputeGroth16Proof (prover/gadget/prove.go:310-366 computeH) ------------------------------------------
 * gnark-crypto ecc/bn254/fr/fft (reference go.mod:7): Domain, FFT, FFTInverse; DIT = 0, DIF = 1 as fft.Decimation.
 *
 * gkrb200ec_fft_domain_init:  fft.NewDomain(m, 1, true) as Groth16's setup builds it
 *     (pkg/gnark/notinternal/backend/bn254/groth16/setup.go:98): cardinality = next power of two >= m (at most 2^26), Generator =
 *     g^(2^(28 - log n)), FinerGenerator^2 = Generator; twiddle tables are built on the device and stay there.  Replaces the
 *     context's previous domain.
 * gkrb200ec_fft / gkrb200ec_fft_inverse:  domain.FFT(a, decimation, coset) / domain.FFTInverse(a, decimation, coset), in place on
 *     a host slice of exactly `cardinality` Montgomery elements; coset 0 or 1.  DIF: natural input, bit-reversed output; DIT:
 *     bit-reversed input, natural output -- no bit-reversal pass ever runs, as in the reference.
 * gkrb200ec_compute_h:  h = computeH(a, b, c, &pk.Domain): a, b, c hold n_in <= cardinality Montgomery elements (the solved
 *     L.w, R.w, O.w per constraint; zero padding is added on the device); h has `cardinality` elements in REGULAR form, coefficients
 *     in bit-reversed order -- exactly what the reference passes to krs2.MultiExp(pk.G1.Z, h, ...) (prove.go:221; pk.G1.Z is
 *     stored bit-reversed, setup.go:229).  h_out (host, may be NULL) receives a copy; d_h_out (may be NULL) receives the DEVICE
 *     pointer of h, valid until the next FFT / computeH call on this context, ready for gkrb200ec_g1_multiexp_device(...,
 *     GKRB200EC_SCALARS_REGULAR, ...): h never has to leave the GPU.                                                               */
#define GKRB200EC_DIT 0
#define GKRB200EC_DIF 1
int gkrb200ec_fft_domain_init(gkrb200ec_ctx *ctx, uint64_t m);
uint64_t gkrb200ec_fft_domain_cardinality(gkrb200ec_ctx *ctx);
int gkrb200ec_fft(gkrb200ec_ctx *ctx, uint64_t *a, size_t n, int decimation, int coset);
int gkrb200ec_fft_inverse(gkrb200ec_ctx *ctx, uint64_t *a, size_t n, int decimation, int coset);
int gkrb200ec_compute_h(gkrb200ec_ctx *ctx, const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n_in, uint64_t *h_out,
                        const void **d_h_out);

/* ---- ComputeGroth16Proof (prover/gadget/prove.go:100-306) in one call -------------------------------------------------------
 * The proving key's point arrays are resident base slots (uploaded once with gkrb200ec_g1_set_bases / gkrb200ec_g2_set_bases,
 * infinity-filtered exactly as groth16.ProvingKey holds pk.G1.A, pk.G1.B, pk.G2.B; pk.G1.Z whole and bit-reversed as setup.go:229
 * leaves it); a fft domain for the constraint count must be set (gkrb200ec_fft_domain_init).
 *   h   = computeH(a, b, c)                                                        prove.go:126
 *   ar  = MultiExp(pk.G1.A, wireValuesA) + pk.G1.Alpha + r * pk.G1.Delta            prove.go:199-210  -> proof.Ar
 *   bs1 = MultiExp(pk.G1.B, wireValuesB) + pk.G1.Beta  + s * pk.G1.Delta            prove.go:186-196
 *   krs = -(r s) * pk.G1.Delta + MultiExp(pk.G1.Z, h) + s * ar + r * bs1            prove.go:212-262  -> grothProof.Krs (the fork
 *         removed the pk.G1.K term, :224-231; the caller adds KrsPrivNotGkr, :94)
 *   bs  = MultiExp(pk.G2.B, wireValuesB) + s * pk.G2.Delta + pk.G2.Beta             prove.go:265-292  -> proof.Bs
 * r, s: fr.Element (Montgomery) sampled by the caller -- prove.go:154-161 draws them with SetRandom, which stays in Go; with them
 * given, the outputs are a deterministic function of the inputs (and bit-exact with the reference for the same r, s).
 * wire_values_a / wire_values_b: the filtered wire values (prove.go:136-157), in `scalar_form`; a, b, c: solution.A/B/C, Montgomery. */
typedef struct {
    int slot_g1_a, slot_g1_b, slot_g1_z, slot_g2_b;
    const uint64_t *g1_alpha, *g1_beta, *g1_delta; /* 8 words each */
    const uint64_t *g2_beta, *g2_delta;            /* 16 words each */
} gkrb200ec_groth16_pk;
int gkrb200ec_groth16_prove(gkrb200ec_ctx *ctx, const gkrb200ec_groth16_pk *pk, const uint64_t *a, const uint64_t *b, const uint64_t *c,
                            size_t n_constraints, const uint64_t *wire_values_a, size_t n_a, const uint64_t *wire_values_b, size_t n_b,
                            int scalar_form, const uint64_t *r, const uint64_t *s, uint64_t *ar_out, uint64_t *bs_out, uint64_t *krs_out);

/* Tuning / test hooks: force the window width c (2..16, 0 = cost model) and the accumulation task size (0 = twice the mean
 * bucket load).  The result never depends on them.                                                                             */
int gkrb200ec_set_plan(gkrb200ec_ctx *ctx, int window_bits, int task_size);

typedef struct {
    uint64_t launches_total;   /* kernels launched by this context so far */
    uint64_t msm_calls;
    uint32_t last_n, last_c, last_windows, last_task_size;
    uint64_t last_tasks_max;
    uint64_t workspace_bytes;
    uint64_t h2d_bytes, d2h_bytes;
    double last_device_ms;     /* CUDA-event time of the last multi-exponentiation's kernels */
    uint64_t fft_calls;        /* FFT / FFTInverse / computeH calls */
    double last_fft_device_ms; /* CUDA-event time of the last FFT / FFTInverse / computeH's kernels */
} gkrb200ec_stats;
int gkrb200ec_get_stats(gkrb200ec_ctx *ctx, gkrb200ec_stats *out);

#ifdef __cplusplus
}
#endif
#endif
