/* CPU restatement of the gkr-mimc prover hot path -- ORACLE (test infrastructure).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * link, load or execute this file.  The product (gkr-mimc_b200/) never does.
 *
 * The reference is 100 % Go and no Go toolchain exists in this image, so the reference cannot
 * be compiled here (no oracle/_ref).  This file restates its algorithms in plain C, one function
 * per reference function, each citing the file:line it follows (paths relative to the reference
 * root).  It keeps the reference's CPU work decomposition (TryDispatch task splitting, the
 * thresholds of sumcheck/prover.go:14-19, 128-entry sub-chunks of sumcheck/algo.go:9, deep copies
 * of re-used layers) so that it can double as the timed CPU baseline ("port").
 *
 * Pinning (tests/test_oracle.py): hash/hash_test.go:21-27 MimcHash([12]) KAT,
 * poly/multilin_test.go:12-31 fold golden, poly/lagrange_test.go:10-29 basis property,
 * snark/polynomial/univariate_test.go:39-45, and word-for-word agreement with the independent
 * Python big-int restatement oracle/pyref.py + tests/golden fixtures.  Proof bytes themselves have
 * no golden in the reference (it has none); they are pinned by exact arithmetic + verifier
 * acceptance (gkr/verifier.go restated below).
 *
 * All fr_t values crossing this API are Montgomery-form 4x64 LE limbs (the Go memory layout).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <sched.h>
#include <stdatomic.h>
#include "fr.h"

#define MIMC_ROUNDS 91 /* hash/mimc.go:8 */
#define N_LAYERS 94    /* examples/mimc.go:13 */

static const fr_t ARKS[MIMC_ROUNDS] = {
#include "arks.inc"
};

/* ------------------------------------------------------------------ worker pool (sumcheck/worker.go:8-26)
 * g_ncpu stands in for runtime.NumCPU(): g_ncpu-1 persistent pthreads + the calling thread pull task
 * indices from a shared counter, like the reference's goroutines pull closures from jobQueue. */
#define MAX_THREADS 512
typedef void (*range_fn)(size_t start, size_t stop, void *ctx, int task);
static int g_ncpu = 1;
static struct {
    pthread_t th[MAX_THREADS];
    int n_workers;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    unsigned long gen;
    int shutdown;
    range_fn f;
    void *ctx;
    size_t per, extra;
    atomic_uint_fast64_t desc; /* (run id << 32) | n_tasks of the current run: ONE word, so a worker validates its ticket against a consistent pair */
    atomic_uint_fast64_t next; /* (run id << 32) | next task index: the ticket counter */
    atomic_size_t done;
} g_pool = {.mu = PTHREAD_MUTEX_INITIALIZER, .cv = PTHREAD_COND_INITIALIZER};

/* A ticket is valid iff it carries the run id of the descriptor word AND its index is below that run's n_tasks.  A worker that
 * drew a ticket in an earlier run (an overshoot past that run's n_tasks) and was descheduled before looking at it can therefore
 * never mistake it for a task of a later run.  (An earlier version reset a plain counter to 0 per run and compared the stale
 * index with the NEW run's n_tasks: the task ran twice, `done` reached n_tasks one task early and the caller went on while a
 * task was still writing -- a ~0.3 % corruption of the 91-claim eq table under CPU contention.)  A valid ticket is a real,
 * not yet executed task of the current run, so the run cannot end -- and the descriptor fields cannot change -- under it. */
static void pool_drain(void) {
    for (;;) {
        const uint_fast64_t t = atomic_fetch_add(&g_pool.next, 1);
        const uint_fast64_t d = atomic_load(&g_pool.desc);
        if ((t >> 32) != (d >> 32) || (uint32_t)t >= (uint32_t)d) return;
        const size_t i = (uint32_t)t;
        size_t off = i < g_pool.extra ? i : g_pool.extra;
        size_t start = i * g_pool.per + off;
        size_t stop = start + g_pool.per + (i < g_pool.extra ? 1 : 0);
        g_pool.f(start, stop, g_pool.ctx, (int)i);
        atomic_fetch_add(&g_pool.done, 1);
    }
}
static void *pool_worker(void *arg) {
    (void)arg;
    unsigned long seen = 0;
    for (;;) {
        pthread_mutex_lock(&g_pool.mu);
        while (g_pool.gen == seen && !g_pool.shutdown) pthread_cond_wait(&g_pool.cv, &g_pool.mu);
        seen = g_pool.gen;
        int stop = g_pool.shutdown;
        pthread_mutex_unlock(&g_pool.mu);
        if (stop) return NULL;
        pool_drain();
    }
}
static void pool_stop(void) {
    if (!g_pool.n_workers) return;
    pthread_mutex_lock(&g_pool.mu);
    g_pool.shutdown = 1;
    pthread_cond_broadcast(&g_pool.cv);
    pthread_mutex_unlock(&g_pool.mu);
    for (int i = 0; i < g_pool.n_workers; i++) pthread_join(g_pool.th[i], NULL);
    g_pool.n_workers = 0;
    g_pool.shutdown = 0;
}
/* run tasks [0,n_tasks): task i covers per (+1 for the first `extra`) iterations */
static void pool_run(size_t n_tasks, size_t per, size_t extra, range_fn f, void *ctx) {
    /* Nobody reads f/ctx/per/extra now: every ticket of the previous run has been executed (that run waited for `done`), and
     * tickets of this run only exist after the store to `next` below, which publishes the fields and the descriptor word. */
    const uint_fast64_t run = (atomic_load(&g_pool.desc) >> 32) + 1;
    g_pool.f = f;
    g_pool.ctx = ctx;
    g_pool.per = per;
    g_pool.extra = extra;
    atomic_store(&g_pool.done, 0);
    atomic_store(&g_pool.desc, (run << 32) | (uint_fast64_t)(uint32_t)n_tasks);
    atomic_store(&g_pool.next, run << 32);
    if (g_pool.n_workers) {
        pthread_mutex_lock(&g_pool.mu);
        g_pool.gen++;
        pthread_cond_broadcast(&g_pool.cv);
        pthread_mutex_unlock(&g_pool.mu);
    }
    pool_drain();
    while (atomic_load(&g_pool.done) < n_tasks) sched_yield();
}
void orc_set_threads(int n) {
    if (n < 1) n = 1;
    if (n > MAX_THREADS) n = MAX_THREADS;
    if (n == g_ncpu && g_pool.n_workers == n - 1) return;
    pool_stop();
    g_ncpu = n;
    for (int i = 0; i < n - 1; i++) {
        if (pthread_create(&g_pool.th[i], NULL, pool_worker, NULL)) break;
        g_pool.n_workers++;
    }
}
int orc_get_threads(void) { return g_ncpu; }

/* ------------------------------------------------------------------ field exports (tests) */
void orc_fr_mul(const fr_t *a, const fr_t *b, fr_t *z) { fr_mul(z, a, b); }
void orc_fr_mul_portable(const fr_t *a, const fr_t *b, fr_t *z) { fr_mul_portable(z, a, b); }
const char *orc_fr_mul_kind(void) { return ORC_FR_MUL_KIND; }
/* single-thread cost of one fr.Element.Mul in ns: [0] throughput form (4 independent chains), [1] latency form (one dependent chain) */
void orc_bench_fr_mul(size_t iters, double out_ns[2]) {
    struct timespec t0, t1;
    fr_t a[4], b;
    for (int k = 0; k < 4; k++) fr_set_u64(&a[k], 0x1234567 + (uint64_t)k);
    fr_set_u64(&b, 0xabcdef01);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (size_t i = 0; i < iters; i++)
        for (int k = 0; k < 4; k++) fr_mul(&a[k], &a[k], &b);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    out_ns[0] = ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / (4.0 * (double)iters);
    fr_add(&b, &a[0], &a[1]);
    fr_add(&b, &b, &a[2]);
    fr_add(&b, &b, &a[3]);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (size_t i = 0; i < 4 * iters; i++) fr_mul(&b, &b, &b);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    out_ns[1] = ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / (4.0 * (double)iters);
    if (fr_is_zero(&b)) out_ns[1] = -out_ns[1]; /* keeps the chain alive */
}
void orc_fr_add(const fr_t *a, const fr_t *b, fr_t *z) { fr_add(z, a, b); }
void orc_fr_sub(const fr_t *a, const fr_t *b, fr_t *z) { fr_sub(z, a, b); }
void orc_fr_inv(const fr_t *a, fr_t *z) { fr_inv(z, a); }
void orc_to_mont(const fr_t *in, fr_t *out, size_t n) { for (size_t i = 0; i < n; i++) fr_to_mont(&out[i], &in[i]); }
void orc_from_mont(const fr_t *in, fr_t *out, size_t n) { for (size_t i = 0; i < n; i++) fr_from_mont(&out[i], &in[i]); }

/* ------------------------------------------------------------------ hash/ */
/* hash/poseidon.go:129-135 SBoxInplace: x^7 = ((x^2 * x)^2) * x */
static inline void sbox(fr_t *x) {
    fr_t t = *x;
    fr_sqr(x, x);
    fr_mul(x, x, &t);
    fr_sqr(x, x);
    fr_mul(x, x, &t);
}
/* hash/mimc.go:31-39 */
static void mimc_keyed_permutation(fr_t *res, const fr_t *x, const fr_t *key) {
    *res = *x;
    for (int i = 0; i < MIMC_ROUNDS; i++) {
        fr_add(res, res, key);
        fr_add(res, res, &ARKS[i]);
        sbox(res);
    }
}
/* hash/mimc.go:24-28 + :43-49 */
static void mimc_update(fr_t *state, const fr_t *block) {
    fr_t ns;
    mimc_keyed_permutation(&ns, block, state);
    fr_add(&ns, &ns, state); /* MimcBlockCipher adds the key */
    fr_add(state, state, &ns);
    fr_add(state, state, block);
}
/* hash/mimc.go:11-18 ; common/challenge.go:10-12 */
void orc_mimc_hash(const fr_t *in, size_t n, fr_t *out) {
    fr_t s;
    fr_set_zero(&s);
    for (size_t i = 0; i < n; i++) mimc_update(&s, &in[i]);
    *out = s;
}
void orc_mimc_keyed_permutation(const fr_t *x, const fr_t *key, fr_t *out) { mimc_keyed_permutation(out, x, key); }

/* common/common.go:49-55 */
void orc_random_fr_array(fr_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) fr_set_u64(&out[i], ((uint64_t)i * (uint64_t)i) ^ 0xf45c9df123fULL);
}

/* ------------------------------------------------------------------ dispatch (common/parallelize.go) */
/* common/parallelize.go:49-88 TryDispatch: returns the number of tasks run (0 => caller runs inline).
 * Tasks are executed by the g_ncpu pool workers pulling from a shared queue (sumcheck/worker.go:14-26). */
static int try_dispatch(size_t n_iter, size_t min_task, range_fn f, void *ctx) {
    size_t n_tasks = (size_t)g_ncpu * 8;
    size_t per = n_iter / n_tasks;
    if (per < min_task) {
        per = min_task;
        n_tasks = n_iter / per;
    }
    if (n_tasks <= 1) return 0;
    pool_run(n_tasks, per, n_iter - n_tasks * per, f, ctx);
    return (int)n_tasks;
}
/* common/parallelize.go:9-44 Parallelize */
static void parallelize(size_t n_iter, range_fn f, void *ctx) {
    size_t n_tasks = (size_t)g_ncpu;
    size_t per = n_iter / n_tasks;
    if (per < 1) {
        per = 1;
        n_tasks = n_iter;
    }
    pool_run(n_tasks, per, n_iter - n_tasks * per, f, ctx);
}

/* ------------------------------------------------------------------ poly/ */
/* poly/multilin.go:26-36 FoldChunk (clobbers the top half like the reference) */
static void fold_chunk(fr_t *tab, size_t len, const fr_t *r, size_t start, size_t stop) {
    size_t mid = len / 2;
    fr_t *bot = tab, *top = tab + mid;
    for (size_t i = start; i < stop; i++) {
        fr_sub(&top[i], &top[i], &bot[i]);
        fr_mul(&top[i], &top[i], r);
        fr_add(&bot[i], &bot[i], &top[i]);
    }
}
/* poly/multilin.go:19-23 Fold; the caller then uses len/2 entries */
void orc_fold(fr_t *tab, size_t len, const fr_t *r) { fold_chunk(tab, len, r, 0, len / 2); }

/* poly/multilin.go:59-66 Evaluate */
void orc_evaluate(const fr_t *tab, size_t len, const fr_t *coords, size_t n_coords, fr_t *out) {
    fr_t *c = (fr_t *)malloc(len * sizeof(fr_t));
    memcpy(c, tab, len * sizeof(fr_t));
    size_t l = len;
    for (size_t k = 0; k < n_coords; k++) {
        orc_fold(c, l, &coords[k]);
        l /= 2;
    }
    *out = c[0];
    free(c);
}

/* poly/eq.go:19-32 EvalEq */
void orc_eval_eq(const fr_t *q, const fr_t *h, size_t n, fr_t *out) {
    fr_t res, nxt, one, sum;
    fr_set_one(&one);
    fr_set_one(&res);
    for (size_t i = 0; i < n; i++) {
        fr_mul(&nxt, &q[i], &h[i]);
        fr_add(&nxt, &nxt, &nxt);
        fr_add(&nxt, &nxt, &one);
        fr_add(&sum, &q[i], &h[i]);
        fr_sub(&nxt, &nxt, &sum);
        fr_mul(&res, &res, &nxt);
    }
    *out = res;
}

/* poly/eq.go:41-59 FoldedEqTable (multiplier == NULL => one) */
void orc_folded_eq_table(fr_t *t, const fr_t *q, size_t n, const fr_t *multiplier) {
    if (multiplier)
        t[0] = *multiplier;
    else
        fr_set_one(&t[0]);
    for (size_t i = 0; i < n; i++) {
        for (size_t j = 0; j < ((size_t)1 << i); j++) {
            size_t J = j << (n - i);
            size_t JN = J + ((size_t)1 << (n - 1 - i));
            fr_mul(&t[JN], &q[i], &t[J]);
            fr_sub(&t[J], &t[J], &t[JN]);
        }
    }
}

static size_t log2_ceil(size_t x) {
    size_t l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}

/* poly/eq.go:62-89 ChunkOfEqTable */
void orc_chunk_of_eq_table(fr_t *eq, size_t chunk_id, size_t chunk_size, const fr_t *q, size_t n, const fr_t *multiplier) {
    size_t n_chunks = ((size_t)1 << n) / chunk_size;
    size_t log_n = log2_ceil(n_chunks);
    fr_t one, tmp, r;
    fr_set_one(&one);
    r = multiplier ? *multiplier : one;
    for (size_t k = 0; k < log_n; k++) {
        const fr_t *rho = &q[log_n - k - 1];
        if ((chunk_id >> k) & 1) {
            fr_mul(&r, &r, rho);
        } else {
            fr_sub(&tmp, &one, rho);
            fr_mul(&r, &r, &tmp);
        }
    }
    orc_folded_eq_table(eq + chunk_id * chunk_size, q + log_n, n - log_n, &r);
}

/* poly/lagrange.go:31-39 EvalUnivariate */
void orc_eval_univariate(const fr_t *coeffs, size_t n, const fr_t *x, fr_t *out) {
    fr_t res = coeffs[n - 1];
    for (size_t i = n - 1; i-- > 0;) {
        fr_mul(&res, &res, x);
        fr_add(&res, &res, &coeffs[i]);
    }
    *out = res;
}

#define MAX_DOMAIN 12 /* poly/lagrange.go:21 */
static fr_t g_lagrange[MAX_DOMAIN + 1][MAX_DOMAIN][MAX_DOMAIN];
static int g_lagrange_ready = 0;

/* poly/lagrange.go:42-92 LagrangeCoefficient */
static void lagrange_coefficient(int domain, fr_t out[MAX_DOMAIN][MAX_DOMAIN]) {
    fr_t bin[MAX_DOMAIN][2];
    for (int i = 0; i < domain; i++) {
        fr_t ic;
        fr_set_u64(&ic, (uint64_t)i);
        fr_neg(&bin[i][0], &ic);
        fr_set_one(&bin[i][1]);
    }
    for (int l = 0; l < domain; l++) {
        fr_t acc[MAX_DOMAIN], upd[MAX_DOMAIN], tmp;
        memset(acc, 0, sizeof acc);
        fr_set_one(&acc[0]);
        for (int i = 0; i < domain; i++) {
            if (i == l) continue;
            memset(upd, 0, sizeof upd);
            for (int j = 0; j < domain; j++) {
                int kmax = (domain - j) < 2 ? (domain - j) : 2;
                for (int k = 0; k < kmax; k++) {
                    fr_mul(&tmp, &acc[j], &bin[i][k]);
                    fr_add(&upd[j + k], &upd[j + k], &tmp);
                }
            }
            memcpy(acc, upd, sizeof acc);
        }
        fr_t lf, norm;
        fr_set_u64(&lf, (uint64_t)l);
        orc_eval_univariate(acc, (size_t)domain, &lf, &norm);
        fr_inv(&norm, &norm);
        for (int i = 0; i < domain; i++) fr_mul(&out[l][i], &acc[i], &norm);
    }
}
/* poly/lagrange.go:25-30 initLagrangePolynomials */
static void init_lagrange(void) {
    if (g_lagrange_ready) return;
    for (int d = 1; d <= MAX_DOMAIN; d++) lagrange_coefficient(d, g_lagrange[d]);
    g_lagrange_ready = 1;
}
void orc_lagrange_coefficient(int domain, fr_t *out /* domain*domain */) {
    init_lagrange();
    for (int l = 0; l < domain; l++)
        for (int j = 0; j < domain; j++) out[l * domain + j] = g_lagrange[domain][l][j];
}
/* poly/lagrange.go:96-111 InterpolateOnRange */
void orc_interpolate_on_range(const fr_t *values, size_t n, fr_t *out) {
    init_lagrange();
    fr_t tmp;
    for (size_t j = 0; j < n; j++) fr_set_zero(&out[j]);
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {
            fr_mul(&tmp, &g_lagrange[n][i][j], &values[i]);
            fr_add(&out[j], &out[j], &tmp);
        }
}

/* ------------------------------------------------------------------ circuit/gates */
enum { GATE_IDENTITY = 0, GATE_CIPHER = 1 };
typedef struct {
    int kind;
    fr_t ark;
} gate_t;
/* gates/cipher.go:68-70, gates/copy.go:30-32 */
static int gate_degree(const gate_t *g) { return g->kind == GATE_CIPHER ? 7 : 1; }
static int gate_arity(const gate_t *g) { return g->kind == GATE_CIPHER ? 2 : 1; }
/* gates/cipher.go:25-42 EvalBatch, gates/copy.go:15-17 */
static void gate_eval_batch(const gate_t *g, fr_t *res, size_t n, const fr_t *const *xs) {
    if (g->kind == GATE_IDENTITY) {
        memcpy(res, xs[0], n * sizeof(fr_t));
        return;
    }
    const fr_t *ls = xs[0], *rs = xs[1];
    fr_t tmp;
    for (size_t i = 0; i < n; i++) {
        fr_add(&tmp, &rs[i], &g->ark);
        fr_add(&tmp, &tmp, &ls[i]);
        fr_sqr(&res[i], &tmp);
        fr_mul(&res[i], &res[i], &tmp);
        fr_sqr(&res[i], &res[i]);
        fr_mul(&res[i], &res[i], &tmp);
    }
}
/* gates/cipher.go:45-55 Eval, gates/copy.go:20-22 */
static void gate_eval(const gate_t *g, fr_t *res, const fr_t *const *xs) {
    if (g->kind == GATE_IDENTITY) {
        *res = *xs[0];
        return;
    }
    fr_t tmp;
    fr_add(&tmp, xs[1], &g->ark);
    fr_add(&tmp, &tmp, xs[0]);
    fr_sqr(res, &tmp);
    fr_mul(res, res, &tmp);
    fr_sqr(res, res);
    fr_mul(res, res, &tmp);
}

/* ------------------------------------------------------------------ sumcheck/ */
#define FOLDING_MIN_TASK (1u << 10)     /* sumcheck/prover.go:15 */
#define PARTIAL_EVAL_MIN_TASK (1u << 6) /* :16 */
#define EQ_TABLE_CHUNK (1u << 8)        /* :17 */
#define ADD_INPLACE_MIN_CHUNK (1u << 10) /* :18 */
#define EVAL_SUB_CHUNK 128              /* sumcheck/algo.go:9 */
#define MAX_EVALS 9
#define MAX_INPUTS 2

typedef struct {
    fr_t *X[MAX_INPUTS];
    int n_inputs;
    fr_t *Eq;
    size_t len; /* current table length */
    gate_t gate;
    int degree; /* gate degree + 1 (sumcheck/prover.go:95) */
} instance_t;

/* sumcheck/algo.go:54-205 getPartialPolyChunk */
static void get_partial_poly_chunk(const instance_t *inst, size_t start, size_t stop, fr_t *evals) {
    int n_evals = inst->degree + 1;
    int n_in = inst->n_inputs;
    size_t mid = inst->len / 2;
    fr_t tmp_evals[EVAL_SUB_CHUNK], tmp_eqs[EVAL_SUB_CHUNK], d_eqs[EVAL_SUB_CHUNK];
    fr_t tmp_xs[EVAL_SUB_CHUNK * MAX_INPUTS], d_xs[EVAL_SUB_CHUNK * MAX_INPUTS];
    const fr_t *buf[MAX_INPUTS];
    fr_t v;
    for (int t = 0; t < n_evals; t++) fr_set_zero(&evals[t]);

    for (size_t s0 = start; s0 < stop; s0 += EVAL_SUB_CHUNK) {
        size_t s1 = s0 + EVAL_SUB_CHUNK < stop ? s0 + EVAL_SUB_CHUNK : stop;
        size_t len = s1 - s0;
        /* t = 0 : read the bottom halves in place (algo.go:107-125) */
        for (int k = 0; k < n_in; k++) buf[k] = inst->X[k] + s0;
        gate_eval_batch(&inst->gate, tmp_evals, len, buf);
        for (size_t x = 0; x < len; x++) {
            fr_mul(&v, &inst->Eq[s0 + x], &tmp_evals[x]);
            fr_add(&evals[0], &evals[0], &v);
        }
        /* t = 1 : top halves (algo.go:127-147) */
        for (int k = 0; k < n_in; k++) buf[k] = inst->X[k] + s0 + mid;
        gate_eval_batch(&inst->gate, tmp_evals, len, buf);
        for (size_t x = 0; x < len; x++) {
            fr_mul(&v, &inst->Eq[s0 + mid + x], &tmp_evals[x]);
            fr_add(&evals[1], &evals[1], &v);
        }
        /* t >= 2 by repeated addition of the differences (algo.go:149-199) */
        memcpy(tmp_eqs, inst->Eq + s0 + mid, len * sizeof(fr_t));
        for (size_t x = 0; x < len; x++) fr_sub(&d_eqs[x], &inst->Eq[s0 + mid + x], &inst->Eq[s0 + x]);
        for (int k = 0; k < n_in; k++) {
            size_t off = (size_t)k * len;
            for (size_t x = 0; x < len; x++) fr_sub(&d_xs[off + x], &inst->X[k][s0 + mid + x], &inst->X[k][s0 + x]);
            memcpy(tmp_xs + off, inst->X[k] + s0 + mid, len * sizeof(fr_t));
            buf[k] = tmp_xs + off;
        }
        for (int t = 2; t < n_evals; t++) {
            for (size_t x = 0; x < len; x++) fr_add(&tmp_eqs[x], &tmp_eqs[x], &d_eqs[x]);
            for (size_t kx = 0; kx < (size_t)n_in * len; kx++) fr_add(&tmp_xs[kx], &tmp_xs[kx], &d_xs[kx]);
            gate_eval_batch(&inst->gate, tmp_evals, len, buf);
            for (size_t x = 0; x < len; x++) {
                fr_mul(&v, &tmp_eqs[x], &tmp_evals[x]);
                fr_add(&evals[t], &evals[t], &v);
            }
        }
    }
}

typedef struct {
    const instance_t *inst;
    fr_t *partials; /* [n_tasks][MAX_EVALS] */
} pe_ctx_t;
static void pe_task(size_t start, size_t stop, void *c, int task) {
    pe_ctx_t *ctx = (pe_ctx_t *)c;
    get_partial_poly_chunk(ctx->inst, start, stop, ctx->partials + (size_t)task * MAX_EVALS);
}
/* sumcheck/prover.go:148-163 dispatchPartialEvals + :236-245 consumeAccumulate */
static void dispatch_partial_evals(const instance_t *inst, fr_t *evals) {
    size_t mid = inst->len / 2;
    size_t max_tasks = (size_t)g_ncpu * 8 + 1; /* TryDispatch never makes more than 8*NumCPU tasks */
    fr_t *partials = (fr_t *)malloc(max_tasks * MAX_EVALS * sizeof(fr_t));
    pe_ctx_t ctx = {inst, partials};
    int n_tasks = try_dispatch(mid, PARTIAL_EVAL_MIN_TASK, pe_task, &ctx);
    int n_evals = inst->degree + 1;
    if (n_tasks < 1) {
        get_partial_poly_chunk(inst, 0, mid, evals);
    } else {
        for (int t = 0; t < n_evals; t++) evals[t] = partials[t];
        for (int i = 1; i < n_tasks; i++)
            for (int t = 0; t < n_evals; t++) fr_add(&evals[t], &evals[t], &partials[(size_t)i * MAX_EVALS + t]);
    }
    free(partials);
}

typedef struct {
    instance_t *inst;
    const fr_t *r;
} fold_ctx_t;
/* sumcheck/algo.go:46-51 foldChunk */
static void fold_task(size_t start, size_t stop, void *c, int task) {
    (void)task;
    fold_ctx_t *ctx = (fold_ctx_t *)c;
    fold_chunk(ctx->inst->Eq, ctx->inst->len, ctx->r, start, stop);
    for (int k = 0; k < ctx->inst->n_inputs; k++) fold_chunk(ctx->inst->X[k], ctx->inst->len, ctx->r, start, stop);
}
/* sumcheck/prover.go:167-190 dispatchFolding */
static void dispatch_folding(instance_t *inst, const fr_t *r) {
    size_t mid = inst->len / 2;
    fold_ctx_t ctx = {inst, r};
    if (try_dispatch(mid, FOLDING_MIN_TASK, fold_task, &ctx) < 1) fold_task(0, mid, &ctx, 0);
    inst->len = mid;
}

typedef struct {
    fr_t *eq;
    const fr_t *q;
    size_t n;
    const fr_t *mult;
} eq_ctx_t;
/* sumcheck/algo.go:209-215 computeEqTableJob */
static void eq_task(size_t start, size_t stop, void *c, int task) {
    (void)task;
    eq_ctx_t *ctx = (eq_ctx_t *)c;
    for (size_t id = start; id < stop; id++) orc_chunk_of_eq_table(ctx->eq, id, EQ_TABLE_CHUNK, ctx->q, ctx->n, ctx->mult);
}
/* sumcheck/prover.go:193-212 dispatchEqTable */
static void dispatch_eq_table(fr_t *eq, size_t len, const fr_t *q, size_t n, const fr_t *mult) {
    size_t nb_chunks = len / EQ_TABLE_CHUNK;
    eq_ctx_t ctx = {eq, q, n, mult};
    if (try_dispatch(nb_chunks, 1, eq_task, &ctx) < 1) orc_folded_eq_table(eq, q, n, mult);
}
typedef struct {
    fr_t *a;
    const fr_t *b;
} add_ctx_t;
/* sumcheck/algo.go:219-223 addInPlace */
static void add_task(size_t start, size_t stop, void *c, int task) {
    (void)task;
    add_ctx_t *ctx = (add_ctx_t *)c;
    for (size_t i = start; i < stop; i++) fr_add(&ctx->a[i], &ctx->a[i], &ctx->b[i]);
}
/* sumcheck/prover.go:215-232 dispatchAdditions */
static void dispatch_additions(fr_t *a, const fr_t *b, size_t len) {
    add_ctx_t ctx = {a, b};
    if (try_dispatch(len, ADD_INPLACE_MIN_CHUNK, add_task, &ctx) < 1) add_task(0, len, &ctx, 0);
}

/* sumcheck/prover.go:102-144 makeEqTable; returns rho (zero if no claims) */
static void make_eq_table(instance_t *inst, const fr_t *claims, size_t n_claims, const fr_t *qprimes, size_t n_q, size_t bn, fr_t *rho_out) {
    if (n_claims != n_q && n_q > 1) {
        fprintf(stderr, "oracle: multi-instance with %zu qPrimes but %zu claims\n", n_q, n_claims);
        abort();
    }
    dispatch_eq_table(inst->Eq, inst->len, qprimes, bn, NULL);
    fr_set_zero(rho_out);
    if (n_claims < 1) return;
    fr_t rho, mult;
    orc_mimc_hash(claims, n_claims, &rho);
    mult = rho;
    if (n_q > 1) {
        fr_t *tmp = (fr_t *)malloc(inst->len * sizeof(fr_t));
        for (size_t i = 1; i < n_q; i++) {
            dispatch_eq_table(tmp, inst->len, qprimes + i * bn, bn, &mult);
            dispatch_additions(inst->Eq, tmp, inst->len);
            fr_mul(&mult, &mult, &rho);
        }
        free(tmp);
    }
    *rho_out = rho;
}

/* test hook: the (multi-claim) eq table of sumcheck/prover.go:102-144 */
void orc_make_eq_table(const fr_t *claims, size_t n_claims, const fr_t *qprimes, size_t n_q, size_t bn, fr_t *eq_out, fr_t *rho_out) {
    instance_t inst;
    memset(&inst, 0, sizeof inst);
    inst.Eq = eq_out;
    inst.len = (size_t)1 << bn;
    make_eq_table(&inst, claims, n_claims, qprimes, n_q, bn, rho_out);
}

/* test hook: one round's evaluations, sumcheck/algo.go:54-205 over the whole table */
void orc_partial_evals(const fr_t *eq, const fr_t *x0, const fr_t *x1, size_t len, int gate_kind, const fr_t *ark, fr_t *evals_out) {
    instance_t inst;
    memset(&inst, 0, sizeof inst);
    inst.Eq = (fr_t *)eq;
    inst.X[0] = (fr_t *)x0;
    inst.X[1] = (fr_t *)x1;
    inst.gate.kind = gate_kind;
    if (ark) inst.gate.ark = *ark;
    inst.n_inputs = gate_arity(&inst.gate);
    inst.len = len;
    inst.degree = gate_degree(&inst.gate) + 1;
    dispatch_partial_evals(&inst, evals_out);
}

/* sumcheck/prover.go:46-90 Prove.
 * X[k] (n_inputs tables of 2^bn) are CONSUMED (folded in place) like the reference.
 * proof_out: bn * (degree+2) coefficients low->high; challenges_out: bn; final_claims_out: 1+n_inputs. */
void orc_sumcheck_prove(fr_t *x0, fr_t *x1, size_t bn, const fr_t *qprimes, size_t n_q, const fr_t *claims, size_t n_claims, int gate_kind,
                        const fr_t *ark, fr_t *proof_out, fr_t *challenges_out, fr_t *final_claims_out) {
    instance_t inst;
    memset(&inst, 0, sizeof inst);
    inst.gate.kind = gate_kind;
    if (ark) inst.gate.ark = *ark;
    inst.n_inputs = gate_arity(&inst.gate);
    inst.X[0] = x0;
    inst.X[1] = x1;
    inst.len = (size_t)1 << bn;
    inst.degree = gate_degree(&inst.gate) + 1;
    inst.Eq = (fr_t *)malloc(inst.len * sizeof(fr_t));
    int n_evals = inst.degree + 1;
    fr_t rho, evals[MAX_EVALS], r;
    make_eq_table(&inst, claims, n_claims, qprimes, n_q, bn, &rho);
    for (size_t k = 0; k < bn; k++) {
        dispatch_partial_evals(&inst, evals);
        fr_t *coeffs = proof_out + k * (size_t)n_evals;
        orc_interpolate_on_range(evals, (size_t)n_evals, coeffs);
        orc_mimc_hash(coeffs, (size_t)n_evals, &r);
        dispatch_folding(&inst, &r);
        challenges_out[k] = r;
    }
    final_claims_out[0] = inst.Eq[0];
    for (int k = 0; k < inst.n_inputs; k++) final_claims_out[1 + k] = inst.X[k][0];
    free(inst.Eq);
}

/* sumcheck/verifier.go:28-65 Verify. returns 0 if accepted, (round+1) of the first failing round otherwise. */
int orc_sumcheck_verify(const fr_t *claims, size_t n_claims, const fr_t *proof, size_t bn, size_t n_coeffs, fr_t *challenges_out,
                        fr_t *final_claim_out, fr_t *recomb_out) {
    fr_t expected, rho, zero, one, a0, a1, r;
    fr_set_zero(&zero);
    fr_set_one(&one);
    orc_mimc_hash(claims, n_claims, &rho);
    orc_eval_univariate(claims, n_claims, &rho, &expected);
    for (size_t i = 0; i < bn; i++) {
        const fr_t *p = proof + i * n_coeffs;
        orc_eval_univariate(p, n_coeffs, &zero, &a0);
        orc_eval_univariate(p, n_coeffs, &one, &a1);
        fr_add(&a0, &a0, &a1);
        if (!fr_eq(&a0, &expected)) return (int)i + 1;
        orc_mimc_hash(p, n_coeffs, &r);
        challenges_out[i] = r;
        orc_eval_univariate(p, n_coeffs, &r, &expected);
    }
    *final_claim_out = expected;
    *recomb_out = rho;
    return 0;
}

/* sumcheck/instance.go:49-68 Evaluation (brute force; tests only) */
void orc_evaluation(int gate_kind, const fr_t *ark, const fr_t *qprimes, size_t n_q, size_t bn, const fr_t *claims, size_t n_claims,
                    const fr_t *x0, const fr_t *x1, fr_t *out) {
    size_t len = (size_t)1 << bn;
    gate_t g;
    g.kind = gate_kind;
    if (ark) g.ark = *ark; else fr_set_zero(&g.ark);
    fr_t *eq = (fr_t *)malloc(len * sizeof(fr_t)), rho, tmp, res;
    orc_make_eq_table(claims, n_claims, qprimes, n_q, bn, eq, &rho);
    fr_set_zero(&res);
    for (size_t n = 0; n < len; n++) {
        const fr_t *buf[2] = {&x0[n], x1 ? &x1[n] : NULL};
        gate_eval(&g, &tmp, buf);
        fr_mul(&tmp, &tmp, &eq[n]);
        fr_add(&res, &res, &tmp);
    }
    *out = res;
    free(eq);
}

/* ------------------------------------------------------------------ circuit: the MiMC circuit */
/* examples/mimc.go:10-37: In lists. layer 0,1 inputs; 2 = Identity(0); i+3 = Cipher(Arks[i])(2, i==0 ? 1 : i+2) */
static int layer_in(int layer, int in[2]) {
    if (layer < 2) return 0;
    if (layer == 2) {
        in[0] = 0;
        return 1;
    }
    in[0] = 2;
    in[1] = layer == 3 ? 1 : layer - 1;
    return 2;
}
/* circuit/circuit.go:28-44 BuildCircuit: Out lists (sorted ascending by construction) */
static int layer_out(int layer, int *out /* up to 91 */) {
    int n = 0;
    for (int l = 0; l < N_LAYERS; l++) {
        int in[2];
        int k = layer_in(l, in);
        for (int j = 0; j < k; j++)
            if (in[j] == layer) out[n++] = l;
    }
    return n;
}
void orc_mimc_circuit_out(int layer, int *out, int *n_out) { *n_out = layer_out(layer, out); }

typedef struct {
    const gate_t *g;
    fr_t *res;
    const fr_t *in[2];
} ev_ctx_t;
static void ev_task(size_t start, size_t stop, void *c, int task) {
    (void)task;
    ev_ctx_t *ctx = (ev_ctx_t *)c;
    const fr_t *xs[2] = {ctx->in[0] + start, ctx->in[1] ? ctx->in[1] + start : NULL};
    gate_eval_batch(ctx->g, ctx->res + start, stop - start, xs);
}

/* circuit/assignment.go:12-32 Assign + circuit/circuit.go:48-64 Layer.Evaluate.
 * layers_out: N_LAYERS pointers, each to n entries (caller-allocated). */
void orc_mimc_assign(const fr_t *key, const fr_t *msg, size_t n, fr_t **layers) {
    memcpy(layers[0], key, n * sizeof(fr_t));
    memcpy(layers[1], msg, n * sizeof(fr_t));
    for (int l = 2; l < N_LAYERS; l++) {
        int in[2];
        int k = layer_in(l, in);
        gate_t g;
        g.kind = l == 2 ? GATE_IDENTITY : GATE_CIPHER;
        if (l > 2) g.ark = ARKS[l - 3]; else fr_set_zero(&g.ark);
        ev_ctx_t ctx = {&g, layers[l], {layers[in[0]], k > 1 ? layers[in[1]] : NULL}};
        parallelize(n, ev_task, &ctx);
    }
}

/* ------------------------------------------------------------------ gkr/ */
/* Flat proof vector exactly in the order of prover/gadget/hints.go:236-271 (GkrProofToVec):
 * all SumcheckProofs[l][k][j], then all Claims[l][j], then all QPrimes[l][j][k]; length 1006*bn+183
 * (hints.go:76-116).  Words are Montgomery form here (the Go in-memory values); orc_from_mont gives
 * the ToBigIntRegular values the hint writes. */
size_t orc_proof_vec_len(size_t bn) { return 1006 * bn + 183; }

typedef struct {
    size_t bn;
    fr_t *sumcheck[N_LAYERS];   /* bn * n_coeffs(layer) */
    fr_t *claims[N_LAYERS];     /* len(Out) */
    fr_t *qprimes[N_LAYERS];    /* len(Out or 1) * bn */
    int n_out[N_LAYERS];
    int n_q[N_LAYERS];
    int have[N_LAYERS];
} proof_t;

static int n_coeffs_of_layer(int l) { return l < 2 ? 0 : (l == 2 ? 3 : 9); }

static void proof_alloc(proof_t *p, size_t bn) {
    memset(p, 0, sizeof *p);
    p->bn = bn;
    int outs[N_LAYERS];
    for (int l = 0; l < N_LAYERS; l++) {
        p->n_out[l] = layer_out(l, outs);
        p->n_q[l] = l == N_LAYERS - 1 ? 1 : p->n_out[l];
        p->sumcheck[l] = (fr_t *)calloc(bn * (size_t)n_coeffs_of_layer(l) + 1, sizeof(fr_t));
        p->claims[l] = (fr_t *)calloc((size_t)p->n_out[l] + 1, sizeof(fr_t));
        p->qprimes[l] = (fr_t *)calloc((size_t)p->n_q[l] * bn + 1, sizeof(fr_t));
    }
}
static void proof_free(proof_t *p) {
    for (int l = 0; l < N_LAYERS; l++) {
        free(p->sumcheck[l]);
        free(p->claims[l]);
        free(p->qprimes[l]);
    }
}
static void proof_to_vec(const proof_t *p, fr_t *vec) {
    size_t c = 0, bn = p->bn;
    for (int l = 0; l < N_LAYERS; l++) {
        size_t n = bn * (size_t)n_coeffs_of_layer(l);
        memcpy(vec + c, p->sumcheck[l], n * sizeof(fr_t));
        c += n;
    }
    for (int l = 0; l < N_LAYERS; l++) {
        memcpy(vec + c, p->claims[l], (size_t)p->n_out[l] * sizeof(fr_t));
        c += (size_t)p->n_out[l];
    }
    for (int l = 0; l < N_LAYERS; l++) {
        size_t n = (size_t)p->n_q[l] * bn;
        memcpy(vec + c, p->qprimes[l], n * sizeof(fr_t));
        c += n;
    }
}
static void proof_from_vec(proof_t *p, const fr_t *vec) {
    size_t c = 0, bn = p->bn;
    for (int l = 0; l < N_LAYERS; l++) {
        size_t n = bn * (size_t)n_coeffs_of_layer(l);
        memcpy(p->sumcheck[l], vec + c, n * sizeof(fr_t));
        c += n;
    }
    for (int l = 0; l < N_LAYERS; l++) {
        memcpy(p->claims[l], vec + c, (size_t)p->n_out[l] * sizeof(fr_t));
        c += (size_t)p->n_out[l];
    }
    for (int l = 0; l < N_LAYERS; l++) {
        size_t n = (size_t)p->n_q[l] * bn;
        memcpy(p->qprimes[l], vec + c, n * sizeof(fr_t));
        c += n;
    }
}

/* gkr/prover.go:21-91 Prove + updateWithSumcheck over an assignment produced by orc_mimc_assign.
 * layers[0..92] are consumed (folded in place), layers[93] is left intact -- as in the reference.
 * Re-used inputs are deep-copied first (circuit/assignment.go:35-57). */
void orc_gkr_prove_mimc(fr_t **layers, size_t bn, const fr_t *qprime, fr_t *vec_out) {
    size_t n = (size_t)1 << bn;
    proof_t p;
    proof_alloc(&p, bn);
    memcpy(p.qprimes[N_LAYERS - 1], qprime, bn * sizeof(fr_t));
    fr_t *challenges = (fr_t *)malloc((bn + 1) * sizeof(fr_t));
    for (int layer = N_LAYERS - 1; layer >= 0; layer--) {
        int in[2], outs[N_LAYERS];
        int k = layer_in(layer, in);
        if (k == 0) break; /* circuit/circuit.go:70-79 IsInputLayer */
        fr_t *X[2] = {NULL, NULL};
        int copied[2] = {0, 0};
        for (int i = 0; i < k; i++) {
            layer_out(in[i], outs);
            if (outs[0] == layer) {
                X[i] = layers[in[i]];
            } else { /* DeepCopyLarge */
                X[i] = (fr_t *)malloc(n * sizeof(fr_t));
                memcpy(X[i], layers[in[i]], n * sizeof(fr_t));
                copied[i] = 1;
            }
        }
        fr_t fin[3];
        const fr_t *ark = layer > 2 ? &ARKS[layer - 3] : NULL;
        orc_sumcheck_prove(X[0], X[1], bn, p.qprimes[layer], (size_t)p.n_q[layer], p.claims[layer],
                           layer == N_LAYERS - 1 ? 0 : (size_t)p.n_out[layer], layer == 2 ? GATE_IDENTITY : GATE_CIPHER, ark,
                           p.sumcheck[layer], challenges, fin);
        for (int i = 0; i < k; i++) {
            int inp = in[i];
            int no = layer_out(inp, outs);
            int at = -1;
            for (int j = 0; j < no; j++)
                if (outs[j] == layer) at = j;
            if (at < 0) abort(); /* gkr/prover.go:83-85 */
            p.claims[inp][at] = fin[1 + i];
            memcpy(p.qprimes[inp] + (size_t)at * bn, challenges, bn * sizeof(fr_t));
            if (copied[i]) free(X[i]);
        }
    }
    proof_to_vec(&p, vec_out);
    proof_free(&p);
    free(challenges);
}

/* gkr/verifier.go:15-132 Verify for the MiMC circuit, from the flat vector.
 * returns 0 when accepted; 1000+layer on a sumcheck failure, 2000+layer on a qPrime mismatch,
 * 3000+layer on a final-claim mismatch, 4000+layer on an input-layer failure, 1 on qPrime mismatch. */
int orc_gkr_verify_mimc(const fr_t *vec, size_t bn, const fr_t *in0, const fr_t *in1, const fr_t *outputs, const fr_t *qprime) {
    size_t n = (size_t)1 << bn;
    proof_t p;
    proof_alloc(&p, bn);
    proof_from_vec(&p, vec);
    int rc = 0;
    fr_t *next_q = (fr_t *)malloc((bn + 1) * sizeof(fr_t));
    fr_t claims93[1];
    if (memcmp(qprime, p.qprimes[N_LAYERS - 1], bn * sizeof(fr_t)) != 0) {
        rc = 1;
        goto done;
    }
    orc_evaluate(outputs, n, qprime, bn, &claims93[0]);
    for (int layer = N_LAYERS - 1; layer >= 0 && !rc; layer--) {
        int in[2], outs[N_LAYERS];
        int k = layer_in(layer, in);
        if (k == 0) break;
        const fr_t *cl = layer == N_LAYERS - 1 ? claims93 : p.claims[layer];
        size_t ncl = layer == N_LAYERS - 1 ? 1 : (size_t)p.n_out[layer];
        fr_t next_claim, rho;
        if (orc_sumcheck_verify(cl, ncl, p.sumcheck[layer], bn, (size_t)n_coeffs_of_layer(layer), next_q, &next_claim, &rho)) {
            rc = 1000 + layer;
            break;
        }
        const fr_t *sub[2] = {NULL, NULL};
        for (int i = 0; i < k; i++) {
            int no = layer_out(in[i], outs), at = -1;
            for (int j = 0; j < no; j++)
                if (outs[j] == layer) at = j;
            if (memcmp(p.qprimes[in[i]] + (size_t)at * bn, next_q, bn * sizeof(fr_t)) != 0) rc = 2000 + layer;
            sub[i] = &p.claims[in[i]][at];
        }
        if (rc) break;
        gate_t g;
        g.kind = layer == 2 ? GATE_IDENTITY : GATE_CIPHER;
        if (layer > 2) g.ark = ARKS[layer - 3]; else fr_set_zero(&g.ark);
        fr_t expected, eq_eval;
        gate_eval(&g, &expected, sub);
        size_t nq = (size_t)p.n_q[layer];
        fr_t *tmp = (fr_t *)malloc(nq * sizeof(fr_t));
        for (size_t i = 0; i < nq; i++) orc_eval_eq(p.qprimes[layer] + i * bn, next_q, bn, &tmp[i]);
        orc_eval_univariate(tmp, nq, &rho, &eq_eval);
        free(tmp);
        fr_mul(&expected, &expected, &eq_eval);
        if (!fr_eq(&expected, &next_claim)) rc = 3000 + layer;
    }
    if (!rc) { /* gkr/verifier.go:120-132 testInitialRound */
        const fr_t *ins[2] = {in0, in1};
        for (int layer = 0; layer < 2; layer++) {
            fr_t actual;
            orc_evaluate(ins[layer], n, p.qprimes[layer], bn, &actual);
            if (!fr_eq(&actual, &p.claims[layer][0])) {
                rc = 4000 + layer;
                break;
            }
        }
    }
done:
    free(next_q);
    proof_free(&p);
    return rc;
}

/* Whole reference flow for the CPU baseline: Assign (circuit/assignment.go:12) + Prove (gkr/prover.go:21).
 * out93 (n entries) receives a[93]; vec_out the flat proof.  Allocates the 94-layer assignment like the
 * reference's pool does. */
int orc_assign_and_prove_mimc(const fr_t *key, const fr_t *msg, size_t bn, const fr_t *qprime, fr_t *out93, fr_t *vec_out) {
    size_t n = (size_t)1 << bn;
    fr_t *layers[N_LAYERS];
    for (int l = 0; l < N_LAYERS; l++) {
        layers[l] = (fr_t *)malloc(n * sizeof(fr_t));
        if (!layers[l]) return -1;
    }
    orc_mimc_assign(key, msg, n, layers);
    if (out93) memcpy(out93, layers[N_LAYERS - 1], n * sizeof(fr_t));
    orc_gkr_prove_mimc(layers, bn, qprime, vec_out);
    for (int l = 0; l < N_LAYERS; l++) free(layers[l]);
    return 0;
}
