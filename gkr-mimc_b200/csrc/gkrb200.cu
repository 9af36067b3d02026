// gkrb200: host driver + C ABI of the B200-native GKR-MiMC prover.  See include/gkrb200.h for the contract.
//
// Structure mirrors the reference call stack:
//   gkrb200_gkr_prove_mimc   ~ gkr.Prove / updateWithSumcheck        (gkr/prover.go:21-91)
//   Ctx::sumcheck            ~ sumcheck.Prove                        (sumcheck/prover.go:46-90)
//   Ctx::build_eq            ~ makeEqTable                           (sumcheck/prover.go:102-144)
//   round loop               ~ dispatchPartialEvals / InterpolateOnRange / GetChallenge / dispatchFolding
// The device does all O(N) work; the host keeps only the O(bn) serial transcript.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <linux/futex.h>
#include <nccl.h>
#include <sys/syscall.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/gkrb200.h"
#include "kernels.cuh"
#include "transcript.hpp"

using gkr::FrRaw;
namespace H = gkr::host;

static_assert(sizeof(FrRaw) == 32 && sizeof(H::Fr) == 32, "element image must be 32 bytes");

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(x)                                                                                             \
    do {                                                                                                        \
        cudaError_t e_ = (x);                                                                                   \
        if (e_ != cudaSuccess) return fail(GKRB200_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(x)             \
    do {                   \
        int rc_ = (x);     \
        if (rc_) return rc_; \
    } while (0)

extern "C" const char* gkrb200_last_error(void) { return g_err; }
extern "C" const char* gkrb200_version(void) { return "gkrb200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------ NCCL (dlopen'ed)
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::once_flag once;
    bool ok = false;
    bool load() {  // several pipelines (host threads) of one process initialise their communicators concurrently
        std::call_once(once, [this] {
            // torch's bundled libnccl.so.2 is reused when it is already mapped in the process (same SONAME)
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!h) return;
            GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
            CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
            CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
            AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
            Send = (decltype(Send))dlsym(h, "ncclSend");
            Recv = (decltype(Recv))dlsym(h, "ncclRecv");
            GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
            GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
            GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
            ok = GetUniqueId && CommInitRank && CommDestroy && AllGather && Send && Recv && GroupStart && GroupEnd && GetErrorString;
        });
        return ok;
    }
};
static Nccl g_nccl;
#define NCCL_TRY(x)                                                                                               \
    do {                                                                                                          \
        ncclResult_t r_ = (x);                                                                                    \
        if (r_ != ncclSuccess) return fail(GKRB200_ERR_COMM, "%s failed: %s", #x, g_nccl.GetErrorString(r_));     \
    } while (0)

// ------------------------------------------------------------------------------------------------ context
enum KClass { KC_ASSIGN = 0, KC_EQ = 1, KC_ROUND = 2, KC_FOLD = 3, KC_MULTIEQ = 4, KC_STAGING = 5, KC_MISC = 6, KC_N = 8 };

static constexpr int N_LAYERS = GKRB200_MIMC_LAYERS;
static constexpr int MAX_CLAIMS = 91;
static constexpr int ROUND_BLOCK = 128;
static constexpr int ROUND_MINB = 5;  // resident blocks per SM the round kernels are compiled for
static constexpr int MAX_EV = 9;
static constexpr int TAIL_MAX_FWD = 32;
static constexpr int CF_MINB1_FWD = 3;

static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct gkrb200_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int max_bn = 0;
    size_t cap = 0;  // entries per table slot

    // device arena
    FrRaw* arena = nullptr;
    FrRaw* layers = nullptr;     // [93][cap]: slot 0 = a[0] key (== a[2]), slot 1 = a[1] msg, slot s>=2 = a[s+1]
    FrRaw* eq = nullptr;         // [cap]
    FrRaw* scratch[3] = {};      // [cap/2] each, contiguous
    FrRaw* hi = nullptr;         // [MAX_CLAIMS][2^ceil(max_bn/2)]
    FrRaw* lo = nullptr;
    FrRaw* d_q = nullptr;        // [MAX_CLAIMS*max_bn]
    FrRaw* d_mults = nullptr;    // [MAX_CLAIMS]
    FrRaw* partials = nullptr;   // [max_grid][MAX_EV]
    unsigned int* ticket = nullptr;
    FrRaw* d_local = nullptr;    // [64] this rank's contribution (multi-GPU); [16..19) residual entries of the last fold
    FrRaw* d_all = nullptr;      // [8*64]
    FrRaw* d_resid = nullptr;    // [3][TAIL_MAX] residual tables after the last device fold
    int cf_blocks_per_sm1[2] = {CF_MINB1_FWD, CF_MINB1_FWD};  // out-of-line build (128-thread blocks), NM = 7 / 8
    int cf_blocks_per_sm_inl[2] = {2, 1};                       // inlined build (256-thread blocks)
    int cf_blocks_cap = 0;           // option: cap on the above (0 = none)
    uint32_t* partials_w = nullptr;  // 8 x 17 64-bit limb-column sums of the factored cipher round (zero between launches)
    int max_grid = 0;

    // pinned, device-mapped result slot: tagged 64-bit words only (see publish_word in kernels.cuh)
    FrRaw* h_result = nullptr;  // [512] (16 KiB: 8 ranks x 8 wide sums x 17 tagged 64-bit words fit)
    uint32_t seq = 0;
    FrRaw* d_chal = nullptr;        // [2] challenges delivered to the device by the round kernels' last blocks (leader mode)
    unsigned int* d_err = nullptr;  // device flag raised when a challenge did not arrive in time
    H::Fr* h_stage = nullptr;  // pinned staging for qprimes/mults uploads [MAX_CLAIMS*(max_bn+1)]

    // assignment state
    size_t n_local = 0;  // entries per table on this rank
    int bn = -1;         // global log2 batch of the current assignment
    bool sharded = false;

    // multi-GPU
    int rank = 0, world = 1, log_world = 0;
    ncclComm_t comm = nullptr;
    // Exchange window: one POSIX shared-memory segment per communicator, mapped by every rank of the box and registered
    // with CUDA as mapped pinned memory.  Rank g's kernels publish their round sums straight into slot [parity][g]
    // (the same tagged stores they use for the single-GPU result slot) and every rank's host thread reads all slots:
    // the per-round "all-reduce" of a few hundred bytes costs no collective, no extra launch and no device-side wait.
    // Two parities: a rank can run at most one exchange ahead of the slowest reader (it needs that reader's next
    // contribution before it can publish again).
    uint8_t* x_base = nullptr;  // host mapping
    uint8_t* x_dev = nullptr;   // device alias
    size_t x_size = 0;
    uint32_t xseq = 0;          // tag of the last exchange (identical on every rank: all ranks run the same exchanges)
    bool use_window = false;    // false: NCCL all-gather + publish kernel (GKRB200_OPT_EXCHANGE = 1, or no usable /dev/shm)
    static constexpr size_t XSLOT_DATA = 16384, XSLOT = XSLOT_DATA + 64;
    uint8_t* xslot_h(uint32_t tag, int g) const { return x_base + ((size_t)(tag & 1) * (size_t)world + (size_t)g) * XSLOT; }
    uint8_t* xslot_d(uint32_t tag, int g) const { return x_dev + ((size_t)(tag & 1) * (size_t)world + (size_t)g) * XSLOT; }
    bool windowed(int W) const { return W > 1 && use_window; }
    // One transcript per proof (SURVEY.md section 8e "not sharded: the transcript ... run once, broadcast"): the LEADER rank of the
    // communicator reads every rank's round sums from the window, runs interpolation + MimcHash and writes the challenge
    // back into the window as 8 tagged words ([2 parities] x 64 bytes after the sum slots); the last block of EVERY rank's
    // round kernel polls it and hands it to the next launch through device memory (ChalWait in kernels.cuh).  Follower
    // ranks enqueue a whole layer of launches and sleep until the leader posts the layer's header (round polynomials,
    // challenges, final claims: host-only part of the window, one ring entry per layer).
    int leader = 0;
    bool replicated_transcript = false;  // GKRB200_OPT_TRANSCRIPT = 1: every rank runs the transcript in lockstep (round 1 behaviour)
    uint64_t hseq = 0;                   // sequence number of the last layer header posted / consumed
    bool lead_mode(int W) const { return windowed(W) && !replicated_transcript; }
    size_t xchal_off() const { return 2 * (size_t)world * XSLOT; }
    volatile uint64_t* xchal_h(uint32_t tag) const { return (volatile uint64_t*)(x_base + xchal_off() + (size_t)(tag & 1) * 64); }
    const unsigned long long* xchal_d(uint32_t tag) const { return (const unsigned long long*)(x_dev + xchal_off() + (size_t)(tag & 1) * 64); }
    struct LayerHeader {
        volatile uint64_t seq;
        uint64_t pad[3];
        H::Fr fin[4];
        H::Fr challenges[32];
        H::Fr coeffs[32 * 9];
    };
    LayerHeader* xheader(int layer) const { return (LayerHeader*)(x_base + xchal_off() + 256) + layer; }
    static size_t xwindow_size(int world) { return 2 * (size_t)world * XSLOT + 256 + (size_t)GKRB200_MIMC_LAYERS * sizeof(LayerHeader); }
    gkr::ChalWait chal_wait(uint32_t tag, int k) const { return gkr::ChalWait{xchal_d(tag), d_chal + (k & 1), d_err, tag}; }
    void post_challenge(uint32_t tag, const H::Fr& r) const {
        volatile uint64_t* w = xchal_h(tag);
        for (int l = 0; l < 8; l++) w[l] = ((uint64_t)tag << 32) | (uint32_t)(r.l[l >> 1] >> (32 * (l & 1)));
    }
    int post_header(int layer, const H::Fr* coeffs, size_t n_coeffs, const H::Fr* challenges, int bn, const H::Fr* fin, int n_fin);
    int wait_header(int layer, H::Fr* coeffs, size_t n_coeffs, H::Fr* challenges, int bn, H::Fr* fin, int n_fin);
    int carve(size_t new_cap);
    // grow-only device staging for the host-table entry points (sumcheck_prove, fold, round_eval, fr_batch, convert): a
    // cudaMalloc/cudaFree pair per call costs tens of milliseconds at 64 MB (cudaFree synchronises the device)
    FrRaw* io_buf = nullptr;
    size_t io_cap = 0;
    int io_reserve(size_t n_elems, FrRaw** out) {
        if (n_elems > io_cap) {
            if (io_buf) {
                cudaStreamSynchronize(stream);
                cudaFree(io_buf);
                io_buf = nullptr;
                io_cap = 0;
            }
            cudaError_t e = cudaMalloc(&io_buf, n_elems * sizeof(FrRaw));
            if (e != cudaSuccess) {
                io_buf = nullptr;
                return fail(GKRB200_ERR_OOM, "cudaMalloc of a %.1f MiB staging buffer failed: %s", n_elems * 32.0 / (1 << 20), cudaGetErrorString(e));
            }
            io_cap = n_elems;
        }
        *out = io_buf;
        return 0;
    }
    int cap_bn = 0;  // log2(cap): what unsharded operations on this context can hold

    // instrumentation
    gkrb200_stats st{};
    bool profiling = false;
    struct Ev {
        cudaEvent_t a, b;
        int cls;
    };
    std::vector<Ev> ev_pool;
    size_t ev_used = 0;

    H::Lagrange lagrange;

    FrRaw* slot(int layer) const {  // device table of Assignment[layer]
        int s = layer == 2 ? 0 : (layer < 2 ? layer : layer - 1);
        return layers + (size_t)s * cap;
    }
    int eff_world() const { return sharded ? world : 1; }

    int ev_flush();
    void prof_begin(int cls);
    void prof_end();
    int wait_words(uint32_t seq, size_t n_words);
    int wait_words_at(const volatile uint64_t* w, uint32_t seq, size_t n_words);
    int upload(FrRaw* dst, const void* src, size_t n_elems);

    int build_eq(const H::Fr* qprimes, size_t n_q, int bn_local, const H::Fr* mults, FrRaw* out);
    int sumcheck(const FrRaw* x0, const FrRaw* x1, int bn_total, const H::Fr* qprimes, size_t n_q, const H::Fr* claims, size_t n_claims,
                 int gate, const H::Fr& ark, bool use_shards, H::Fr* proof_out, H::Fr* challenges_out, H::Fr* final_out,
                 const H::Fr* trusted_claim = nullptr);
    size_t par8_max_pairs = 8192;  // rounds with at most this many pairs spread one pair over 8 lanes
    size_t inline_min_pairs = 0;   // one-thread-per-pair rounds with at least this many pairs run the kernel with the inlined multiplier
    bool const_fold = true;        // option: folds by the constant-multiplier product when the host knows the challenge at launch
    bool force_generic = false;  // test hook: run cipher layers through the generic evaluate-at-9-points kernel
    int exchange_and_fetch_wide(int nm, int wl, int W, uint32_t tag, H::Fr* out);
    int mle_eval(const FrRaw* table, int bn_total, const H::Fr* point, bool use_shards, H::Fr* out);
    int fetch_residual(const FrRaw* const* cur, int ntab, size_t lres, bool folded_by_r, const H::Fr& r, int W, H::Fr (*tabs)[TAIL_MAX_FWD]);
    int residual_enqueue(const FrRaw* const* cur, int ntab, size_t lres, bool folded_by_r, const H::Fr& r, const FrRaw* r_dev, int W, uint32_t* tag_out);
    int residual_collect(int ntab, size_t lres, int W, uint32_t tag, H::Fr (*tabs)[TAIL_MAX_FWD]);
    int tail_len = TAIL_MAX_FWD;     // option: residual length (entries over all ranks) handed to the host
    int sumcheck_cf(const FrRaw* x0, const FrRaw* x1, int bn_total, const H::Fr* q, const H::Fr* trusted_claim, const H::Fr& ark, bool use_shards,
                    H::Fr* proof_out, H::Fr* challenges_out, H::Fr* final_out);
};

// ------------------------------------------------------------------------------------------------ profiling helpers
void gkrb200_ctx::prof_begin(int cls) {
    st.launches_total++;
    st.launches[cls]++;
    if (!profiling) return;
    if (ev_used == ev_pool.size()) {
        if (ev_pool.size() >= 4096) {
            ev_flush();
        } else {
            Ev e;
            cudaEventCreate(&e.a);
            cudaEventCreate(&e.b);
            ev_pool.push_back(e);
        }
    }
    ev_pool[ev_used].cls = cls;
    cudaEventRecord(ev_pool[ev_used].a, stream);
}
void gkrb200_ctx::prof_end() {
    if (!profiling) return;
    cudaEventRecord(ev_pool[ev_used].b, stream);
    ev_used++;
}
int gkrb200_ctx::ev_flush() {
    if (ev_used) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        for (size_t i = 0; i < ev_used; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev_pool[i].a, ev_pool[i].b);
            st.kernel_ms[ev_pool[i].cls] += ms;
        }
        ev_used = 0;
    }
    return 0;
}
#define LAUNCH(ctx, cls, kernel, grid, block, smem, ...)                           \
    do {                                                                           \
        (ctx)->prof_begin(cls);                                                    \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);           \
        (ctx)->prof_end();                                                         \
    } while (0)

// spin until every tagged 64-bit word of h_result carries `want` in its upper half (see publish_word in kernels.cuh)
int gkrb200_ctx::wait_words(uint32_t want, size_t n_words) { return wait_words_at((const volatile uint64_t*)h_result, want, n_words); }
int gkrb200_ctx::wait_words_at(const volatile uint64_t* w, uint32_t want, size_t n_words) {
    const double t0 = now_ms();
    unsigned spins = 0;
    for (;;) {
        bool all = true;
        for (size_t i = n_words; i-- > 0;)  // the last word usually lands last
            if ((uint32_t)(w[i] >> 32) != want) {
                all = false;
                break;
            }
        if (all) break;
        _mm_pause();
        if ((++spins & 0xfff) == 0) {
            cudaError_t q = cudaStreamQuery(stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) return fail(GKRB200_ERR_CUDA, "stream error while waiting: %s", cudaGetErrorString(q));
            if (now_ms() - t0 > 60000.0) return fail(GKRB200_ERR_CUDA, "timeout waiting for device result %u", want);
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    st.wait_ms += now_ms() - t0;
    return 0;
}

int gkrb200_ctx::upload(FrRaw* dst, const void* src, size_t n_elems) {
    st.h2d_bytes += n_elems * sizeof(FrRaw);
    CUDA_TRY(cudaMemcpyAsync(dst, src, n_elems * sizeof(FrRaw), cudaMemcpyHostToDevice, stream));
    return 0;
}

static inline int grid_for(size_t work_items, int block, int max_grid) {
    size_t g = (work_items + (size_t)block - 1) / (size_t)block;
    if (g < 1) g = 1;
    if (g > (size_t)max_grid) g = (size_t)max_grid;
    return (int)g;
}


// ------------------------------------------------------------------------------------------------ factored cipher round: launch plumbing
static constexpr int CF_BLOCK = 128;  // PAR == 8 kernels
static constexpr int CF_BLOCK1 = 128; // PAR == 1 kernels: 7-8 accumulators x 17 limbs x 128 threads = 61-70 KB of shared memory per block (64-thread blocks measured 30 % slower)
static constexpr int CF_MINB1 = 3;    // -> 3 blocks (12 warps) per SM
#ifndef GKR_CF_MINB_INL
#define GKR_CF_MINB_INL 2
#endif
static constexpr int CF_MINB_INL = GKR_CF_MINB_INL;  // resident 256-thread blocks per SM of the inlined-multiplier build (=> 128 registers per thread)
static constexpr int CF_MINB8 = 4;
static constexpr int CF_WL1 = 17;   // limbs per accumulator, PAR == 1 (plain 512-bit products summed)
static constexpr int CF_WL8 = 9;    // PAR == 8 (reduced products summed)
static constexpr size_t CF_SMEM_PAR8 = (8 * 9 + (CF_BLOCK / 32) * 8 * 9 + (CF_BLOCK / 8) * 8 * 9) * 4;
#ifndef GKR_CF_BLOCK_INL
#define GKR_CF_BLOCK_INL 256
#endif
static constexpr int CF_BLOCK_INL = GKR_CF_BLOCK_INL;  // inlined-multiplier kernels: two 256-thread blocks (16 warps) per SM, 16-limb shared accumulators
static inline size_t cf_smem_v(int nm, int variant) {  // variant: see CF_V_* below
    return variant == 1 ? CF_SMEM_PAR8 : (variant == 0 ? (size_t)nm * 16 * CF_BLOCK_INL * 4 : (size_t)nm * CF_WL1 * CF_BLOCK1 * 4);
}

typedef void (*cf_kernel_t)(const gkr::RoundCfArgs);
// variant: 0 = one thread per pair, multiplier inlined (big rounds); 1 = eight lanes per pair (small rounds);
//          2 = one thread per pair, multiplier out of line (mid-size rounds / A-B reference)
//          3 = variant 0 with the folds done by the constant-multiplier product (the challenge is known to the host at launch)
enum { CF_V_INL = 0, CF_V_PAR8 = 1, CF_V_CALL = 2, CF_V_INL_KFOLD = 3 };
static cf_kernel_t cf_kernel(bool fold, int nm, int variant) {
    using namespace gkr;
    if (variant == CF_V_INL_KFOLD) {
        if (!fold) variant = CF_V_INL;
        else return nm == 8 ? (cf_kernel_t)k_round_cf<true, 8, 1, CF_BLOCK_INL, CF_MINB_INL, true, true>
                            : (cf_kernel_t)k_round_cf<true, 7, 1, CF_BLOCK_INL, CF_MINB_INL, true, true>;
    }
#define CF_K1(F, N) k_round_cf<F, N, 1, CF_BLOCK_INL, CF_MINB_INL, true>
#define CF_K8(F, N) k_round_cf<F, N, 8, CF_BLOCK, CF_MINB8, false>
#define CF_KC(F, N) k_round_cf<F, N, 1, CF_BLOCK1, CF_MINB1, false>
    static const cf_kernel_t tab[2][2][3] = {{{CF_K1(false, 7), CF_K8(false, 7), CF_KC(false, 7)}, {CF_K1(false, 8), CF_K8(false, 8), CF_KC(false, 8)}},
                                             {{CF_K1(true, 7), CF_K8(true, 7), CF_KC(true, 7)}, {CF_K1(true, 8), CF_K8(true, 8), CF_KC(true, 8)}}};
#undef CF_K1
#undef CF_K8
#undef CF_KC
    return tab[fold ? 1 : 0][nm == 8 ? 1 : 0][variant];
}
static int set_cf_attrs() {
    for (int f = 0; f < 2; f++)
        for (int n = 7; n <= 8; n++)
            for (int v = 0; v < 4; v++)
                CUDA_TRY(cudaFuncSetAttribute(cf_kernel(f, n, v), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cf_smem_v(n, v == CF_V_INL_KFOLD ? CF_V_INL : v)));
    return 0;
}

// x (< 2^256, any) mod q
static H::Fr reduce_256(H::Fr x) {
    for (;;) {
        H::ull s0, s1, s2, s3;
        unsigned char b = _subborrow_u64(0, x.l[0], H::Q[0], &s0);
        b = _subborrow_u64(b, x.l[1], H::Q[1], &s1);
        b = _subborrow_u64(b, x.l[2], H::Q[2], &s2);
        b = _subborrow_u64(b, x.l[3], H::Q[3], &s3);
        if (b) return x;
        x = H::Fr{{s0, s1, s2, s3}};
    }
}
// Device sums (one per rank; tagged 64-bit words with the 32-bit limb in the low half) -> canonical field element.
//   wl == 9 : 288-bit sum of canonical Montgomery values       -> value mod q
//   wl == 17: 544-bit sum of PLAIN products x*y of Montgomery values -> REDC(sum) = sum * 2^-256 mod q, which is exactly the
//             sum of the Montgomery products fr.Mul would have produced one by one.
static H::Fr wide_to_fr(const volatile uint64_t* w, int wl, int n_ranks, size_t rank_stride_words) {
    uint64_t acc[20] = {0};
    for (int g = 0; g < n_ranks; g++)
        for (int l = 0; l < wl; l++) acc[l] += (uint32_t)w[(size_t)g * rank_stride_words + l];
    for (int l = 0; l < wl + 1; l++) {  // carry-normalise to 32-bit limbs
        acc[l + 1] += acc[l] >> 32;
        acc[l] &= 0xffffffffu;
    }
    const H::Fr R2{{H::R2[0], H::R2[1], H::R2[2], H::R2[3]}};
    auto limb64 = [&](int i) { return acc[2 * i] | (acc[2 * i + 1] << 32); };
    const H::Fr lo = reduce_256(H::Fr{{limb64(0), limb64(1), limb64(2), limb64(3)}});
    if (wl == 9) {
        const uint64_t hi = limb64(4);  // < 2^36
        return H::add(lo, H::mul(H::Fr{{hi, 0, 0, 0}}, R2));  // hi * 2^256 mod q = mont(hi, R^2)
    }
    const H::Fr mid = reduce_256(H::Fr{{limb64(4), limb64(5), limb64(6), limb64(7)}});
    const uint64_t top = limb64(8);  // < 2^36
    // sum = lo + 2^256 * (mid + 2^256 * top)  =>  sum * 2^-256 = lo * 2^-256 + mid + top * 2^256
    return H::add(H::add(H::mul(lo, H::Fr{{1, 0, 0, 0}}), mid), H::mul(H::Fr{{top, 0, 0, 0}}, R2));
}

// The constant-multiplier table of a fold challenge (fr_mul_const in fr_device.cuh): K_i = r * 2^(32i+64) * 2^-256 mod q.
// H::mul is the Montgomery product x*y*2^-256, so K_i = H::mul(r, c_i) with c_i the plain integer 2^(32i+64) mod q.
static gkr::FrConstMul const_mul_table(const H::Fr& r) {
    static const struct Pows {
        H::Fr c[8];
        Pows() {
            H::Fr v{{0, 1, 0, 0}};  // 2^64
            for (int i = 0; i < 8; i++) {
                c[i] = v;
                for (int b = 0; b < 32; b++) v = H::dbl(v);  // * 2^32 mod q
            }
        }
    } pw;
    gkr::FrConstMul t;
    for (int i = 0; i < 8; i++) {
        const H::Fr k = H::mul(r, pw.c[i]);
        for (int l = 0; l < 8; l++) t.k[i][l] = (uint32_t)(k.l[l >> 1] >> (32 * (l & 1)));
    }
    return t;
}

// test hook: the table the fold kernels receive for a challenge r (64 words: K_0..K_7, 8 x 32-bit limbs each, low limb first)
extern "C" int gkrb200_const_mul_table(const uint64_t* r, uint32_t* out64) {
    if (!r || !out64) return fail(GKRB200_ERR_ARG, "null argument");
    H::Fr rr;
    memcpy(&rr, r, 32);
    const gkr::FrConstMul t = const_mul_table(rr);
    memcpy(out64, t.k, sizeof t.k);
    return 0;
}

// ------------------------------------------------------------------------------------------------ init / free
static int init_ctx(gkrb200_ctx* c, int device, int max_bn, void* stream, int world);
static size_t shard_cap(int max_bn, int world) {  // entries per table a rank of `world` holds (floor: see gkrb200_comm_init)
    size_t want = (size_t)1 << max_bn;
    int lw = 0;
    while ((1 << lw) < world) lw++;
    if (world > 1) want = std::max<size_t>(want >> lw, std::min<size_t>(want, 64));
    return want;
}
extern "C" int gkrb200_init(gkrb200_ctx** out, int device, int max_bn, void* stream) { return gkrb200_init_shard(out, device, max_bn, stream, 1); }
extern "C" int gkrb200_init_shard(gkrb200_ctx** out, int device, int max_bn, void* stream, int world) {
    if (world < 1 || world > 8 || (world & (world - 1))) return fail(GKRB200_ERR_ARG, "bad world %d (a power of two <= 8)", world);
    if (!out || max_bn < 0 || max_bn > 26) return fail(GKRB200_ERR_ARG, "bad arguments to gkrb200_init (max_bn=%d)", max_bn);
    // the host transcript's multiplier is MULX/ADCX/ADOX assembly (fr_host.hpp): refuse to start on a CPU without them instead of SIGILL
    if (!__builtin_cpu_supports("adx") || !__builtin_cpu_supports("bmi2"))
        return fail(GKRB200_ERR_STATE, "this host CPU lacks ADX/BMI2, which the Fiat-Shamir transcript's field multiplier needs");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(GKRB200_ERR_CUDA, "no CUDA device available (%s); gkrb200 has no CPU fallback", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(GKRB200_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    gkrb200_ctx* c = new gkrb200_ctx();
    const int rc = init_ctx(c, device, max_bn, stream, world);
    if (rc) {
        gkrb200_free(c);  // releases whatever was built before the failure
        return rc;
    }
    *out = c;
    return 0;
}
static int init_ctx(gkrb200_ctx* c, int device, int max_bn, void* stream, int world) {
    c->device = device;
    c->max_bn = max_bn;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        return fail(GKRB200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    }
    c->n_sm = prop.multiProcessorCount;
    c->max_grid = c->n_sm * ROUND_MINB;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    TRY(c->carve(shard_cap(max_bn, world)));
    CUDA_TRY(cudaMalloc(&c->ticket, 128));
    CUDA_TRY(cudaMemset(c->ticket, 0, 128));
    c->d_err = c->ticket + 16;
    CUDA_TRY(cudaHostAlloc((void**)&c->h_result, 512 * sizeof(FrRaw), cudaHostAllocMapped));
    memset(c->h_result, 0, 512 * sizeof(FrRaw));
    CUDA_TRY(cudaHostAlloc((void**)&c->h_stage, (size_t)MAX_CLAIMS * (max_bn + 2) * sizeof(H::Fr), cudaHostAllocDefault));
    // opt in to the dynamic shared memory the round kernels need
    const int smem9 = 9 * 9 * ROUND_BLOCK * 4, smem3 = 3 * 17 * ROUND_BLOCK * 4;
    CUDA_TRY(cudaFuncSetAttribute(gkr::k_round<gkr::GATE_CIPHER, false, ROUND_BLOCK, ROUND_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem9));
    CUDA_TRY(cudaFuncSetAttribute(gkr::k_round<gkr::GATE_CIPHER, true, ROUND_BLOCK, ROUND_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem9));
    CUDA_TRY(cudaFuncSetAttribute(gkr::k_round<gkr::GATE_IDENTITY, false, ROUND_BLOCK, ROUND_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    CUDA_TRY(cudaFuncSetAttribute(gkr::k_round<gkr::GATE_IDENTITY, true, ROUND_BLOCK, ROUND_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    TRY(set_cf_attrs());
    for (int n = 7; n <= 8; n++) {  // resident blocks per SM of the PAR == 1 kernels (shared-memory bound): the grid is exactly one wave
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cf_kernel(true, n, CF_V_CALL), CF_BLOCK1, cf_smem_v(n, CF_V_CALL)));
        c->cf_blocks_per_sm1[n - 7] = nb > 0 ? nb : 1;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cf_kernel(true, n, CF_V_INL), CF_BLOCK_INL, cf_smem_v(n, CF_V_INL)));
        c->cf_blocks_per_sm_inl[n - 7] = nb > 0 ? nb : 1;
    }
    return 0;
}

static void comm_teardown(gkrb200_ctx* c);
// (Re)builds the device arena for tables of `new_cap` entries: one allocation, carved into the 93 layer tables, the eq table,
// the ping-pong scratch and the small per-layer buffers.  A sharded context holds 1/world of every table, so
// gkrb200_comm_init calls this again with cap = 2^max_bn / world (the arena must shrink with the shard).
int gkrb200_ctx::carve(size_t new_cap) {
    if (arena) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaFree(arena));
        arena = nullptr;
    }
    cap = new_cap;
    cap_bn = 0;
    while (((size_t)1 << (cap_bn + 1)) <= cap) cap_bn++;
    const size_t half = cap / 2 > 0 ? cap / 2 : 1;
    const int nsmall = (max_bn + 1) / 2;
    const size_t small = (size_t)1 << nsmall;
    const size_t pw_elems = ((size_t)8 * 17 * 8 + 31) / 32;  // 8 x 17 64-bit limb-column sums
    size_t total = 93 * cap + cap + 3 * half + 2 * MAX_CLAIMS * small + (size_t)MAX_CLAIMS * (max_bn + 1) + MAX_CLAIMS +
                   (size_t)max_grid * MAX_EV + 64 + 8 * 64 + 3 * 32 + 64 + pw_elems + 2;
    cudaError_t me = cudaMalloc(&arena, total * sizeof(FrRaw));
    if (me != cudaSuccess) {
        arena = nullptr;
        return fail(GKRB200_ERR_OOM, "cudaMalloc of %.2f GiB arena failed: %s", total * 32.0 / (1 << 30), cudaGetErrorString(me));
    }
    FrRaw* p = arena;
    layers = p; p += 93 * cap;
    eq = p; p += cap;
    for (int i = 0; i < 3; i++) { scratch[i] = p; p += half; }
    hi = p; p += MAX_CLAIMS * small;
    lo = p; p += MAX_CLAIMS * small;
    d_q = p; p += (size_t)MAX_CLAIMS * (max_bn + 1);
    d_mults = p; p += MAX_CLAIMS;
    partials = p; p += (size_t)max_grid * MAX_EV;
    d_local = p; p += 64;
    d_all = p; p += 8 * 64;
    d_resid = p; p += 3 * 32;
    partials_w = (uint32_t*)p; p += pw_elems;
    d_chal = p; p += 2;
    CUDA_TRY(cudaMemset(partials_w, 0, 8 * 17 * 8));
    bn = -1;  // whatever assignment the old arena held is gone
    return 0;
}

extern "C" void gkrb200_free(gkrb200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    comm_teardown(c);
    for (auto& e : c->ev_pool) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    cudaFree(c->arena);
    cudaFree(c->io_buf);
    cudaFree(c->ticket);
    cudaFreeHost(c->h_result);
    cudaFreeHost(c->h_stage);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int gkrb200_comm_unique_id(uint8_t id_out[128]) {
    if (!g_nccl.load()) return fail(GKRB200_ERR_COMM, "cannot load libnccl.so.2 (or a symbol is missing)");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return 0;
}
// releases the communicator and the exchange window of a context (comm_init on a context that already has them, and free)
static void comm_teardown(gkrb200_ctx* c) {
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    c->comm = nullptr;
    if (c->x_base) {
        cudaHostUnregister(c->x_base);
        munmap(c->x_base, c->x_size);
    }
    c->x_base = c->x_dev = nullptr;
    c->x_size = 0;
    c->use_window = false;
}

extern "C" int gkrb200_comm_init(gkrb200_ctx* c, int rank, int world, const uint8_t uid[128]) {
    if (!c || world < 1 || world > 8 || (world & (world - 1)) || rank < 0 || rank >= world)
        return fail(GKRB200_ERR_ARG, "bad rank/world %d/%d (world must be a power of two <= 8)", rank, world);
    if (world > 1 && !uid) return fail(GKRB200_ERR_ARG, "null nccl unique id");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    comm_teardown(c);  // a second comm_init on the same context replaces the first communicator
    c->rank = rank;
    c->world = world;
    c->leader = 0;
    c->log_world = 0;
    while ((1 << c->log_world) < world) c->log_world++;
    // a sharded context holds 1/world of every table: shrink (or restore) the arena accordingly.  64 entries is the floor
    // (batches of at most `world` hashes run unsharded, and tiny sharded batches stage the full inputs).
    if (shard_cap(c->max_bn, world) != c->cap) TRY(c->carve(shard_cap(c->max_bn, world)));
    if (world == 1) return 0;
    if (!g_nccl.load()) return fail(GKRB200_ERR_COMM, "cannot load libnccl.so.2 (or a symbol is missing)");
    ncclUniqueId id;
    memcpy(id.internal, uid, 128);
    // ---- exchange window (see gkrb200_ctx): opened by every rank BEFORE the NCCL rendezvous, unlinked by rank 0 after it
    char name[64];
    {
        uint64_t h = 1469598103934665603ull;  // FNV-1a of the unique id: every communicator gets its own segment
        for (int i = 0; i < 128; i++) h = (h ^ uid[i]) * 1099511628211ull;
        snprintf(name, sizeof name, "/gkrb200-%016llx", (unsigned long long)h);
    }
    const size_t xsize = gkrb200_ctx::xwindow_size(world);
    // every exit path below closes the descriptor and (rank 0) unlinks the segment
    struct Shm {
        int fd = -1;
        const char* name = nullptr;
        bool owner = false;
        ~Shm() {
            if (fd >= 0) close(fd);
            if (owner && name) shm_unlink(name);
        }
    } shm;
    shm.name = name;
    if (rank == 0) {
        shm_unlink(name);
        shm.fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        shm.owner = shm.fd >= 0;
        if (shm.fd >= 0 && ftruncate(shm.fd, (off_t)xsize) != 0) {  // ftruncate zero-fills: tag 0 is never used by an exchange
            close(shm.fd);
            shm.fd = -1;
        }
    }
    // NCCL's rendezvous doubles as the barrier "rank 0 has created the segment"
    NCCL_TRY(g_nccl.CommInitRank(&c->comm, world, id, rank));
    if (rank != 0) {
        shm.fd = shm_open(name, O_RDWR, 0600);
        struct stat sb;
        if (shm.fd >= 0 && (fstat(shm.fd, &sb) != 0 || (size_t)sb.st_size != xsize)) {
            close(shm.fd);
            shm.fd = -1;
        }
    }
    uint8_t ok = 0;
    if (shm.fd >= 0) {
        void* base = mmap(nullptr, xsize, PROT_READ | PROT_WRITE, MAP_SHARED, shm.fd, 0);
        if (base != MAP_FAILED) {
            void* dev = nullptr;
            if (cudaHostRegister(base, xsize, cudaHostRegisterMapped | cudaHostRegisterPortable) == cudaSuccess &&
                cudaHostGetDevicePointer(&dev, base, 0) == cudaSuccess && dev) {
                c->x_base = (uint8_t*)base;
                c->x_dev = (uint8_t*)dev;
                c->x_size = xsize;
                ok = 1;
            } else {
                cudaGetLastError();
                munmap(base, xsize);
            }
        }
    }
    // all ranks must agree on the exchange path: gather everybody's verdict (this is also the barrier before the unlink)
    CUDA_TRY(cudaMemcpyAsync(c->d_local, &ok, 1, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(g_nccl.AllGather(c->d_local, c->d_all, 1, ncclUint8, c->comm, c->stream));
    uint8_t oks[8] = {0};
    CUDA_TRY(cudaMemcpyAsync(oks, c->d_all, (size_t)world, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    bool all_ok = true;
    for (int g = 0; g < world; g++) all_ok = all_ok && oks[g] == 1;
    c->use_window = all_ok;
    c->xseq = 0;
    c->hseq = 0;
    // Leader mode pre-enqueues a whole layer of mutually waiting launches per context.  If several contexts' streams share one
    // hardware work queue, a waiting launch of one proof can hold back a ready launch of another whose sums the first proof's
    // leader (through another rank) is waiting for.  One queue per stream avoids that; say so once if the process did not ask for it.
    if (all_ok) {
        static std::once_flag warned;
        const char* mc = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
        if (!mc || atoi(mc) < 16)
            std::call_once(warned, [] {
                fprintf(stderr, "gkrb200: note: with several proofs in flight per GPU set CUDA_DEVICE_MAX_CONNECTIONS=32 before CUDA "
                                "initialises (one hardware work queue per stream; see INTEGRATION.md section 5)\n");
            });
    }
    if (!all_ok && getenv("GKRB200_VERBOSE")) fprintf(stderr, "gkrb200: rank %d: no shared exchange window (ok=%d), using NCCL all-gather\n", rank, (int)ok);
    return 0;
}

// Which rank runs the transcript of this communicator's proofs (default 0).  Set identically on every rank; with several
// proofs in flight (one communicator each) rotating the leader spreads the host work over the ranks' processes.
extern "C" int gkrb200_comm_set_leader(gkrb200_ctx* c, int leader_rank) {
    if (!c || leader_rank < 0 || leader_rank >= c->world) return fail(GKRB200_ERR_ARG, "leader rank %d out of range", leader_rank);
    c->leader = leader_rank;
    return 0;
}

extern "C" int gkrb200_comm_exchange_mode(gkrb200_ctx* c) { return (!c || c->world <= 1) ? -1 : (c->use_window ? 0 : 1); }

// ------------------------------------------------------------------------------------------------ K1 assign
static int assign_common(gkrb200_ctx* c, size_t n_local) {
    const int block = 128;
    const int grid = grid_for(n_local, block, c->n_sm * 16);
    LAUNCH(c, KC_ASSIGN, gkr::k_mimc_assign, grid, block, 0, c->slot(0), c->slot(1), c->slot(3), c->cap, n_local);
    c->st.fr_mul_assign += 364ull * n_local;
    CUDA_TRY(cudaGetLastError());
    return 0;
}
// shards: the call splits the table over the ranks of the communicator (the assignment); otherwise it must fit this rank's arena
static int check_n(gkrb200_ctx* c, size_t n, int* bn_out, bool shards = false) {
    *bn_out = 0;
    if (!c) return fail(GKRB200_ERR_ARG, "null context");
    if (n == 0 || (n & (n - 1))) return fail(GKRB200_ERR_ARG, "table size %zu is not a power of two", n);
    int bn = 0;
    while (((size_t)1 << bn) < n) bn++;
    const int limit = (shards && c->world > 1 && bn > c->log_world) ? std::min(c->max_bn, c->cap_bn + c->log_world) : c->cap_bn;
    if (bn > limit) return fail(GKRB200_ERR_OOM, "batch 2^%d exceeds the context capacity 2^%d", bn, limit);
    *bn_out = bn;
    return 0;
}

// Host key/msg (full tables, Go layout) -> this rank's layer-0 / layer-1 tables.  Sharded: the rank uploads only ITS
// contiguous 1/world slice of each table over PCIe (total H2D = one copy of the inputs, however many GPUs), de-interleaves it
// by owner (entry i belongs to rank i mod world) and the ranks swap the blocks with one grouped NCCL send/recv over NVLink --
// the only bandwidth-relevant transfer between GPUs on this path.
static int stage_inputs(gkrb200_ctx* c, const uint64_t* key, const uint64_t* msg, size_t n, int bn) {
    c->sharded = c->world > 1 && bn > c->log_world;
    c->bn = bn;
    if (!c->sharded) {
        c->n_local = n;
        TRY(c->upload(c->slot(0), key, n));
        TRY(c->upload(c->slot(1), msg, n));
        return 0;
    }
    const size_t W = (size_t)c->world, nl = n / W;
    c->n_local = nl;
    FrRaw* st_key = c->eq;          // cap entries
    FrRaw* st_msg = c->scratch[0];  // 3*cap/2 contiguous entries
    if (nl >= W && nl % W == 0) {
        const size_t blk = nl / W;
        TRY(c->upload(st_key, key + 4 * (size_t)c->rank * nl, nl));
        TRY(c->upload(st_msg, msg + 4 * (size_t)c->rank * nl, nl));
        FrRaw* snd_key = c->slot(3);  // layers 3 and 4 are free until the assignment kernel runs
        FrRaw* snd_msg = c->slot(4);
        const int grid = grid_for(nl, 256, c->n_sm * 8);
        LAUNCH(c, KC_STAGING, gkr::k_destripe, grid, 256, 0, st_key, snd_key, nl, (int)W);
        LAUNCH(c, KC_STAGING, gkr::k_destripe, grid, 256, 0, st_msg, snd_msg, nl, (int)W);
        CUDA_TRY(cudaGetLastError());
        const double t0 = now_ms();
        NCCL_TRY(g_nccl.GroupStart());
        for (size_t peer = 0; peer < W; peer++) {
            NCCL_TRY(g_nccl.Send(snd_key + peer * blk, blk * sizeof(FrRaw), ncclUint8, (int)peer, c->comm, c->stream));
            NCCL_TRY(g_nccl.Recv(c->slot(0) + peer * blk, blk * sizeof(FrRaw), ncclUint8, (int)peer, c->comm, c->stream));
            NCCL_TRY(g_nccl.Send(snd_msg + peer * blk, blk * sizeof(FrRaw), ncclUint8, (int)peer, c->comm, c->stream));
            NCCL_TRY(g_nccl.Recv(c->slot(1) + peer * blk, blk * sizeof(FrRaw), ncclUint8, (int)peer, c->comm, c->stream));
        }
        NCCL_TRY(g_nccl.GroupEnd());
        c->st.comm_ms += now_ms() - t0;
        c->st.launches_total++;
        c->st.launches[KC_STAGING]++;
        return 0;
    }
    // tiny sharded batches (n < world^2): every rank stages the full tables and keeps entries {i : i mod world == rank}
    TRY(c->upload(st_key, key, n));
    TRY(c->upload(st_msg, msg, n));
    const int grid = grid_for(nl, 256, c->n_sm * 8);
    LAUNCH(c, KC_STAGING, gkr::k_take_shard, grid, 256, 0, st_key, c->slot(0), nl, c->world, c->rank);
    LAUNCH(c, KC_STAGING, gkr::k_take_shard, grid, 256, 0, st_msg, c->slot(1), nl, c->world, c->rank);
    return 0;
}

extern "C" int gkrb200_mimc_assign(gkrb200_ctx* c, const uint64_t* key, const uint64_t* msg, size_t n, uint64_t* out93) {
    int bn;
    TRY(check_n(c, n, &bn, true));
    if (!key || !msg) return fail(GKRB200_ERR_ARG, "null input table");
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(stage_inputs(c, key, msg, n, bn));
    TRY(assign_common(c, c->n_local));
    if (out93) {
        CUDA_TRY(cudaMemcpyAsync(out93, c->slot(93), c->n_local * sizeof(FrRaw), cudaMemcpyDeviceToHost, c->stream));
        c->st.d2h_bytes += c->n_local * sizeof(FrRaw);
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gkrb200_mimc_assign_device(gkrb200_ctx* c, const void* d_key, const void* d_msg, size_t n) {
    int bn;
    TRY(check_n(c, n, &bn, true));
    if (!d_key || !d_msg) return fail(GKRB200_ERR_ARG, "null input table");
    CUDA_TRY(cudaSetDevice(c->device));
    c->sharded = c->world > 1 && bn > c->log_world;
    c->bn = bn;
    if (!c->sharded) {
        c->n_local = n;
        CUDA_TRY(cudaMemcpyAsync(c->slot(0), d_key, n * sizeof(FrRaw), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->slot(1), d_msg, n * sizeof(FrRaw), cudaMemcpyDeviceToDevice, c->stream));
    } else {
        c->n_local = n / (size_t)c->world;
        const int grid = grid_for(c->n_local, 256, c->n_sm * 8);
        LAUNCH(c, KC_STAGING, gkr::k_take_shard, grid, 256, 0, (const FrRaw*)d_key, c->slot(0), c->n_local, c->world, c->rank);
        LAUNCH(c, KC_STAGING, gkr::k_take_shard, grid, 256, 0, (const FrRaw*)d_msg, c->slot(1), c->n_local, c->world, c->rank);
    }
    TRY(assign_common(c, c->n_local));
    return 0;  // asynchronous on the context's stream
}

extern "C" int gkrb200_assign_layer_to_host(gkrb200_ctx* c, int layer, uint64_t* dst, size_t n) {
    if (!c || !dst || layer < 0 || layer >= N_LAYERS) return fail(GKRB200_ERR_ARG, "bad layer %d", layer);
    if (c->bn < 0) return fail(GKRB200_ERR_STATE, "no assignment in this context");
    if (n != c->n_local) return fail(GKRB200_ERR_ARG, "layer has %zu entries on this rank, caller asked for %zu", c->n_local, n);
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(dst, c->slot(layer), n * sizeof(FrRaw), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------ K2/K5 eq table
// out[x] = sum_j mults[j] * eq(q_j, x) over bn_local variables (mults == nullptr: single table with seed 1)
int gkrb200_ctx::build_eq(const H::Fr* qprimes, size_t n_q, int bnl, const H::Fr* mults, FrRaw* out) {
    if (n_q < 1 || n_q > (size_t)MAX_CLAIMS) return fail(GKRB200_ERR_ARG, "unsupported number of claims %zu (max %d)", n_q, MAX_CLAIMS);
    const int nh = bnl / 2, nl = bnl - nh;  // nl <= ceil(max_bn/2)
    const size_t n = (size_t)1 << bnl;
    // stage q's (and multipliers) through pinned memory
    size_t nq_el = n_q * (size_t)bnl;
    if (nq_el) memcpy(h_stage, qprimes, nq_el * sizeof(H::Fr));
    if (mults) memcpy(h_stage + nq_el, mults, n_q * sizeof(H::Fr));
    if (nq_el) TRY(upload(d_q, h_stage, nq_el));
    if (mults) TRY(upload(d_mults, h_stage + nq_el, n_q));
    const int cls = n_q > 1 ? KC_MULTIEQ : KC_EQ;
    LAUNCH(this, cls, gkr::k_eq_small, (int)(2 * n_q), 256, 0, d_q, bnl, nh, nl, mults ? d_mults : nullptr, hi, lo);
    const int grid = grid_for(n, 256, n_sm * 8);
    LAUNCH(this, cls, gkr::k_eq_expand, grid, gkr::EQX_BLOCK, 0, hi, lo, nh, nl, (int)n_q, out, n);
    CUDA_TRY(cudaGetLastError());
    // the staging buffer is reused by the next call: make sure the copies above are done
    return 0;
}

// ------------------------------------------------------------------------------------------------ host tail (multi-GPU residual)
// The last log2(world) rounds run on the gathered world-entry tables: at most 4 pairs.
static void host_round_eval(const H::Fr* eq, const H::Fr* x0, const H::Fr* x1, size_t len, int gate, const H::Fr& ark, H::Fr* evals) {
    const int nev = gate == gkr::GATE_CIPHER ? 9 : 3;
    const size_t mid = len / 2;
    for (int t = 0; t < nev; t++) evals[t] = H::zero();
    for (size_t x = 0; x < mid; x++) {
        H::Fr e = eq[x], de = H::sub(eq[x + mid], eq[x]);
        H::Fr s = x0[x], ds = H::sub(x0[x + mid], x0[x]);
        if (gate == gkr::GATE_CIPHER) {
            s = H::add(H::add(x1[x], ark), s);
            H::Fr s1 = H::add(H::add(x1[x + mid], ark), x0[x + mid]);
            ds = H::sub(s1, s);
        }
        for (int t = 0; t < nev; t++) {
            H::Fr g = s;
            if (gate == gkr::GATE_CIPHER) {
                H::Fr t2 = H::sqr(s);
                g = H::mul(H::sqr(H::mul(t2, s)), s);
            }
            evals[t] = H::add(evals[t], H::mul(e, g));
            e = H::add(e, de);
            s = H::add(s, ds);
        }
    }
}
static void host_fold(H::Fr* t, size_t len, const H::Fr& r) {
    const size_t mid = len / 2;
    for (size_t i = 0; i < mid; i++) t[i] = H::add(t[i], H::mul(r, H::sub(t[i + mid], t[i])));
}

// poly/eq.go:41-59 FoldedEqTable on the host (tiny tables of the tail), seeded with `seed`
static void host_eq_table(const H::Fr* q, int n, const H::Fr& seed, H::Fr* out) {
    out[0] = seed;
    for (int i = 0; i < n; i++) {
        const size_t step = (size_t)1 << (n - 1 - i);
        for (size_t j = 0; j < ((size_t)1 << i); j++) {
            const size_t J = j << (n - i);
            out[J + step] = H::mul(q[i], out[J]);
            out[J] = H::sub(out[J], out[J + step]);
        }
    }
}

static constexpr int TAIL_MAX = TAIL_MAX_FWD;  // largest residual table (entries, all ranks together) finished on the host

// Rounds k0..bn-1 on the host over the gathered residual tables of length rl = 2^(bn-k0) (at most TAIL_MAX entries):
// below that size a device round trip (launch + PCIe latency) costs more than the handful of multiplications.
// Same per-round steps as the device rounds: evaluate at 0..deg (sumcheck/algo.go:54-205), InterpolateOnRange,
// GetChallenge, Fold.  Leaves [Eq(r), X0(r), X1(r)] in final_out.
static void host_tail(const H::Lagrange& lagrange, H::Fr* te, H::Fr* t0v, H::Fr* t1v, size_t rl, int gate, const H::Fr& ark, int k0, int bn,
                      H::Fr* proof_out, H::Fr* challenges_out, H::Fr* final_out, gkrb200_stats& st) {
    const int nev = gate == gkr::GATE_CIPHER ? 9 : 3;
    const int nin = gate == gkr::GATE_CIPHER ? 2 : 1;
    H::Fr evals[MAX_EV];
    for (int k = k0; k < bn; k++) {
        const double t0 = now_ms();
        host_round_eval(te, t0v, t1v, rl, gate, ark, evals);
        H::Fr* coeffs = proof_out + (size_t)k * nev;
        lagrange.interpolate(evals, nev, coeffs);
        const H::Fr r = H::mimc_hash(coeffs, nev);
        challenges_out[k] = r;
        host_fold(te, rl, r);
        host_fold(t0v, rl, r);
        if (nin > 1) host_fold(t1v, rl, r);
        rl /= 2;
        st.transcript_ms += now_ms() - t0;
        st.rounds++;
    }
    final_out[0] = te[0];
    final_out[1] = t0v[0];
    if (nin > 1) final_out[2] = t1v[0];
}

// Residual tables for the host tail.  cur[t] (t < ntab) are this rank's device tables; when folded_by_r they have length
// 2*lres and are folded once more with the last device challenge (r, or *r_dev when the challenge lives in device memory),
// otherwise they have length lres and are taken as is.  residual_enqueue only ENQUEUES the device work (and, through the
// window, the publication of this rank's entries under a fresh exchange tag); residual_collect waits for every rank's entries:
// tabs[t][j*W + g] = rank g's entry j (strided sharding).
int gkrb200_ctx::residual_enqueue(const FrRaw* const* cur, int ntab, size_t lres, bool folded_by_r, const H::Fr& r, const FrRaw* r_dev, int W,
                                  uint32_t* tag_out) {
    if (folded_by_r) {
        gkr::FoldArgs f{};
        f.n_tables = ntab;
        for (int i = 0; i < ntab; i++) {
            f.src[i] = cur[i];
            f.dst[i] = d_resid + (size_t)i * lres;
        }
        f.half = lres;
        memcpy(&f.r, &r, 32);
        f.r_dev = r_dev;
        if (!r_dev) f.rk = const_mul_table(r);
        LAUNCH(this, KC_FOLD, gkr::k_fold, 1, 128, 0, f);
    } else {
        for (int i = 0; i < ntab; i++)
            CUDA_TRY(cudaMemcpyAsync(d_resid + (size_t)i * lres, cur[i], lres * sizeof(FrRaw), cudaMemcpyDeviceToDevice, stream));
    }
    *tag_out = 0;
    if (windowed(W)) {
        // every rank publishes its residual entries (tagged 32-bit limbs) into its slot of the exchange window
        const uint32_t tag = ++xseq;
        const int n_limbs = (int)((size_t)ntab * lres * 8);
        LAUNCH(this, KC_MISC, gkr::k_publish_tagged, 1, 256, 0, (const uint32_t*)d_resid, n_limbs, (unsigned long long*)xslot_d(tag, rank), tag);
        *tag_out = tag;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int gkrb200_ctx::residual_collect(int ntab, size_t lres, int W, uint32_t tag, H::Fr (*tabs)[TAIL_MAX]) {
    const size_t per_rank = (size_t)ntab * lres;
    if (windowed(W)) {
        const size_t n_limbs = per_rank * 8;
        for (int g = 0; g < W; g++) TRY(wait_words_at((const volatile uint64_t*)xslot_h(tag, g), tag, n_limbs));
        st.d2h_bytes += per_rank * (size_t)W * sizeof(FrRaw);
        for (int g = 0; g < W; g++) {
            const volatile uint64_t* w = (const volatile uint64_t*)xslot_h(tag, g);
            for (int t = 0; t < ntab; t++)
                for (size_t j = 0; j < lres; j++) {
                    const volatile uint64_t* e = w + ((size_t)t * lres + j) * 8;
                    H::Fr v;
                    for (int l = 0; l < 4; l++) v.l[l] = (uint64_t)(uint32_t)e[2 * l] | ((uint64_t)(uint32_t)e[2 * l + 1] << 32);
                    tabs[t][j * (size_t)W + g] = v;
                }
        }
        return 0;
    }
    const FrRaw* src = d_resid;
    if (W > 1) {
        const double t0 = now_ms();
        NCCL_TRY(g_nccl.AllGather(d_resid, d_all, per_rank * sizeof(FrRaw), ncclUint8, comm, stream));
        st.comm_ms += now_ms() - t0;
        src = d_all;
    }
    CUDA_TRY(cudaMemcpyAsync(h_stage, src, per_rank * (size_t)W * sizeof(FrRaw), cudaMemcpyDeviceToHost, stream));
    const double t1 = now_ms();
    CUDA_TRY(cudaStreamSynchronize(stream));
    st.wait_ms += now_ms() - t1;
    st.d2h_bytes += per_rank * (size_t)W * sizeof(FrRaw);
    for (int g = 0; g < W; g++)
        for (int t = 0; t < ntab; t++)
            for (size_t j = 0; j < lres; j++) tabs[t][j * (size_t)W + g] = h_stage[((size_t)g * ntab + t) * lres + j];
    return 0;
}
int gkrb200_ctx::fetch_residual(const FrRaw* const* cur, int ntab, size_t lres, bool folded_by_r, const H::Fr& r, int W, H::Fr (*tabs)[TAIL_MAX]) {
    uint32_t tag;
    TRY(residual_enqueue(cur, ntab, lres, folded_by_r, r, nullptr, W, &tag));
    return residual_collect(ntab, lres, W, tag, tabs);
}

// ---- layer headers (leader mode): the leader posts what the followers need to stay in step and to assemble the same proof
int gkrb200_ctx::post_header(int layer, const H::Fr* coeffs, size_t n_coeffs, const H::Fr* challenges, int bn_, const H::Fr* fin, int n_fin) {
    LayerHeader* h = xheader(layer);
    if (n_coeffs > 32 * 9 || bn_ > 32 || n_fin > 4) return fail(GKRB200_ERR_STATE, "internal: layer header overflow");
    if (n_coeffs) memcpy(h->coeffs, coeffs, n_coeffs * sizeof(H::Fr));
    if (bn_) memcpy(h->challenges, challenges, (size_t)bn_ * sizeof(H::Fr));
    memcpy(h->fin, fin, (size_t)n_fin * sizeof(H::Fr));
    std::atomic_thread_fence(std::memory_order_release);
    h->seq = ++hseq;
    // wake the followers sleeping on this header (a shared, non-private futex: the window is mapped by every rank's process)
    syscall(SYS_futex, (uint32_t*)&h->seq, FUTEX_WAKE, 0x7fffffff, nullptr, nullptr, 0);
    return 0;
}
// Followers sleep here (they have nothing to compute): short naps (or, with GKRB200_FUTEX set, a futex in the shared window), so a
// proof in flight costs ONE host core (its leader's), not one per rank.
int gkrb200_ctx::wait_header(int layer, H::Fr* coeffs, size_t n_coeffs, H::Fr* challenges, int bn_, H::Fr* fin, int n_fin) {
    LayerHeader* h = xheader(layer);
    const uint64_t want = ++hseq;
    const double t0 = now_ms();
    unsigned naps = 0;
    static const bool use_futex = getenv("GKRB200_FUTEX") != nullptr;  // opt-in: sleep on the header word instead of short naps
    for (;;) {
        const uint64_t have = h->seq;
        if (have == want) break;
        if (use_futex) {
            // sleep in the kernel until the leader posts (FUTEX_WAKE in post_header); the timeout only bounds the error checks below
            struct timespec ts = {0, 2000000};  // 2 ms
            syscall(SYS_futex, (uint32_t*)&h->seq, FUTEX_WAIT, (uint32_t)have, &ts, nullptr, 0);
        } else {
            struct timespec ts = {0, 20000};  // 20 us naps: the configuration measured on 2 and 8 x B200
            nanosleep(&ts, nullptr);
        }
        if ((++naps & (use_futex ? 0x3fu : 0x3ffu)) == 0) {
            cudaError_t q = cudaStreamQuery(stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) return fail(GKRB200_ERR_CUDA, "stream error while waiting for the leader: %s", cudaGetErrorString(q));
            if (now_ms() - t0 > 60000.0) return fail(GKRB200_ERR_COMM, "timeout waiting for the leader's header of layer %d (have %llu, want %llu)", layer,
                                                     (unsigned long long)h->seq, (unsigned long long)want);
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (n_coeffs) memcpy(coeffs, h->coeffs, n_coeffs * sizeof(H::Fr));
    if (bn_) memcpy(challenges, h->challenges, (size_t)bn_ * sizeof(H::Fr));
    memcpy(fin, h->fin, (size_t)n_fin * sizeof(H::Fr));
    st.follow_wait_ms += now_ms() - t0;
    return 0;
}

// ------------------------------------------------------------------------------------------------ sumcheck.Prove
// x0/x1: device tables of this rank (n_local = 2^(bn_total - log_world) entries when use_shards), never modified.
// qprimes: n_q * bn_total.  Returns bn_total*(nev) coefficients, bn_total challenges, 1+arity final claims.
//
// Two drivers share the code below.  LOCKSTEP (single GPU, or GKRB200_OPT_TRANSCRIPT = 1): launch round k, wait for its sums,
// run the transcript, pass r_k to launch k+1 by value.  LEADER MODE (sharded, default): every rank enqueues ALL device rounds
// of the sumcheck up front -- round k+1 reads r_k from device memory, where the last block of round k put it after polling
// the window (ChalWait) -- and only the communicator's leader rank runs the transcript loop; followers return right after
// enqueueing (outputs untouched) and pick the results up from the layer header.
int gkrb200_ctx::sumcheck(const FrRaw* x0, const FrRaw* x1, int bn, const H::Fr* qprimes, size_t n_q, const H::Fr* claims, size_t n_claims,
                          int gate, const H::Fr& ark, bool use_shards, H::Fr* proof_out, H::Fr* challenges_out, H::Fr* final_out,
                          const H::Fr* trusted_claim) {
    // sumcheck/prover.go:113-115
    if (n_claims != n_q && n_q > 1)
        return fail(GKRB200_ERR_ARG, "provided a multi-instance %zu but the number of claims does not match %zu", n_q, n_claims);
    if (gate == gkr::GATE_CIPHER && n_q == 1 && !force_generic)
        return sumcheck_cf(x0, x1, bn, qprimes, trusted_claim, ark, use_shards, proof_out, challenges_out, final_out);
    const int W = use_shards ? world : 1, LW = use_shards ? log_world : 0;
    const int bnl = bn - LW;  // rounds run on the device
    const int nev = gate == gkr::GATE_CIPHER ? 9 : 3;
    const int nin = gate == gkr::GATE_CIPHER ? 2 : 1;
    const int wl = gate == gkr::GATE_CIPHER ? 9 : 17;  // limbs per published sum: canonical values / unreduced sums of plain products
    const size_t n_local = (size_t)1 << bnl;
    const bool lead = lead_mode(W), am_leader = !lead || rank == leader;

    // ---- makeEqTable (sumcheck/prover.go:102-144)
    std::vector<H::Fr> mults;
    bool have_mults = false;
    if (n_q > 1 || W > 1) {
        mults.assign(n_q, H::one());
        have_mults = true;
        if (n_claims >= 1 && n_q > 1) {
            const double t0 = now_ms();
            const H::Fr rho = H::mimc_hash(claims, n_claims);  // prover.go:128 (once per sumcheck; every rank holds the claims)
            st.transcript_ms += now_ms() - t0;
            H::Fr m = rho;
            for (size_t j = 1; j < n_q; j++) {
                mults[j] = m;
                m = H::mul(m, rho);
            }
        }
        if (W > 1) {
            // this rank's slice of eq(q_j, .) carries the factor of its fixed low address bits:
            // global index = j*W + rank, address bit b of rank pairs with q[bn-1-b]
            for (size_t j = 0; j < n_q; j++) {
                const H::Fr* q = qprimes + j * (size_t)bn;
                H::Fr s = mults[j];
                for (int b = 0; b < LW; b++) {
                    const H::Fr& qb = q[bn - 1 - b];
                    s = H::mul(s, ((rank >> b) & 1) ? qb : H::sub(H::one(), qb));
                }
                mults[j] = s;
            }
        }
    }
    {
        // local table over the first bnl variables of every q_j
        std::vector<H::Fr> ql;
        const H::Fr* qsrc = qprimes;
        if (LW > 0) {
            ql.resize(n_q * (size_t)bnl);
            for (size_t j = 0; j < n_q; j++) memcpy(&ql[j * (size_t)bnl], qprimes + j * (size_t)bn, (size_t)bnl * sizeof(H::Fr));
            qsrc = ql.data();
        }
        TRY(build_eq(qsrc, n_q, bnl, have_mults ? mults.data() : nullptr, eq));
        CUDA_TRY(cudaStreamSynchronize(stream));  // h_stage / ql may be reused
    }

    // ---- rounds on the device
    const FrRaw* cur[3] = {eq, x0, x1};
    FrRaw* dstp[3] = {scratch[0], scratch[1], scratch[2]};
    size_t len = n_local;
    H::Fr evals[MAX_EV], r = H::zero();
    // the last rounds (residual table of at most tail_len entries over all ranks) are finished on the host
    int tail_bits = 0;
    while (((size_t)2 << tail_bits) * (size_t)W <= (size_t)tail_len && tail_bits < bnl) tail_bits++;
    const int kdev = bnl - tail_bits;
    uint32_t tags[32];
    auto launch_round = [&](int k) -> int {
        gkr::RoundArgs a{};
        const bool do_fold = k > 0;
        const size_t half = len / (do_fold ? 4 : 2);
        for (int i = 0; i < 3; i++) {
            a.src[i] = cur[i];
            a.dst[i] = dstp[i];
        }
        a.half = half;
        memcpy(&a.r, &r, 32);
        a.r_dev = (lead && do_fold) ? d_chal + ((k - 1) & 1) : nullptr;
        memcpy(&a.ark, &ark, 32);
        const uint32_t tag = windowed(W) ? ++xseq : ++seq;
        tags[k] = tag;
        a.red.partials = partials;
        a.red.ticket = ticket;
        a.red.result = windowed(W) ? (unsigned long long*)xslot_d(tag, rank) : (W > 1 ? (unsigned long long*)d_local : (unsigned long long*)h_result);
        a.red.seq = tag;
        a.red.chal = lead ? chal_wait(tag, k) : gkr::ChalWait{nullptr, nullptr, nullptr, 0};
        a.partials_w = partials_w;
        const int grid = grid_for(half, ROUND_BLOCK, max_grid);
        const size_t smem = (size_t)nev * wl * ROUND_BLOCK * 4;
        if (gate == gkr::GATE_CIPHER) {
            auto kf = do_fold ? gkr::k_round<gkr::GATE_CIPHER, true, ROUND_BLOCK, ROUND_MINB> : gkr::k_round<gkr::GATE_CIPHER, false, ROUND_BLOCK, ROUND_MINB>;
            LAUNCH(this, KC_ROUND, kf, grid, ROUND_BLOCK, smem, a);
            st.fr_mul_round += (uint64_t)half * (45 + (do_fold ? 6 : 0));
            st.bytes_round += (uint64_t)half * 32 * (do_fold ? (12 + 6) : 6);
        } else {
            auto kf = do_fold ? gkr::k_round<gkr::GATE_IDENTITY, true, ROUND_BLOCK, ROUND_MINB> : gkr::k_round<gkr::GATE_IDENTITY, false, ROUND_BLOCK, ROUND_MINB>;
            LAUNCH(this, KC_ROUND, kf, grid, ROUND_BLOCK, smem, a);
            st.fr_mul_round += (uint64_t)half * (3 * 64 + (do_fold ? 4 * 136 : 0)) / 136;  // three plain products (64 of 136 wide MACs each)
            st.bytes_round += (uint64_t)half * 32 * (do_fold ? (8 + 4) : 4);
        }
        CUDA_TRY(cudaGetLastError());
        if (do_fold) {
            for (int i = 0; i < 3; i++) cur[i] = dstp[i];
            len /= 2;
        }
        return 0;
    };
    auto transcript_round = [&](int k) -> int {
        TRY(exchange_and_fetch_wide(nev, wl, W, tags[k], evals));
        const double t0 = now_ms();
        H::Fr* coeffs = proof_out + (size_t)k * nev;
        lagrange.interpolate(evals, nev, coeffs);  // poly/lagrange.go:96
        r = H::mimc_hash(coeffs, nev);             // common/challenge.go:10
        challenges_out[k] = r;
        if (lead) post_challenge(tags[k], r);
        st.transcript_ms += now_ms() - t0;
        st.rounds++;
        return 0;
    };
    uint32_t rtag = 0;
    const size_t lres = (size_t)1 << tail_bits;
    if (lead) {
        for (int k = 0; k < kdev; k++) TRY(launch_round(k));
        TRY(residual_enqueue(cur, 1 + nin, lres, kdev > 0, r, kdev > 0 ? d_chal + ((kdev - 1) & 1) : nullptr, W, &rtag));
        if (!am_leader) return 0;
        for (int k = 0; k < kdev; k++) TRY(transcript_round(k));
    } else {
        for (int k = 0; k < kdev; k++) {
            TRY(launch_round(k));
            TRY(transcript_round(k));
        }
        TRY(residual_enqueue(cur, 1 + nin, lres, kdev > 0, r, nullptr, W, &rtag));
    }
    // ---- residual tables (folded with the last device challenge) -> host, gathered over the ranks; host finishes
    H::Fr tabs[3][TAIL_MAX];
    TRY(residual_collect(1 + nin, lres, W, rtag, tabs));
    host_tail(lagrange, tabs[0], tabs[1], tabs[2], lres * (size_t)W, gate, ark, kdev, bn, proof_out, challenges_out, final_out, st);
    return 0;
}

// Wide variant for the factored cipher round: nm 288- or 544-bit sums per rank, as tagged words.
int gkrb200_ctx::exchange_and_fetch_wide(int nm, int wl, int W, uint32_t tag, H::Fr* out) {
    const size_t words = (size_t)nm * wl;
    const volatile uint64_t* w = (const volatile uint64_t*)h_result;
    size_t rank_stride = words;
    if (windowed(W)) {
        for (int g = 0; g < W; g++) TRY(wait_words_at((const volatile uint64_t*)xslot_h(tag, g), tag, words));
        w = (const volatile uint64_t*)xslot_h(tag, 0);
        rank_stride = XSLOT / 8;
    } else {
        if (W > 1) {
            const double t0 = now_ms();
            NCCL_TRY(g_nccl.AllGather(d_local, d_all, words * 8, ncclUint8, comm, stream));
            st.launches_total++;
            st.launches[KC_MISC]++;
            gkr::k_publish_words<<<1, 128, 0, stream>>>((const unsigned long long*)d_all, (int)(words * W), (unsigned long long*)h_result);
            st.comm_ms += now_ms() - t0;
        }
        TRY(wait_words(tag, words * (size_t)W));
    }
    for (int i = 0; i < nm; i++) out[i] = wide_to_fr(w + (size_t)i * wl, wl, W, rank_stride);
    st.d2h_bytes += words * 8 * (size_t)W;
    return 0;
}

// montgomery-form small integers
static const H::Fr& binom7(int i) {
    static const H::Fr t[8] = {H::from_u64(1), H::from_u64(7), H::from_u64(21), H::from_u64(35), H::from_u64(35), H::from_u64(21), H::from_u64(7), H::from_u64(1)};
    return t[i];
}
// out[i] = 1/v[i] (Montgomery batch inversion: one field inversion per layer); ok[i] = false where v[i] == 0
static void batch_inverse(const H::Fr* v, int n, H::Fr* out, bool* ok) {
    H::Fr pre[32], acc = H::one();
    for (int i = 0; i < n; i++) {
        pre[i] = acc;
        ok[i] = !H::is_zero(v[i]);
        if (ok[i]) acc = H::mul(acc, v[i]);
    }
    H::Fr ia = H::inv(acc);
    for (int i = n - 1; i >= 0; i--) {
        if (!ok[i]) {
            out[i] = H::zero();
            continue;
        }
        out[i] = H::mul(ia, pre[i]);
        ia = H::mul(ia, v[i]);
    }
}

// ------------------------------------------------------------------------------------------------ sumcheck.Prove, cipher gate, one claim
// Same outputs as Ctx::sumcheck (sumcheck/prover.go:46-90) for gate = CipherGate and a single qPrime, through the factored
// round described above k_round_cf in kernels.cuh.  trusted_claim: the value of sum_x eq(q,x)*gate(X0,X1)(x) when the caller
// KNOWS it (gkr.Prove: Claims[layer][0], produced by this prover for a consistent assignment) or nullptr; with it round 0
// also skips the 8th coefficient sum.  Rounds k >= 1 always carry their own running claim S_{k-1}(r_{k-1}).
int gkrb200_ctx::sumcheck_cf(const FrRaw* x0, const FrRaw* x1, int bn, const H::Fr* q, const H::Fr* trusted_claim, const H::Fr& ark, bool use_shards,
                             H::Fr* proof_out, H::Fr* challenges_out, H::Fr* final_out) {
    const int W = use_shards ? world : 1, LW = use_shards ? log_world : 0;
    const int bnl = bn - LW;
    if (bnl > 26) return fail(GKRB200_ERR_ARG, "too many variables (%d)", bnl);
    const bool lead = lead_mode(W), am_leader = !lead || rank == leader;
    const H::Fr one = H::one();
    // this rank's slice of eq(q,.) carries the factor of its fixed low address bits
    H::Fr seed = one;
    for (int b = 0; b < LW; b++) {
        const H::Fr& qb = q[bn - 1 - b];
        seed = H::mul(seed, ((rank >> b) & 1) ? qb : H::sub(one, qb));
    }
    const int c = bnl / 2;  // variables in the low suffix table
    int tail_bits = 0;  // the last rounds (residual table of at most tail_len entries over all ranks) are finished on the host
    while (((size_t)2 << tail_bits) * (size_t)W <= (size_t)tail_len && tail_bits < bnl) tail_bits++;
    const int kdev = bnl - tail_bits;
    if (kdev > 0) {
        gkr::EqSuffixArgs ea{};
        memcpy(ea.q, q, (size_t)bnl * 32);
        memcpy(&ea.seed, &seed, 32);
        ea.n = bnl;
        ea.c = c;
        ea.outB = lo;
        ea.outA = hi;
        LAUNCH(this, KC_EQ, gkr::k_eq_suffix, 2, 256, 0, ea);
        CUDA_TRY(cudaGetLastError());
    }
    H::Fr qinv[32];
    bool qinv_ok[32];
    batch_inverse(q, bnl, qinv, qinv_ok);

    const FrRaw* cur[2] = {x0, x1};
    FrRaw* dstp[2] = {scratch[1], scratch[2]};
    H::Fr ck = one, cl = trusted_claim ? *trusted_claim : H::zero(), r = H::zero();
    H::Fr m[8], sc[8];
    uint32_t tags[32];
    int nms[32];
    bool par8s[32];
    // round k runs on 7 sums when its claim is known (every round but an untrusted round 0) and q_k is invertible
    auto launch_round = [&](int k) -> int {
        const int mk = bnl - 1 - k;  // variables of x'
        const size_t half = (size_t)1 << mk;
        const bool do_fold = k > 0;
        const int nm = ((k > 0 || trusted_claim != nullptr) && qinv_ok[k]) ? 7 : 8;
        const bool par8 = half <= par8_max_pairs;
        nms[k] = nm;
        par8s[k] = par8;
        gkr::RoundCfArgs a{};
        for (int i = 0; i < 2; i++) {
            a.src[i] = cur[i];
            a.dst[i] = dstp[i];
        }
        a.half = half;
        memcpy(&a.r, &r, 32);
        a.r_dev = (lead && do_fold) ? d_chal + ((k - 1) & 1) : nullptr;
        memcpy(&a.ark, &ark, 32);
        if (mk > c) {
            a.tA = hi + ((size_t)1 << (mk - c));
            a.tB = lo + ((size_t)1 << c);
        } else {
            a.tA = nullptr;
            a.tB = lo + ((size_t)1 << mk);
        }
        a.c = c;
        const uint32_t tag = windowed(W) ? ++xseq : ++seq;
        tags[k] = tag;
        a.red.partials = partials_w;
        a.red.ticket = ticket;
        a.red.result = windowed(W) ? (unsigned long long*)xslot_d(tag, rank) : (W > 1 ? (unsigned long long*)d_local : (unsigned long long*)h_result);
        a.red.seq = tag;
        a.red.chal = lead ? chal_wait(tag, k) : gkr::ChalWait{nullptr, nullptr, nullptr, 0};
        int variant = par8 ? CF_V_PAR8 : (half >= inline_min_pairs ? CF_V_INL : CF_V_CALL);
        const bool inl = variant == CF_V_INL;
        if (inl && do_fold && !a.r_dev && const_fold) {  // the host knows r: fold by the constant-multiplier product
            a.rk = const_mul_table(r);
            variant = CF_V_INL_KFOLD;
        }
        const int blk = par8 ? CF_BLOCK : (inl ? CF_BLOCK_INL : CF_BLOCK1);
        int bps1 = inl ? cf_blocks_per_sm_inl[nm - 7] : cf_blocks_per_sm1[nm - 7];
        if (cf_blocks_cap > 0 && bps1 > cf_blocks_cap) bps1 = cf_blocks_cap;
        const int grid = grid_for(par8 ? half * 8 : half, blk, n_sm * (par8 ? CF_MINB8 : bps1));  // at most one resident wave
        LAUNCH(this, KC_ROUND, cf_kernel(do_fold, nm, variant), grid, blk, cf_smem_v(nm, inl ? CF_V_INL : variant), a);
        CUDA_TRY(cudaGetLastError());
        // algorithmic multiplier work in units of one Montgomery product (136 wide multiply-adds): 9 (NM = 8: 11) full products,
        // NM plain 512-bit products of 64 multiply-adds each, the eq factor product and four folds
        st.fr_mul_round += (uint64_t)half * (136 * ((nm == 8 ? 11 : 9) + (mk > c ? 1 : 0) + (do_fold ? 4 : 0)) + 64 * nm) / 136;
        st.bytes_round += (uint64_t)half * 32 * (do_fold ? (8 + 4) : 4);
        if (do_fold) {
            cur[0] = dstp[0];
            cur[1] = dstp[1];
        }
        return 0;
    };
    auto transcript_round = [&](int k) -> int {
        const int nm = nms[k];
        TRY(exchange_and_fetch_wide(nm, par8s[k] ? CF_WL8 : CF_WL1, W, tags[k], m));
        const double t0 = now_ms();
        const H::Fr w0 = H::sub(one, q[k]), w1 = H::sub(q[k], w0);  // eq(q_k, t) = w0 + w1*t
        if (nm == 7) {
            // p(0) + p(1) = claim  =>  S(1) = (cl - w0*S(0)) / q_k ;  m_7 = S(1) - sum_{i<7} C(7,i) m_i
            H::Fr acc = H::mul(H::sub(cl, H::mul(w0, m[0])), qinv[k]);
            for (int i = 0; i < 7; i++) {
                sc[i] = i == 0 ? m[0] : H::mul(binom7(i), m[i]);
                acc = H::sub(acc, sc[i]);
            }
            sc[7] = acc;
        } else {
            for (int i = 0; i < 8; i++) sc[i] = (i == 0 || i == 7) ? m[i] : H::mul(binom7(i), m[i]);
        }
        // p(t) = ck * (w0 + w1 t) * S(t), coefficients low -> high (what InterpolateOnRange yields, poly/lagrange.go:96)
        const H::Fr cw0 = H::mul(ck, w0), cw1 = H::mul(ck, w1);
        H::Fr* coeffs = proof_out + (size_t)k * 9;
        coeffs[0] = H::mul(cw0, sc[0]);
        for (int j = 1; j < 8; j++) coeffs[j] = H::add(H::mul(cw0, sc[j]), H::mul(cw1, sc[j - 1]));
        coeffs[8] = H::mul(cw1, sc[7]);
        r = H::mimc_hash(coeffs, 9);  // common/challenge.go:10
        if (lead) post_challenge(tags[k], r);
        challenges_out[k] = r;
        cl = H::eval_univariate(sc, 8, r);  // next round's claim / c_{k+1}
        ck = H::mul(ck, H::add(w0, H::mul(w1, r)));
        st.transcript_ms += now_ms() - t0;
        st.rounds++;
        return 0;
    };
    uint32_t rtag = 0;
    const size_t lres = (size_t)1 << tail_bits;
    if (lead) {
        for (int k = 0; k < kdev; k++) TRY(launch_round(k));
        TRY(residual_enqueue(cur, 2, lres, kdev > 0, r, kdev > 0 ? d_chal + ((kdev - 1) & 1) : nullptr, W, &rtag));
        if (!am_leader) return 0;
        for (int k = 0; k < kdev; k++) TRY(transcript_round(k));
    } else {
        for (int k = 0; k < kdev; k++) {
            TRY(launch_round(k));
            TRY(transcript_round(k));
        }
        TRY(residual_enqueue(cur, 2, lres, kdev > 0, r, nullptr, W, &rtag));
    }
    // ---- residual X tables -> host (gathered over the ranks); the eq residual is closed-form: c_kdev * eq(q[kdev:], .)
    H::Fr tabs[3][TAIL_MAX];
    TRY(residual_collect(2, lres, W, rtag, tabs + 1));
    host_eq_table(q + kdev, bn - kdev, ck, tabs[0]);
    host_tail(lagrange, tabs[0], tabs[1], tabs[2], lres * (size_t)W, gkr::GATE_CIPHER, ark, kdev, bn, proof_out, challenges_out, final_out, st);
    return 0;
}

// ------------------------------------------------------------------------------------------------ circuit description
// examples/mimc.go:10-37: layer 2 = Identity(0); layer i+3 = Cipher(Arks[i])(2, i == 0 ? 1 : i+2)
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
// circuit/circuit.go:28-44 BuildCircuit: number of consumers of each layer and the position of a consumer in Out
static int n_out(int layer) { return layer == 2 ? 91 : (layer == N_LAYERS - 1 ? 0 : 1); }
static int pos_in_out(int producer, int consumer) { return producer == 2 ? consumer - 3 : 0; }
static int n_coeffs(int layer) { return layer < 2 ? 0 : (layer == 2 ? 3 : 9); }

extern "C" size_t gkrb200_proof_vec_len(int bn) { return (size_t)1006 * (size_t)bn + 183; }

// The library proves ONE circuit, examples.MimcCircuit() (examples/mimc.go:10-37).  The Go shim hands over the description of
// the circuit.Circuit it was called with and panics unless this returns 0 -- so gkr.Prove(c, a, qPrime) on any other circuit
// fails loudly instead of silently proving the MiMC one.  Checks what BuildCircuit (circuit/circuit.go:28-44) and
// IsInputLayer (:70-79) enforce as well: an input layer has no gate and no inputs and at most one consumer.
extern "C" int gkrb200_check_mimc_circuit(int n_layers, const int* n_in, const int* in_flat, const int* gate_kinds, const uint64_t* arks) {
    if (!n_in || !in_flat || !gate_kinds) return fail(GKRB200_ERR_ARG, "null circuit description");
    if (n_layers != N_LAYERS) return fail(GKRB200_ERR_ARG, "circuit has %d layers, the MiMC circuit has %d", n_layers, N_LAYERS);
    int consumers[N_LAYERS] = {0};
    size_t cur = 0;
    for (int l = 0; l < N_LAYERS; l++) {
        int want[2];
        const int k = layer_in(l, want);
        if (n_in[l] != k) return fail(GKRB200_ERR_ARG, "layer %d has %d inputs, the MiMC circuit has %d", l, n_in[l], k);
        for (int i = 0; i < k; i++) {
            const int src = in_flat[cur + (size_t)i];
            if (src != want[i]) return fail(GKRB200_ERR_ARG, "layer %d input %d is layer %d, the MiMC circuit wires layer %d", l, i, src, want[i]);
            consumers[src]++;
        }
        cur += (size_t)k;
        const int gate = k == 0 ? -1 : (l == 2 ? GKRB200_GATE_IDENTITY : GKRB200_GATE_CIPHER);
        if (gate_kinds[l] != gate) return fail(GKRB200_ERR_ARG, "layer %d has gate kind %d, the MiMC circuit has %d (-1 = input layer)", l, gate_kinds[l], gate);
        if (gate == GKRB200_GATE_CIPHER) {
            if (!arks) return fail(GKRB200_ERR_ARG, "null round constants");
            if (memcmp(arks + 4 * (size_t)l, &H::ARKS[l - 3], 32) != 0) return fail(GKRB200_ERR_ARG, "layer %d: the cipher gate's Ark is not hash.Arks[%d]", l, l - 3);
        }
    }
    for (int l = 0; l < 2; l++)  // circuit/circuit.go:37-41: an input layer feeds at most one layer
        if (consumers[l] > 1) return fail(GKRB200_ERR_ARG, "input layer %d has %d consumers", l, consumers[l]);
    return 0;
}

extern "C" int gkrb200_gkr_prove_mimc(gkrb200_ctx* c, const uint64_t* qprime, int bn, uint64_t* proof_vec_out, uint32_t flags) {
    if (!c || !proof_vec_out || (bn > 0 && !qprime)) return fail(GKRB200_ERR_ARG, "null argument");
    if (c->bn < 0) return fail(GKRB200_ERR_STATE, "gkr_prove_mimc called before mimc_assign");
    if (bn != c->bn) return fail(GKRB200_ERR_ARG, "inconsistent sizes : bn is %d but the assignment has 2^%d entries", bn, c->bn);
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t ubn = (size_t)bn;
    // gkr.Proof (gkr/prover.go:14-18) laid out per layer
    std::vector<std::vector<H::Fr>> sc(N_LAYERS), claims(N_LAYERS), qps(N_LAYERS);
    for (int l = 0; l < N_LAYERS; l++) {
        sc[l].assign(ubn * (size_t)n_coeffs(l), H::zero());
        claims[l].assign((size_t)n_out(l), H::zero());
        qps[l].assign((size_t)(l == N_LAYERS - 1 ? 1 : n_out(l)) * ubn, H::zero());
    }
    if (bn) memcpy(qps[N_LAYERS - 1].data(), qprime, ubn * sizeof(H::Fr));  // prover.go:31
    std::vector<H::Fr> challenges(ubn + 1);
    for (int layer = N_LAYERS - 1; layer >= 0; layer--) {
        int in[2];
        const int k = layer_in(layer, in);
        if (k == 0) break;  // input layer: the proof is complete (prover.go:35-39)
        const int gate = layer == 2 ? gkr::GATE_IDENTITY : gkr::GATE_CIPHER;
        const H::Fr ark = layer > 2 ? H::ARKS[layer - 3] : H::zero();
        const size_t n_q = layer == N_LAYERS - 1 ? 1 : (size_t)n_out(layer);
        const size_t n_cl = layer == N_LAYERS - 1 ? 0 : (size_t)n_out(layer);  // Claims[93] is nil (prover.go:27)
        H::Fr fin[4];
        // Claims[layer][0] was produced by this prover from the assignment it computed itself: it IS sum eq*gate of this layer
        const H::Fr* trusted = (gate == gkr::GATE_CIPHER && n_cl == 1) ? claims[layer].data() : nullptr;
        TRY(c->sumcheck(c->slot(in[0]), k > 1 ? c->slot(in[1]) : nullptr, bn, qps[layer].data(), n_q, claims[layer].data(), n_cl, gate, ark,
                        c->sharded, sc[layer].data(), challenges.data(), fin, trusted));
        if (c->sharded && c->lead_mode(c->world)) {
            // one transcript per proof: the leader posts this layer's round polynomials, challenges and final claims; the
            // followers (which only enqueued their kernels) pick them up and stay in step
            if (c->rank == c->leader) TRY(c->post_header(layer, sc[layer].data(), sc[layer].size(), challenges.data(), bn, fin, 1 + k));
            else TRY(c->wait_header(layer, sc[layer].data(), sc[layer].size(), challenges.data(), bn, fin, 1 + k));
        }
        for (int i = 0; i < k; i++) {  // prover.go:66-90
            const int at = pos_in_out(in[i], layer);
            claims[in[i]][(size_t)at] = fin[1 + i];
            if (bn) memcpy(&qps[in[i]][(size_t)at * ubn], challenges.data(), ubn * sizeof(H::Fr));
        }
    }
    if (c->sharded && c->lead_mode(c->world)) {  // a round kernel that gave up waiting for its challenge leaves a mark
        unsigned int derr = 0;
        CUDA_TRY(cudaMemcpyAsync(&derr, c->d_err, sizeof derr, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (derr) {
            cudaMemsetAsync(c->d_err, 0, sizeof derr, c->stream);
            return fail(GKRB200_ERR_COMM, "a round kernel timed out waiting for the leader's challenge");
        }
    }
    // GkrProofToVec order (prover/gadget/hints.go:236-271)
    H::Fr* out = (H::Fr*)proof_vec_out;
    size_t cur = 0;
    for (int l = 0; l < N_LAYERS; l++) { if (!sc[l].empty()) memcpy(out + cur, sc[l].data(), sc[l].size() * 32); cur += sc[l].size(); }
    for (int l = 0; l < N_LAYERS; l++) { if (!claims[l].empty()) memcpy(out + cur, claims[l].data(), claims[l].size() * 32); cur += claims[l].size(); }
    for (int l = 0; l < N_LAYERS; l++) { if (!qps[l].empty()) memcpy(out + cur, qps[l].data(), qps[l].size() * 32); cur += qps[l].size(); }
    if (cur != gkrb200_proof_vec_len(bn)) return fail(GKRB200_ERR_STATE, "internal: proof vector length %zu != %zu", cur, gkrb200_proof_vec_len(bn));
    if (flags & GKRB200_PROOF_REGULAR)
        for (size_t i = 0; i < cur; i++) out[i] = H::from_mont(out[i]);
    return 0;
}


// ------------------------------------------------------------------------------------------------ hint I/O
extern "C" int gkrb200_convert(gkrb200_ctx* c, const uint64_t* in, size_t n, uint64_t* out, int to_mont) {
    if (!c || (n && (!in || !out))) return fail(GKRB200_ERR_ARG, "null argument");
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    FrRaw* d = nullptr;
    TRY(c->io_reserve(n, &d));
    int rc = c->upload(d, in, n);
    if (!rc) {
        LAUNCH(c, KC_STAGING, gkr::k_convert, grid_for(n, 256, c->n_sm * 8), 256, 0, d, d, n, to_mont ? 1 : 0);
        cudaError_t e = cudaMemcpyAsync(out, d, n * sizeof(FrRaw), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(GKRB200_ERR_CUDA, "convert: %s", cudaGetErrorString(e));
        c->st.d2h_bytes += n * sizeof(FrRaw);
    }
    return rc;
}

extern "C" int gkrb200_mimc_assign_ex(gkrb200_ctx* c, const uint64_t* key, const uint64_t* msg, size_t n, uint64_t* out, uint32_t flags) {
    int bn;
    TRY(check_n(c, n, &bn, true));
    if (!key || !msg) return fail(GKRB200_ERR_ARG, "null input table");
    if (flags & ~(GKRB200_IO_INPUT_REGULAR | GKRB200_IO_OUTPUT_REGULAR | GKRB200_IO_OUTPUT_HASH)) return fail(GKRB200_ERR_ARG, "unknown flags 0x%x", flags);
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(stage_inputs(c, key, msg, n, bn));
    if (flags & GKRB200_IO_INPUT_REGULAR) {  // SetBigInt of every input on the device (prover/gadget/hints.go:202-205)
        const int grid = grid_for(c->n_local, 256, c->n_sm * 8);
        LAUNCH(c, KC_STAGING, gkr::k_convert, grid, 256, 0, c->slot(0), c->slot(0), c->n_local, 1);
        LAUNCH(c, KC_STAGING, gkr::k_convert, grid, 256, 0, c->slot(1), c->slot(1), c->n_local, 1);
    }
    TRY(assign_common(c, c->n_local));
    if (out) {
        const FrRaw* src = c->slot(93);
        if (flags & (GKRB200_IO_OUTPUT_REGULAR | GKRB200_IO_OUTPUT_HASH)) {
            LAUNCH(c, KC_STAGING, gkr::k_hash_out, grid_for(c->n_local, 256, c->n_sm * 8), 256, 0, c->slot(93), c->slot(0), c->slot(1), c->eq, c->n_local,
                   (flags & GKRB200_IO_OUTPUT_HASH) ? 1 : 0, (flags & GKRB200_IO_OUTPUT_REGULAR) ? 1 : 0);
            src = c->eq;
        }
        CUDA_TRY(cudaMemcpyAsync(out, src, c->n_local * sizeof(FrRaw), cudaMemcpyDeviceToHost, c->stream));
        c->st.d2h_bytes += c->n_local * sizeof(FrRaw);
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------ MultiLin.Evaluate on the device
// poly/multilin.go:59-66: fold a copy of the table with point[0], point[1], ... (MSB first).  The first fold is out of place
// (the table is never modified), the rest in place in a scratch table; the last <= 32 entries are finished on the host.
int gkrb200_ctx::mle_eval(const FrRaw* table, int bn, const H::Fr* point, bool use_shards, H::Fr* out) {
    const int W = use_shards ? world : 1, LW = use_shards ? log_world : 0;
    const int bnl = bn - LW;
    int tail_bits = 0;
    while (((size_t)2 << tail_bits) * (size_t)W <= (size_t)TAIL_MAX && tail_bits < bnl) tail_bits++;
    const int kdev = bnl - tail_bits;  // folds on the device; the last of them is done by fetch_residual
    const FrRaw* cur = table;
    FrRaw* dst = scratch[0];
    size_t len = (size_t)1 << bnl;
    for (int k = 0; k + 1 < kdev; k++) {
        gkr::FoldArgs f{};
        f.n_tables = 1;
        f.src[0] = cur;
        f.dst[0] = dst;
        f.half = len / 2;
        memcpy(&f.r, &point[k], 32);
        f.rk = const_mul_table(point[k]);
        LAUNCH(this, KC_FOLD, gkr::k_fold, grid_for(f.half, 256, n_sm * 8), 256, 0, f);
        cur = dst;
        len /= 2;
    }
    CUDA_TRY(cudaGetLastError());
    H::Fr tabs[1][TAIL_MAX];
    const FrRaw* curp[1] = {cur};
    TRY(fetch_residual(curp, 1, (size_t)1 << tail_bits, kdev > 0, kdev > 0 ? point[kdev - 1] : H::zero(), W, tabs));
    size_t rl = ((size_t)1 << tail_bits) * (size_t)W;
    for (int k = kdev; k < bn; k++) {
        host_fold(tabs[0], rl, point[k]);
        rl /= 2;
    }
    *out = tabs[0][0];
    return 0;
}

extern "C" int gkrb200_assign_layer_evaluate(gkrb200_ctx* c, int layer, const uint64_t* point, int bn, uint64_t* out) {
    if (!c || !out || layer < 0 || layer >= N_LAYERS || (bn > 0 && !point)) return fail(GKRB200_ERR_ARG, "bad argument");
    if (c->bn < 0) return fail(GKRB200_ERR_STATE, "no assignment in this context");
    if (bn != c->bn) return fail(GKRB200_ERR_ARG, "the assignment has 2^%d entries, the point has %d coordinates", c->bn, bn);
    CUDA_TRY(cudaSetDevice(c->device));
    H::Fr r;
    TRY(c->mle_eval(c->slot(layer), bn, (const H::Fr*)point, c->sharded, &r));
    memcpy(out, &r, 32);
    return 0;
}

extern "C" int gkrb200_mle_evaluate(gkrb200_ctx* c, const uint64_t* table, size_t n, const uint64_t* point, uint64_t* out) {
    int bn;
    TRY(check_n(c, n, &bn));
    if (!table || !out || (bn > 0 && !point)) return fail(GKRB200_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    FrRaw* d = c->eq;  // cap entries: the table is staged in the eq slot, folded into scratch[0]
    TRY(c->upload(d, table, n));
    H::Fr r;
    TRY(c->mle_eval(d, bn, (const H::Fr*)point, false, &r));
    memcpy(out, &r, 32);
    return 0;
}

// ------------------------------------------------------------------------------------------------ gkr.Verify (gkr/verifier.go:15-132)
// sumcheck.Verify (sumcheck/verifier.go:28-65): returns false (with a message) when a round check fails
static bool sumcheck_verify(const H::Fr* claims, size_t n_claims, const H::Fr* proof, int bn, int nco, H::Fr* challenges, H::Fr* final_claim,
                            H::Fr* recomb, char* why, size_t why_len) {
    // recombineMultiClaims (:58-65): always hashes the claims (the len < 1 guard of the reference cannot trigger)
    *recomb = H::mimc_hash(claims, n_claims);
    H::Fr expected = H::eval_univariate(claims, n_claims, *recomb);
    const H::Fr zero = H::zero(), one = H::one();
    for (int i = 0; i < bn; i++) {
        const H::Fr* p = proof + (size_t)i * nco;
        const H::Fr actual = H::add(H::eval_univariate(p, nco, zero), H::eval_univariate(p, nco, one));
        if (!H::eq(expected, actual)) {
            snprintf(why, why_len, "at round %d verifier eval at 0 + 1 differs from the expected value", i);
            return false;
        }
        challenges[i] = H::mimc_hash(p, nco);
        expected = H::eval_univariate(p, nco, challenges[i]);
    }
    *final_claim = expected;
    return true;
}

// gkr.Verify(c, proof, inputs, outputs, qPrime).  io == nullptr: inputs/outputs are the ones of the assignment held by the context
// (layers 0, 1, 93).  Otherwise io[0..2] = the caller's key, msg and outputs tables (host, 2^bn entries each), evaluated on the
// device from the caller's bytes -- the form the hint's self-check needs (prover/gadget/hints.go:225-229: the solver's outputs, not
// the prover's own recomputation, so a faulty assignment cannot vouch for itself).
static int gkr_verify_common(gkrb200_ctx* c, const uint64_t* proof_vec, int bn, const uint64_t* qprime, uint32_t flags, const uint64_t* const* io) {
    if (!c || !proof_vec || (bn > 0 && !qprime)) return fail(GKRB200_ERR_ARG, "null argument");
    if (io) {
        if (!io[0] || !io[1] || !io[2]) return fail(GKRB200_ERR_ARG, "null input/output table");
        if (bn > c->cap_bn) return fail(GKRB200_ERR_OOM, "tables of 2^%d entries exceed this context's capacity 2^%d (a sharded context holds 1/world of a table)", bn, c->cap_bn);
    } else {
        if (c->bn < 0) return fail(GKRB200_ERR_STATE, "gkr_verify_mimc needs the assignment (inputs and outputs) in the context");
        if (bn != c->bn) return fail(GKRB200_ERR_ARG, "inconsistent sizes : bn is %d but the assignment has 2^%d entries", bn, c->bn);
    }
    CUDA_TRY(cudaSetDevice(c->device));
    auto eval_io = [&](int which, int layer, const H::Fr* point, H::Fr* out) -> int {
        if (!io) return c->mle_eval(c->slot(layer), bn, point, c->sharded, out);
        TRY(c->upload(c->eq, io[which], (size_t)1 << bn));  // staged in the eq slot, folded into scratch[0]
        return c->mle_eval(c->eq, bn, point, false, out);
    };
    const size_t ubn = (size_t)bn, total = gkrb200_proof_vec_len(bn);
    std::vector<H::Fr> vec(total);
    memcpy(vec.data(), proof_vec, total * 32);
    if (flags & GKRB200_PROOF_REGULAR)
        for (auto& v : vec) v = H::to_mont(v);
    // slice the flat vector (prover/gadget/hints.go:275-317 GkrProofFromVec order)
    const H::Fr *sc[N_LAYERS], *cl[N_LAYERS], *qp[N_LAYERS];
    size_t cur = 0;
    for (int l = 0; l < N_LAYERS; l++) { sc[l] = vec.data() + cur; cur += ubn * (size_t)n_coeffs(l); }
    for (int l = 0; l < N_LAYERS; l++) { cl[l] = vec.data() + cur; cur += (size_t)n_out(l); }
    for (int l = 0; l < N_LAYERS; l++) { qp[l] = vec.data() + cur; cur += (size_t)(l == N_LAYERS - 1 ? 1 : n_out(l)) * ubn; }
    // gkr/verifier.go:25: the initial qPrime must be the one in the proof
    if (bn && memcmp(qprime, qp[N_LAYERS - 1], ubn * 32) != 0) return fail(GKRB200_ERR_VERIFY, "initial qPrime does not match with the proof");
    // :36 the initial claim is the output MLE at qPrime (the prover does not send it)
    H::Fr out_claim;
    TRY(eval_io(2, N_LAYERS - 1, (const H::Fr*)qprime, &out_claim));
    std::vector<H::Fr> challenges(ubn + 1), tmp(MAX_CLAIMS);
    char why[160];
    for (int layer = N_LAYERS - 1; layer >= 0; layer--) {
        int in[2];
        const int k = layer_in(layer, in);
        if (k == 0) break;  // :40-44
        const bool top = layer == N_LAYERS - 1;
        const size_t n_cl = top ? 1 : (size_t)n_out(layer), n_q = top ? 1 : (size_t)n_out(layer);
        const H::Fr* claims = top ? &out_claim : cl[layer];
        H::Fr final_claim, recomb;
        if (!sumcheck_verify(claims, n_cl, sc[layer], bn, n_coeffs(layer), challenges.data(), &final_claim, &recomb, why, sizeof why))
            return fail(GKRB200_ERR_VERIFY, "error at sumcheck layer %d : %s", layer, why);
        // testSumcheck (:61-117)
        H::Fr sub[2];
        for (int i = 0; i < k; i++) {
            const int at = pos_in_out(in[i], layer);
            if (bn && memcmp(qp[in[i]] + (size_t)at * ubn, challenges.data(), ubn * 32) != 0)
                return fail(GKRB200_ERR_VERIFY, "mismatch for qPrimes between sumcheck and proof at layer %d", layer);
            sub[i] = cl[in[i]][at];
        }
        H::Fr expected;
        if (layer == 2) {
            expected = sub[0];  // IdentityGate.Eval (circuit/gates/copy.go:20-22)
        } else {            // CipherGate.Eval (circuit/gates/cipher.go:45-55)
            const H::Fr t = H::add(H::add(sub[0], sub[1]), H::ARKS[layer - 3]);
            const H::Fr t2 = H::sqr(t);
            expected = H::mul(H::sqr(H::mul(t2, t)), t);
        }
        for (size_t i = 0; i < n_q; i++) tmp[i] = H::eval_eq(qp[layer] + i * ubn, challenges.data(), ubn);
        const H::Fr eq_eval = H::eval_univariate(tmp.data(), n_q, recomb);
        expected = H::mul(expected, eq_eval);
        if (!H::eq(expected, final_claim))
            return fail(GKRB200_ERR_VERIFY, "the expected claim and the final claim of the sumcheck do not match for layer %d", layer);
    }
    // testInitialRound (:120-132) for the two input layers
    for (int layer = 0; layer < 2; layer++) {
        H::Fr actual;
        TRY(eval_io(layer, layer, qp[layer], &actual));
        if (!H::eq(actual, cl[layer][0])) return fail(GKRB200_ERR_VERIFY, "input layer check failed (layer %d)", layer);
    }
    return 0;
}
extern "C" int gkrb200_gkr_verify_mimc(gkrb200_ctx* c, const uint64_t* proof_vec, int bn, const uint64_t* qprime, uint32_t flags) {
    return gkr_verify_common(c, proof_vec, bn, qprime, flags, nullptr);
}
extern "C" int gkrb200_gkr_verify_mimc_io(gkrb200_ctx* c, const uint64_t* proof_vec, int bn, const uint64_t* qprime, uint32_t flags, const uint64_t* key,
                                          const uint64_t* msg, const uint64_t* outputs) {
    const uint64_t* io[3] = {key, msg, outputs};
    return gkr_verify_common(c, proof_vec, bn, qprime, flags, io);
}

// ------------------------------------------------------------------------------------------------ sumcheck.Prove on host tables
static int sumcheck_prove_common(gkrb200_ctx* c, const void* X0, const void* X1, bool on_device, int bn, const uint64_t* qprimes, size_t n_q,
                                 const uint64_t* claims, size_t n_claims, int gate_kind, const uint64_t* ark, uint64_t* proof_out,
                                 uint64_t* challenges_out, uint64_t* final_claims_out) {
    if (!c || !X0 || bn < 0 || n_q < 1 || (bn > 0 && !qprimes) || !final_claims_out) return fail(GKRB200_ERR_ARG, "bad argument");
    if (gate_kind != GKRB200_GATE_IDENTITY && gate_kind != GKRB200_GATE_CIPHER)
        return fail(GKRB200_ERR_ARG, "gate kind %d cannot cross the ABI (only IdentityGate and CipherGate)", gate_kind);
    if (gate_kind == GKRB200_GATE_CIPHER && !X1) return fail(GKRB200_ERR_ARG, "cipher gate needs two input tables");
    if (bn > c->cap_bn) return fail(GKRB200_ERR_OOM, "table 2^%d exceeds the context capacity 2^%d", bn, c->cap_bn);
    if (n_claims && !claims) return fail(GKRB200_ERR_ARG, "null claims");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t n = (size_t)1 << bn;
    const FrRaw *d0 = (const FrRaw*)X0, *d1 = (const FrRaw*)X1;
    FrRaw* d = nullptr;
    int rc = 0;
    if (!on_device) {
        TRY(c->io_reserve(2 * n, &d));
        rc = c->upload(d, X0, n);
        if (!rc && gate_kind == GKRB200_GATE_CIPHER) rc = c->upload(d + n, X1, n);
        d0 = d;
        d1 = d + n;
    }
    H::Fr a = H::zero();
    if (ark) memcpy(&a, ark, 32);
    if (!rc)
        rc = c->sumcheck(d0, gate_kind == GKRB200_GATE_CIPHER ? d1 : nullptr, bn, (const H::Fr*)qprimes, n_q, (const H::Fr*)claims, n_claims,
                         gate_kind, a, false, (H::Fr*)proof_out, (H::Fr*)challenges_out, (H::Fr*)final_claims_out);
    if (d) cudaStreamSynchronize(c->stream);  // the caller's host tables may be released as soon as the call returns
    return rc;
}
extern "C" int gkrb200_sumcheck_prove(gkrb200_ctx* c, const uint64_t* X0, const uint64_t* X1, int bn, const uint64_t* qprimes, size_t n_q,
                                      const uint64_t* claims, size_t n_claims, int gate_kind, const uint64_t* ark, uint64_t* proof_out,
                                      uint64_t* challenges_out, uint64_t* final_claims_out) {
    return sumcheck_prove_common(c, X0, X1, false, bn, qprimes, n_q, claims, n_claims, gate_kind, ark, proof_out, challenges_out, final_claims_out);
}
extern "C" int gkrb200_sumcheck_prove_device(gkrb200_ctx* c, const void* d_X0, const void* d_X1, int bn, const uint64_t* qprimes, size_t n_q,
                                             const uint64_t* claims, size_t n_claims, int gate_kind, const uint64_t* ark, uint64_t* proof_out,
                                             uint64_t* challenges_out, uint64_t* final_claims_out) {
    return sumcheck_prove_common(c, d_X0, d_X1, true, bn, qprimes, n_q, claims, n_claims, gate_kind, ark, proof_out, challenges_out, final_claims_out);
}

// ------------------------------------------------------------------------------------------------ building blocks
extern "C" int gkrb200_eq_table(gkrb200_ctx* c, const uint64_t* qprimes, size_t n_q, int bn, const uint64_t* multipliers, uint64_t* out) {
    if (!c || !out || bn < 0 || (bn > 0 && !qprimes)) return fail(GKRB200_ERR_ARG, "bad argument");
    if (bn > c->cap_bn) return fail(GKRB200_ERR_OOM, "table 2^%d exceeds the context capacity 2^%d", bn, c->cap_bn);
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(c->build_eq((const H::Fr*)qprimes, n_q, bn, (const H::Fr*)multipliers, c->eq));
    CUDA_TRY(cudaMemcpyAsync(out, c->eq, ((size_t)1 << bn) * sizeof(FrRaw), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gkrb200_fold(gkrb200_ctx* c, const uint64_t* table, size_t n, const uint64_t* r, uint64_t* out) {
    int bn;
    TRY(check_n(c, n, &bn));
    if (!table || !r || !out || n < 2) return fail(GKRB200_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    FrRaw* d = nullptr;
    TRY(c->io_reserve(n + n / 2, &d));
    int rc = c->upload(d, table, n);
    if (!rc) {
        gkr::FoldArgs f{};
        f.n_tables = 1;
        f.src[0] = d;
        f.dst[0] = d + n;
        f.half = n / 2;
        memcpy(&f.r, r, 32);
        {
            H::Fr rr;
            memcpy(&rr, r, 32);
            f.rk = const_mul_table(rr);
        }
        LAUNCH(c, KC_FOLD, gkr::k_fold, grid_for(n / 2, 256, c->n_sm * 8), 256, 0, f);
        cudaError_t e = cudaMemcpyAsync(out, d + n, (n / 2) * sizeof(FrRaw), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(GKRB200_ERR_CUDA, "fold: %s", cudaGetErrorString(e));
    }
    return rc;
}

extern "C" int gkrb200_round_eval(gkrb200_ctx* c, const uint64_t* eq, const uint64_t* X0, const uint64_t* X1, size_t n, int gate_kind,
                                  const uint64_t* ark, uint64_t* evals_out) {
    int bn;
    TRY(check_n(c, n, &bn));
    if (!eq || !X0 || !evals_out || n < 2) return fail(GKRB200_ERR_ARG, "bad argument");
    if (gate_kind != GKRB200_GATE_IDENTITY && gate_kind != GKRB200_GATE_CIPHER) return fail(GKRB200_ERR_ARG, "bad gate kind %d", gate_kind);
    if (gate_kind == GKRB200_GATE_CIPHER && !X1) return fail(GKRB200_ERR_ARG, "cipher gate needs two input tables");
    CUDA_TRY(cudaSetDevice(c->device));
    FrRaw* d = nullptr;
    TRY(c->io_reserve(3 * n, &d));
    int rc = c->upload(d, eq, n);
    if (!rc) rc = c->upload(d + n, X0, n);
    if (!rc && gate_kind == GKRB200_GATE_CIPHER) rc = c->upload(d + 2 * n, X1, n);
    if (!rc) {
        gkr::RoundArgs a{};
        a.src[0] = d;
        a.src[1] = d + n;
        a.src[2] = d + 2 * n;
        a.half = n / 2;
        if (ark) memcpy(&a.ark, ark, 32);
        ++c->seq;
        a.red.partials = c->partials;
        a.red.ticket = c->ticket;
        a.red.result = (unsigned long long*)c->h_result;
        a.red.seq = c->seq;
        const int nev = gate_kind == GKRB200_GATE_CIPHER ? 9 : 3, wl = gate_kind == GKRB200_GATE_CIPHER ? 9 : 17;
        a.partials_w = c->partials_w;
        const int grid = grid_for(a.half, ROUND_BLOCK, c->max_grid);
        const size_t smem = (size_t)nev * wl * ROUND_BLOCK * 4;
        auto kf = gate_kind == GKRB200_GATE_CIPHER ? gkr::k_round<gkr::GATE_CIPHER, false, ROUND_BLOCK, ROUND_MINB> : gkr::k_round<gkr::GATE_IDENTITY, false, ROUND_BLOCK, ROUND_MINB>;
        LAUNCH(c, KC_ROUND, kf, grid, ROUND_BLOCK, smem, a);
        H::Fr ev[MAX_EV];
        rc = c->exchange_and_fetch_wide(nev, wl, 1, c->seq, ev);
        if (!rc) memcpy(evals_out, ev, (size_t)nev * 32);
    }
    cudaStreamSynchronize(c->stream);
    return rc;
}

extern "C" int gkrb200_fr_batch(gkrb200_ctx* c, int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
    if (!c || !a || !out || op < 0 || op > 4 || (op < 3 && !b) || n == 0) return fail(GKRB200_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    FrRaw* d = nullptr;
    TRY(c->io_reserve(3 * n, &d));
    int rc = c->upload(d, a, n);
    if (!rc && b) rc = c->upload(d + n, b, n);
    if (!rc) {
        LAUNCH(c, KC_MISC, gkr::k_fr_batch, grid_for(n, 256, c->n_sm * 8), 256, 0, op, d, d + n, d + 2 * n, n);
        cudaError_t e = cudaMemcpyAsync(out, d + 2 * n, n * sizeof(FrRaw), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(GKRB200_ERR_CUDA, "fr_batch: %s", cudaGetErrorString(e));
    }
    return rc;
}

// ------------------------------------------------------------------------------------------------ host transcript pieces
extern "C" int gkrb200_mimc_hash(const uint64_t* in, size_t n, uint64_t* out) {
    if (!out || (n && !in)) return fail(GKRB200_ERR_ARG, "null argument");
    H::Fr r = H::mimc_hash((const H::Fr*)in, n);
    memcpy(out, &r, 32);
    return 0;
}
extern "C" int gkrb200_interpolate(const uint64_t* evals, size_t n, uint64_t* coeffs_out) {
    static const H::Lagrange lag;
    if (!evals || !coeffs_out || n < 1 || n > (size_t)H::Lagrange::MAX_DOMAIN) return fail(GKRB200_ERR_ARG, "domain size %zu unsupported (1..12)", n);
    H::Fr tmp[H::Lagrange::MAX_DOMAIN];
    lag.interpolate((const H::Fr*)evals, (int)n, tmp);
    memcpy(coeffs_out, tmp, n * 32);
    return 0;
}
extern "C" int gkrb200_to_montgomery(const uint64_t* in, size_t n, uint64_t* out) {
    if (n && (!in || !out)) return fail(GKRB200_ERR_ARG, "null argument");
    for (size_t i = 0; i < n; i++) {
        H::Fr v;
        memcpy(&v, in + 4 * i, 32);
        v = H::to_mont(v);
        memcpy(out + 4 * i, &v, 32);
    }
    return 0;
}
extern "C" int gkrb200_from_montgomery(const uint64_t* in, size_t n, uint64_t* out) {
    if (n && (!in || !out)) return fail(GKRB200_ERR_ARG, "null argument");
    for (size_t i = 0; i < n; i++) {
        H::Fr v;
        memcpy(&v, in + 4 * i, 32);
        v = H::from_mont(v);
        memcpy(out + 4 * i, &v, 32);
    }
    return 0;
}

// hash.Arks[round] (hash/ark.go:13-337), the constant examples.MimcCircuit() gives layer round+3 (examples/mimc.go:29)
extern "C" int gkrb200_mimc_ark(int round, uint64_t* out) {
    if (!out || round < 0 || round >= H::MIMC_ROUNDS) return fail(GKRB200_ERR_ARG, "mimc_ark: round %d out of range (0..%d)", round, H::MIMC_ROUNDS - 1);
    memcpy(out, &H::ARKS[round], 32);
    return 0;
}
// poly.EvalUnivariate (poly/lagrange.go:31-39), poly.EvalEq (poly/eq.go:19-32), scalar fr.Element methods, sumcheck.Verify
extern "C" int gkrb200_eval_univariate(const uint64_t* coeffs, size_t n, const uint64_t* x, uint64_t* out) {
    if (!coeffs || !x || !out || n < 1) return fail(GKRB200_ERR_ARG, "null argument or empty polynomial");
    std::vector<H::Fr> c(n);
    memcpy(c.data(), coeffs, n * 32);
    H::Fr xx;
    memcpy(&xx, x, 32);
    const H::Fr r = H::eval_univariate(c.data(), n, xx);
    memcpy(out, &r, 32);
    return 0;
}
extern "C" int gkrb200_eval_eq(const uint64_t* q, const uint64_t* h, size_t n, uint64_t* out) {
    if (!out || (n && (!q || !h))) return fail(GKRB200_ERR_ARG, "null argument");
    std::vector<H::Fr> a(n), b(n);
    if (n) memcpy(a.data(), q, n * 32), memcpy(b.data(), h, n * 32);
    const H::Fr r = H::eval_eq(a.data(), b.data(), n);
    memcpy(out, &r, 32);
    return 0;
}
extern "C" int gkrb200_fr_scalar(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    if (!a || !out || (op <= 2 && !b)) return fail(GKRB200_ERR_ARG, "null argument");
    H::Fr x, y = H::zero(), r;
    memcpy(&x, a, 32);
    if (b) memcpy(&y, b, 32);
    switch (op) {
        case 0: r = H::mul(x, y); break;
        case 1: r = H::add(x, y); break;
        case 2: r = H::sub(x, y); break;
        case 3: {  // the cipher S-box x^7 (circuit/gates/cipher.go:51-54)
            const H::Fr x2 = H::sqr(x);
            r = H::mul(H::sqr(H::mul(x2, x)), x);
            break;
        }
        case 4: r = H::is_zero(x) ? H::zero() : H::inv(x); break;  // fr.Element.Inverse (0 -> 0)
        default: return fail(GKRB200_ERR_ARG, "fr_scalar: unknown op %d", op);
    }
    memcpy(out, &r, 32);
    return 0;
}
extern "C" int gkrb200_sumcheck_verify(const uint64_t* claims, size_t n_claims, const uint64_t* proof, int bn, int n_coeffs_per_round,
                                       uint64_t* challenges_out, uint64_t* final_claim_out, uint64_t* recomb_out) {
    if (!claims || n_claims < 1 || bn < 0 || (bn > 0 && (!proof || !challenges_out)) || !final_claim_out)
        return fail(GKRB200_ERR_ARG, "null argument or no claim");
    if (n_coeffs_per_round < 1 || n_coeffs_per_round > H::Lagrange::MAX_DOMAIN) return fail(GKRB200_ERR_ARG, "bad round polynomial length %d", n_coeffs_per_round);
    std::vector<H::Fr> cl(n_claims), pr((size_t)bn * (size_t)n_coeffs_per_round), ch((size_t)bn + 1);
    memcpy(cl.data(), claims, n_claims * 32);
    if (bn) memcpy(pr.data(), proof, pr.size() * 32);
    H::Fr fin, recomb;
    char why[160];
    if (!sumcheck_verify(cl.data(), n_claims, pr.data(), bn, n_coeffs_per_round, ch.data(), &fin, &recomb, why, sizeof why))
        return fail(GKRB200_ERR_VERIFY, "%s", why);
    if (bn) memcpy(challenges_out, ch.data(), (size_t)bn * 32);
    memcpy(final_claim_out, &fin, 32);
    if (recomb_out) memcpy(recomb_out, &recomb, 32);
    return 0;
}

// ------------------------------------------------------------------------------------------------ instrumentation
extern "C" int gkrb200_stats_reset(gkrb200_ctx* c) {
    if (!c) return fail(GKRB200_ERR_ARG, "null context");
    c->ev_flush();
    memset(&c->st, 0, sizeof c->st);
    return 0;
}
extern "C" int gkrb200_stats_get(gkrb200_ctx* c, gkrb200_stats* out) {
    if (!c || !out) return fail(GKRB200_ERR_ARG, "null argument");
    TRY(c->ev_flush());
    *out = c->st;
    return 0;
}
extern "C" int gkrb200_set_profiling(gkrb200_ctx* c, int on) {
    if (!c) return fail(GKRB200_ERR_ARG, "null context");
    c->ev_flush();
    c->profiling = on != 0;
    return 0;
}

extern "C" int gkrb200_set_option(gkrb200_ctx* c, int option, long value) {
    if (!c) return fail(GKRB200_ERR_ARG, "null context");
    switch (option) {
        case GKRB200_OPT_GENERIC_CIPHER: c->force_generic = value != 0; return 0;
        case GKRB200_OPT_HOST_TAIL_LEN:
            if (value < 1 || value > TAIL_MAX || (value & (value - 1))) return fail(GKRB200_ERR_ARG, "host tail length must be a power of two in 1..%d", TAIL_MAX);
            c->tail_len = (int)value;
            return 0;
        case GKRB200_OPT_PAR8_MAX_PAIRS:
            if (value < 0) return fail(GKRB200_ERR_ARG, "bad value %ld", value);
            c->par8_max_pairs = (size_t)value;
            return 0;
        case GKRB200_OPT_CF_BLOCKS_PER_SM:
            if (value < 0 || value > 32) return fail(GKRB200_ERR_ARG, "bad value %ld", value);
            c->cf_blocks_cap = (int)value;
            return 0;
        case GKRB200_OPT_INLINE_MIN_PAIRS:
            if (value < 0) return fail(GKRB200_ERR_ARG, "bad value %ld", value);
            c->inline_min_pairs = (size_t)value;
            return 0;
        case GKRB200_OPT_CONST_FOLD:
            if (value != 0 && value != 1) return fail(GKRB200_ERR_ARG, "bad value %ld", value);
            c->const_fold = value == 1;
            return 0;
        case GKRB200_OPT_TRANSCRIPT:
            if (value != 0 && value != 1) return fail(GKRB200_ERR_ARG, "bad value %ld", value);
            c->replicated_transcript = value == 1;
            return 0;
        case GKRB200_OPT_EXCHANGE:
            if (value != 0 && value != 1) return fail(GKRB200_ERR_ARG, "bad value %ld", value);
            if (value == 0 && c->world > 1 && !c->x_base) return fail(GKRB200_ERR_STATE, "this communicator has no exchange window");
            c->use_window = value == 0 && c->x_base != nullptr;
            return 0;
        default: return fail(GKRB200_ERR_ARG, "unknown option %d", option);
    }
}

#ifdef GKR_TRACE
extern "C" int gkrb200_trace_get(long long* out16) {
    return cudaMemcpyFromSymbol(out16, gkr::g_trace, 16 * sizeof(long long)) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int gkrb200_microbench(gkrb200_ctx* c, int kind_and_occ, int iters, double* rate_out, double* ms_out) {
    // low byte: kind; next byte (optional): warps per SM to allow (occupancy limited through dynamic shared memory)
    const int kind = kind_and_occ & 0xff, warps = (kind_and_occ >> 8) & 0xff;
    if (!c || !rate_out || iters < 1 || kind < 0 || kind > 4) return fail(GKRB200_ERR_ARG, "bad argument");  // 4: kind 1 through the out-of-line multiplier (call ABI cost)
    CUDA_TRY(cudaSetDevice(c->device));
    int block = 256, grid = c->n_sm * 8;
    size_t smem = 0;
    if (warps && (kind == 1 || kind >= 3)) {
        block = 128;                                  // 4 warps per block
        const int blocks_per_sm = warps / 4 > 0 ? warps / 4 : 1;
        smem = (size_t)(220 * 1024) / blocks_per_sm;  // only blocks_per_sm blocks fit in shared memory
        if (smem > 200 * 1024) smem = 200 * 1024;
        grid = c->n_sm * blocks_per_sm * 2;
        CUDA_TRY(cudaFuncSetAttribute(gkr::k_bench_fr_mul<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(gkr::k_bench_fr_mul<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(gkr::k_bench_fr_mul1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    void* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)grid * block * 32));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {  // first rep is warm-up
        cudaEventRecord(e0, c->stream);
        if (kind == 0) gkr::k_bench_imad_wide<<<grid, block, 0, c->stream>>>((uint64_t*)d, iters, 12345u + rep);
        else if (kind == 1) gkr::k_bench_fr_mul<0><<<grid, block, smem, c->stream>>>((FrRaw*)d, iters, 12345u + rep);
        else if (kind == 4) gkr::k_bench_fr_mul<1><<<grid, block, smem, c->stream>>>((FrRaw*)d, iters, 12345u + rep);
        else if (kind == 2) gkr::k_bench_imad_wide_x<<<grid, block, 0, c->stream>>>((uint64_t*)d, iters, 12345u + rep);
        else gkr::k_bench_fr_mul1<<<grid, block, smem, c->stream>>>((FrRaw*)d, iters, 12345u + rep);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return fail(GKRB200_ERR_CUDA, "microbench: %s", cudaGetErrorString(e));
    const double per_thread_iter = kind == 0 ? 64.0 : (kind == 2 ? 64.0 : 2.0);
    *rate_out = (double)grid * block * (double)iters * per_thread_iter / (best * 1e-3) / 1e9;
    if (ms_out) *ms_out = best;
    c->st.launches_total += 4;
    c->st.launches[KC_MISC] += 4;
    return 0;
}
