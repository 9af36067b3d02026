// libgkrb200ec.so: host driver and C ABI of the Groth16-side operations (include/gkrb200_ec.h): the G1 multi-exponentiation and
// the FFTs of computeH.  The kernels are the per-index bodies of msm.cuh and ntt.cuh run as grids of 128-thread blocks; this file
// adds the CUDA executor, the context (stream, grow-only workspaces, resident bases, the resident fft.Domain), and the host-side
// tail of InitialRandomnessHint.Call (prover/gadget/hints.go:147-192: RawBytes, legacy Keccak-256, fr.SetBytes).
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../../include/gkrb200_ec.h"
#include "msm.cuh"
#include "groth16.hpp"
#include "ntt.cuh"

namespace {

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU_TRY(x)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (x);                                                                                  \
        if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? GKRB200EC_ERR_OOM : GKRB200EC_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); \
    } while (0)

// The executor: every kernel body K::run(index, args...) of msm.cuh / ntt.cuh as a grid of 128-thread blocks on the context's stream.
// (Test seam: tests/emu/ec_hostbuild.cpp compiles THIS FILE with g++ against a stand-in for the CUDA runtime and supplies its own
// executor, so that the driver below -- staging, workspaces, slots, error paths, the Groth16 sequencing -- is exercised on the CPU
// through the same C ABI.  The shipped library is built by nvcc and contains only the executor below.)
#ifndef GKRB200EC_TEST_EXECUTOR
template <class K, class... A>
__global__ void __launch_bounds__(128) k_each(size_t n, A... a) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) K::run(i, a...);
}

struct CudaExec {
    cudaStream_t st;
    cudaError_t err = cudaSuccess;
    template <class K, class... A>
    int launch(size_t n, A... a) {
        if (n == 0) return 0;
        const unsigned blocks = (unsigned)((n + 127) / 128);
        k_each<K, A...><<<blocks, 128, 0, st>>>(n, a...);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        return 1;
    }
    void zero(void* p, size_t bytes) {
        const cudaError_t e = cudaMemsetAsync(p, 0, bytes, st);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
    }
};
#endif

}  // namespace

struct gkrb200ec_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    unsigned char* ws = nullptr;
    size_t ws_bytes = 0;
    uint64_t* d_scalars = nullptr;  // staging for host scalars
    size_t sc_cap = 0;
    uint64_t* d_tmp_points = nullptr;  // staging for one-shot host points
    size_t tp_cap = 0;
    uint64_t* d_small = nullptr;  // 2 input points + the result (Montgomery, regular) of gkrb200ec_g1_add / g2_add: 4 x 16 words at most
    uint64_t* h_pin = nullptr;    // pinned: up to 32 result words; the error flag at word PIN_FLAG
    struct Slot {
        uint64_t* d = nullptr;
        size_t n = 0;
        int words = 0;  // u64 words per point: 8 = G1Affine, 16 = G2Affine
    } slots[GKRB200EC_MAX_SLOTS];
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int c_force = 0, T_force = 0;
    // fft.Domain resident on the device: twiddle tables + split coset tables in one buffer; three work arrays for computeH
    bool have_domain = false;
    ec::NttDomainDev dom{};
    unsigned char* dom_buf = nullptr;
    uint64_t* d_abc = nullptr;  // 3 x cardinality elements
    size_t abc_cap = 0;
    gkrb200ec_stats st{};
};

namespace {

constexpr int PIN_FLAG = 48;  // word index of the error flag in h_pin (64 words)

int ensure(gkrb200ec_ctx* c, void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return 0;
    if (*p) CU_TRY(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    CU_TRY(cudaMalloc(p, bytes));
    *cap = bytes;
    (void)c;
    return 0;
}

// Self-check of every point the device hands back (host side, three field products): a result that does not satisfy the curve
// equation means a base point off the curve or a device fault, and is reported instead of returned.  (A wrong point ON the curve
// cannot be detected this way; the parity tests are what rules that out.)
template <class C>
typename C::El curve_b();
template <>
ec::Big8 curve_b<ec::G1>() {  // y^2 = x^3 + 3
    ec::Big8 t = ec::big_zero();
    t.v[0] = 3;
    return ec::f_to_mont<ec::Fp>(t);
}
template <>
ec::Fp2El curve_b<ec::G2>() {  // y^2 = x^3 + 3 / (9 + u)
    ec::Big8 nine = ec::big_zero(), one = ec::big_zero(), three = ec::big_zero();
    nine.v[0] = 9, one.v[0] = 1, three.v[0] = 3;
    ec::Fp2El d, n;
    d.c0 = ec::f_to_mont<ec::Fp>(nine), d.c1 = ec::f_to_mont<ec::Fp>(one);
    n.c0 = ec::f_to_mont<ec::Fp>(three), n.c1 = ec::big_zero();
    return ec::Fp2Mul<ec::FpMulCall>::mul(n, ec::Fp2Base::inv(d));
}
template <class C>
bool on_curve(const uint64_t* aff_mont) {
    typedef typename C::Field F;
    typedef typename C::MCall M;
    const typename C::Affine a = C::aff_load(aff_mont);
    if (C::aff_is_inf(a)) return true;
    static const typename C::El b = curve_b<C>();
    return F::is_zero(F::sub(M::sqr(a.y), F::add(M::mul(M::sqr(a.x), a.x), b)));
}

// out: affine result, Montgomery (C::AFF_WORDS words) then regular form (C::AFF_WORDS words)
template <class C>
int run_msm(gkrb200ec_ctx* c, const uint64_t* d_points, const uint64_t* d_scalars, size_t n, int form, uint64_t* out) {
    constexpr size_t RES = 2 * C::AFF_WORDS;
    if (n == 0) {
        memset(out, 0, RES * sizeof(uint64_t));
        return 0;
    }
    const ec::MsmPlan pl = ec::msm_make_plan(n, form == GKRB200EC_SCALARS_MONTGOMERY, c->c_force, c->T_force);
    if ((uint64_t)pl.n * pl.W >= 0xffffffffull) return fail(GKRB200EC_ERR_ARG, "multiexp: %zu points x %u windows does not fit 32-bit entry offsets", n, pl.W);
    const ec::MsmWorkspace ws = ec::msm_layout(pl, 8 * C::X_WORDS, 8 * C::AFF_WORDS);
    {
        void* p = c->ws;
        const int rc = ensure(c, &p, &c->ws_bytes, ws.bytes);
        c->ws = (unsigned char*)p;
        if (rc) return rc;
    }
    CudaExec ex{c->stream};
    CU_TRY(cudaEventRecord(c->e0, c->stream));
    const int launches = ec::msm_enqueue<C>(ex, pl, ws, c->ws, d_points, d_scalars);
    CU_TRY(cudaEventRecord(c->e1, c->stream));
    if (ex.err != cudaSuccess) return fail(GKRB200EC_ERR_CUDA, "multiexp launch: %s", cudaGetErrorString(ex.err));
    CU_TRY(cudaMemcpyAsync(c->h_pin, c->ws + ws.out, RES * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(c->h_pin + PIN_FLAG, c->ws + ws.err, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, c->e0, c->e1));
    c->st.launches_total += (uint64_t)launches;
    c->st.msm_calls++;
    c->st.last_n = pl.n, c->st.last_c = pl.c, c->st.last_windows = pl.W, c->st.last_task_size = pl.T;
    c->st.last_tasks_max = pl.max_tasks;
    c->st.workspace_bytes = c->ws_bytes;
    c->st.d2h_bytes += RES * sizeof(uint64_t) + sizeof(uint32_t);
    c->st.last_device_ms = ms;
    const uint32_t flag = (uint32_t)c->h_pin[PIN_FLAG];
    if (flag & ec::MSM_ERR_SCALAR_RANGE) return fail(GKRB200EC_ERR_ARG, "multiexp: a regular-form scalar is not reduced (>= q)");
    if (flag) return fail(GKRB200EC_ERR_CUDA, "multiexp: internal digit overflow (flag %u)", flag);
    memcpy(out, c->h_pin, RES * sizeof(uint64_t));
    if (!on_curve<C>(out)) return fail(GKRB200EC_ERR_CUDA, "multiexp: the result is not on the curve (a base point off the curve, or a device fault)");
    return 0;
}

int stage_scalars(gkrb200ec_ctx* c, const uint64_t* scalars, size_t n) {
    void* p = c->d_scalars;
    const int rc = ensure(c, &p, &c->sc_cap, n * 32);
    c->d_scalars = (uint64_t*)p;
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(c->d_scalars, scalars, n * 32, cudaMemcpyHostToDevice, c->stream));
    c->st.h2d_bytes += n * 32;
    return 0;
}

int check_slot(gkrb200ec_ctx* c, int slot, size_t n, int words) {
    if (!c) return fail(GKRB200EC_ERR_ARG, "null context");
    if (slot < 0 || slot >= GKRB200EC_MAX_SLOTS) return fail(GKRB200EC_ERR_ARG, "base slot %d out of range", slot);
    if (n > c->slots[slot].n) return fail(GKRB200EC_ERR_ARG, "%zu scalars for %zu bases in slot %d", n, c->slots[slot].n, slot);
    if (n && c->slots[slot].words != words) return fail(GKRB200EC_ERR_ARG, "slot %d holds %s points", slot, c->slots[slot].words == 8 ? "G1" : "G2");
    return 0;
}

// out: a + b, Montgomery then regular form (2 * C::AFF_WORDS words)
template <class C>
int add_points(gkrb200ec_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    constexpr size_t AW = C::AFF_WORDS;
    CU_TRY(cudaMemcpyAsync(c->d_small, a, AW * 8, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(c->d_small + AW, b, AW * 8, cudaMemcpyHostToDevice, c->stream));
    CudaExec ex{c->stream};
    c->st.launches_total += (uint64_t)ex.launch<ec::KAddAffine<C>>(1, (const uint64_t*)c->d_small, (const uint64_t*)(c->d_small + AW), c->d_small + 2 * AW);
    if (ex.err != cudaSuccess) return fail(GKRB200EC_ERR_CUDA, "point addition launch: %s", cudaGetErrorString(ex.err));
    CU_TRY(cudaMemcpyAsync(c->h_pin, c->d_small + 2 * AW, 2 * AW * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out, c->h_pin, 2 * AW * 8);
    c->st.h2d_bytes += 2 * AW * 8, c->st.d2h_bytes += 2 * AW * 8;
    if (!on_curve<C>(out)) return fail(GKRB200EC_ERR_CUDA, "point addition: the result is not on the curve (an operand off the curve, or a device fault)");
    return 0;
}

int set_bases(gkrb200ec_ctx* c, int slot, const uint64_t* points, size_t n, int words) {
    if (const int rc = check_slot(c, slot, 0, words)) return rc;
    if (n && !points) return fail(GKRB200EC_ERR_ARG, "null points");
    if (n > GKRB200EC_MAX_POINTS) return fail(GKRB200EC_ERR_ARG, "%zu points: at most %u", n, GKRB200EC_MAX_POINTS);
    CU_TRY(cudaSetDevice(c->device));
    auto& s = c->slots[slot];
    if (s.d) {
        CU_TRY(cudaStreamSynchronize(c->stream));
        CU_TRY(cudaFree(s.d));
        s.d = nullptr, s.n = 0, s.words = 0;
    }
    if (n == 0) return 0;
    CU_TRY(cudaMalloc((void**)&s.d, n * words * 8));
    CU_TRY(cudaMemcpyAsync(s.d, points, n * words * 8, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    s.n = n, s.words = words;
    c->st.h2d_bytes += n * words * 8;
    return 0;
}

// the three host-facing forms of MultiExp, once per curve
template <class C>
int multiexp_device(gkrb200ec_ctx* c, int slot, const void* d_scalars, size_t n, int form, uint64_t* out) {
    if (const int rc = check_slot(c, slot, n, C::AFF_WORDS)) return rc;
    if (!out || (n && !d_scalars)) return fail(GKRB200EC_ERR_ARG, "null pointer");
    if (form != GKRB200EC_SCALARS_REGULAR && form != GKRB200EC_SCALARS_MONTGOMERY) return fail(GKRB200EC_ERR_ARG, "scalar form %d", form);
    CU_TRY(cudaSetDevice(c->device));
    uint64_t r[2 * C::AFF_WORDS];
    if (const int rc = run_msm<C>(c, c->slots[slot].d, (const uint64_t*)d_scalars, n, form, r)) return rc;
    memcpy(out, r, C::AFF_WORDS * 8);
    return 0;
}
template <class C>
int multiexp_host(gkrb200ec_ctx* c, int slot, const uint64_t* scalars, size_t n, int form, uint64_t* out) {
    if (const int rc = check_slot(c, slot, n, C::AFF_WORDS)) return rc;
    if (!out || (n && !scalars)) return fail(GKRB200EC_ERR_ARG, "null pointer");
    CU_TRY(cudaSetDevice(c->device));
    if (n)
        if (const int rc = stage_scalars(c, scalars, n)) return rc;
    return multiexp_device<C>(c, slot, c->d_scalars, n, form, out);
}
template <class C>
int multiexp_points(gkrb200ec_ctx* c, const uint64_t* points, const uint64_t* scalars, size_t n, int form, uint64_t* out) {
    if (!c) return fail(GKRB200EC_ERR_ARG, "null context");
    if (!out || (n && (!points || !scalars))) return fail(GKRB200EC_ERR_ARG, "null pointer");
    if (n > GKRB200EC_MAX_POINTS) return fail(GKRB200EC_ERR_ARG, "%zu points: at most %u", n, GKRB200EC_MAX_POINTS);
    if (form != GKRB200EC_SCALARS_REGULAR && form != GKRB200EC_SCALARS_MONTGOMERY) return fail(GKRB200EC_ERR_ARG, "scalar form %d", form);
    CU_TRY(cudaSetDevice(c->device));
    uint64_t r[2 * C::AFF_WORDS];
    if (n) {
        void* p = c->d_tmp_points;
        const int rc = ensure(c, &p, &c->tp_cap, n * C::AFF_WORDS * 8);
        c->d_tmp_points = (uint64_t*)p;
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(c->d_tmp_points, points, n * C::AFF_WORDS * 8, cudaMemcpyHostToDevice, c->stream));
        c->st.h2d_bytes += n * C::AFF_WORDS * 8;
        if (const int rc2 = stage_scalars(c, scalars, n)) return rc2;
    }
    if (const int rc = run_msm<C>(c, c->d_tmp_points, c->d_scalars, n, form, r)) return rc;
    memcpy(out, r, C::AFF_WORDS * 8);
    return 0;
}

// ---- host side of DeriveRandomnessFromPoint -----------------------------------------------------------------------------------
// legacy Keccak-256: Keccak-f[1600], rate 136 bytes, multi-rate padding with domain byte 0x01 (not SHA-3's 0x06)
void keccak_f1600(uint64_t* a) {
    static const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
                                    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
                                    0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
                                    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                                    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    // rho offsets and pi destinations along the 24-lane cycle starting at lane 1
    static const int RHO[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PI[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int x = 0; x < 5; x++) bc[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) {
            const uint64_t r = bc[(x + 1) % 5];
            const uint64_t t = bc[(x + 4) % 5] ^ ((r << 1) | (r >> 63));
            for (int y = 0; y < 25; y += 5) a[y + x] ^= t;
        }
        uint64_t t = a[1];
        for (int i = 0; i < 24; i++) {
            const int j = PI[i];
            const uint64_t keep = a[j];
            a[j] = (t << RHO[i]) | (t >> (64 - RHO[i]));
            t = keep;
        }
        for (int y = 0; y < 25; y += 5) {
            for (int x = 0; x < 5; x++) bc[x] = a[y + x];
            for (int x = 0; x < 5; x++) a[y + x] = bc[x] ^ (~bc[(x + 1) % 5] & bc[(x + 2) % 5]);
        }
        a[0] ^= RC[round];
    }
}
void keccak256(const uint8_t* data, size_t len, uint8_t* out) {
    uint64_t a[25];
    memset(a, 0, sizeof a);
    const size_t rate = 136;
    uint8_t block[136];
    size_t done = 0;
    bool last = false;
    while (!last) {
        const size_t take = len - done < rate ? len - done : rate;
        last = take < rate;
        memset(block, 0, rate);
        if (take) memcpy(block, data + done, take);
        if (last) {
            block[take] ^= 0x01;
            block[rate - 1] ^= 0x80;
        }
        for (size_t i = 0; i < rate / 8; i++) {
            uint64_t w;
            memcpy(&w, block + 8 * i, 8);  // little-endian host (x86-64 / aarch64)
            a[i] ^= w;
        }
        keccak_f1600(a);
        done += take;
    }
    memcpy(out, a, 32);
}
// regular-form affine words (X: r[0..4), Y: r[4..8)) -> RawBytes
void raw_bytes_from_regular(const uint64_t* r, uint8_t* out) {
    bool inf = true;
    for (int i = 0; i < 8; i++) inf = inf && r[i] == 0;
    memset(out, 0, 64);
    if (inf) {
        out[0] = 0x40;  // mUncompressedInfinity
        return;
    }
    for (int i = 0; i < 32; i++) {
        out[31 - i] = (uint8_t)(r[i / 8] >> (8 * (i % 8)));
        out[63 - i] = (uint8_t)(r[4 + i / 8] >> (8 * (i % 8)));
    }
}
// fr.SetBytes of a 32-byte big-endian value, returned in regular form: the integer mod q
void fr_set_bytes_regular(const uint8_t* be, uint64_t* out) {
    ec::Big8 v;
    for (int i = 0; i < 8; i++) v.v[i] = (uint32_t)be[31 - 4 * i] | ((uint32_t)be[30 - 4 * i] << 8) | ((uint32_t)be[29 - 4 * i] << 16) | ((uint32_t)be[28 - 4 * i] << 24);
    for (int k = 0; k < 6 && !ec::f_is_canonical<ec::FrMod>(v); k++) {  // 2^256 / q < 6
        ec::Big8 s;
        (void)ec::f_sub_mod<ec::FrMod>(s, v);
        v = s;
    }
    ec::big_store(out, v);
}
void derive_from_regular(const uint64_t* reg8, uint64_t* out) {
    uint8_t raw[64], h[32];
    raw_bytes_from_regular(reg8, raw);
    keccak256(raw, 64, h);
    fr_set_bytes_regular(h, out);
}
void regular_from_mont(const uint64_t* g1, uint64_t* reg8) {
    const ec::Big8 x = ec::f_from_mont<ec::Fp>(ec::big_load(g1)), y = ec::f_from_mont<ec::Fp>(ec::big_load(g1 + 4));
    ec::big_store(reg8, x);
    ec::big_store(reg8 + 4, y);
}


// ---- fft.Domain on the device -------------------------------------------------------------------------------------------------
int domain_upload(gkrb200ec_ctx* c, uint32_t log_n) {
    ec::NttDomainHost h;
    if (!ec::ntt_domain_host(log_n, h)) return fail(GKRB200EC_ERR_ARG, "fft domain of 2^%u elements: at most 2^%d", log_n, ec::NTT_MAX_LOG);
    const size_t n = (size_t)1 << log_n, half = n > 1 ? n / 2 : 1;
    struct Piece {
        uint64_t** dst;
        const uint64_t* src;  // null: filled on the device
        size_t words;
    };
    ec::NttDomainDev d{};
    d.log_n = log_n;
    const Piece pieces[] = {
        {&d.tw, nullptr, 4 * half},
        {&d.tw_inv, nullptr, 4 * half},
        {&d.w_lo, h.w_lo.data(), h.w_lo.size()},
        {&d.w_hi, h.w_hi.data(), h.w_hi.size()},
        {&d.wi_lo, h.wi_lo.data(), h.wi_lo.size()},
        {&d.wi_hi, h.wi_hi.data(), h.wi_hi.size()},
        {&d.u_lo, h.u_lo.data(), h.u_lo.size()},
        {&d.u_hi, h.u_hi.data(), h.u_hi.size()},
        {&d.u_hi_n, h.u_hi_n.data(), h.u_hi_n.size()},
        {&d.ui_lo, h.ui_lo.data(), h.ui_lo.size()},
        {&d.ui_hi_n, h.ui_hi_n.data(), h.ui_hi_n.size()},
        {&d.n_inv, h.n_inv, 4},
        {&d.minus_two_inv, h.minus_two_inv, 4},
    };
    size_t total = 0;
    for (const Piece& p : pieces) total += ec::msm_align(p.words * 8);
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (c->dom_buf) CU_TRY(cudaFree(c->dom_buf));
    c->dom_buf = nullptr;
    c->have_domain = false;
    CU_TRY(cudaMalloc((void**)&c->dom_buf, total));
    size_t off = 0;
    for (const Piece& p : pieces) {
        *p.dst = (uint64_t*)(c->dom_buf + off);
        if (p.src) {
            CU_TRY(cudaMemcpyAsync(*p.dst, p.src, p.words * 8, cudaMemcpyHostToDevice, c->stream));
            c->st.h2d_bytes += p.words * 8;
        }
        off += ec::msm_align(p.words * 8);
    }
    CudaExec ex{c->stream};
    c->st.launches_total += (uint64_t)ec::ntt_domain_enqueue(ex, d);
    if (ex.err != cudaSuccess) return fail(GKRB200EC_ERR_CUDA, "fft domain launch: %s", cudaGetErrorString(ex.err));
    CU_TRY(cudaStreamSynchronize(c->stream));  // the host tables die with this frame
    c->dom = d;
    c->have_domain = true;
    return 0;
}
int check_domain(gkrb200ec_ctx* c, size_t n) {
    if (!c) return fail(GKRB200EC_ERR_ARG, "null context");
    if (!c->have_domain) return fail(GKRB200EC_ERR_ARG, "no fft domain: call gkrb200ec_fft_domain_init first");
    if (n != (size_t)1 << c->dom.log_n) return fail(GKRB200EC_ERR_ARG, "%zu elements for a domain of cardinality %zu", n, (size_t)1 << c->dom.log_n);
    return 0;
}
int ensure_abc(gkrb200ec_ctx* c) {
    void* p = c->d_abc;
    const int rc = ensure(c, &p, &c->abc_cap, 3 * ((size_t)32 << c->dom.log_n));
    c->d_abc = (uint64_t*)p;
    return rc;
}
int run_fft(gkrb200ec_ctx* c, uint64_t* a, size_t n, int decimation, int coset, bool inverse) {
    if (const int rc = check_domain(c, n)) return rc;
    if (!a) return fail(GKRB200EC_ERR_ARG, "null pointer");
    if (decimation != GKRB200EC_DIT && decimation != GKRB200EC_DIF) return fail(GKRB200EC_ERR_ARG, "decimation %d", decimation);
    if (coset != 0 && coset != 1) return fail(GKRB200EC_ERR_ARG, "coset %d: the domain has depth 1 (cosets 0 and 1)", coset);
    CU_TRY(cudaSetDevice(c->device));
    if (const int rc = ensure_abc(c)) return rc;
    CU_TRY(cudaMemcpyAsync(c->d_abc, a, n * 32, cudaMemcpyHostToDevice, c->stream));
    CudaExec ex{c->stream};
    CU_TRY(cudaEventRecord(c->e0, c->stream));
    const int launches = ec::fft_enqueue(ex, c->dom, c->d_abc, decimation == GKRB200EC_DIF, coset, inverse);
    CU_TRY(cudaEventRecord(c->e1, c->stream));
    if (ex.err != cudaSuccess) return fail(GKRB200EC_ERR_CUDA, "fft launch: %s", cudaGetErrorString(ex.err));
    CU_TRY(cudaMemcpyAsync(a, c->d_abc, n * 32, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, c->e0, c->e1));
    c->st.launches_total += (uint64_t)launches;
    c->st.fft_calls++;
    c->st.last_fft_device_ms = ms;
    c->st.h2d_bytes += n * 32, c->st.d2h_bytes += n * 32;
    return 0;
}

}  // namespace

extern "C" {

const char* gkrb200ec_version(void) { return "gkrb200ec 0.2 (sm_100a; G1 bucket method in XYZZ with signed digits; radix-8 in-register Fr transforms)"; }
const char* gkrb200ec_last_error(void) { return g_err; }

int gkrb200ec_init(gkrb200ec_ctx** out, int device, void* stream) {
    if (!out) return fail(GKRB200EC_ERR_ARG, "null ctx pointer");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(GKRB200EC_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(GKRB200EC_ERR_ARG, "device %d of %d", device, ndev);
    CU_TRY(cudaSetDevice(device));
    gkrb200ec_ctx* c = new gkrb200ec_ctx();
    c->device = device;
    int rc = 0;
    auto step = [&](cudaError_t e, const char* what) {
        if (rc == 0 && e != cudaSuccess) rc = fail(GKRB200EC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    };
    if (stream) c->stream = (cudaStream_t)stream;
    else {
        step(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "stream");
        c->own_stream = rc == 0;
    }
    step(cudaMalloc((void**)&c->d_small, 64 * sizeof(uint64_t)), "scratch");
    step(cudaMallocHost((void**)&c->h_pin, 64 * sizeof(uint64_t)), "pinned result");
    step(cudaEventCreate(&c->e0), "event");
    step(cudaEventCreate(&c->e1), "event");
    if (rc) {
        gkrb200ec_free(c);
        return rc;
    }
    *out = c;
    return 0;
}

void gkrb200ec_free(gkrb200ec_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& s : c->slots)
        if (s.d) cudaFree(s.d);
    if (c->ws) cudaFree(c->ws);
    if (c->d_scalars) cudaFree(c->d_scalars);
    if (c->d_tmp_points) cudaFree(c->d_tmp_points);
    if (c->d_small) cudaFree(c->d_small);
    if (c->dom_buf) cudaFree(c->dom_buf);
    if (c->d_abc) cudaFree(c->d_abc);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->e0) cudaEventDestroy(c->e0);
    if (c->e1) cudaEventDestroy(c->e1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int gkrb200ec_g1_set_bases(gkrb200ec_ctx* c, int slot, const uint64_t* points, size_t n) { return set_bases(c, slot, points, n, ec::G1::AFF_WORDS); }
int gkrb200ec_g2_set_bases(gkrb200ec_ctx* c, int slot, const uint64_t* points, size_t n) { return set_bases(c, slot, points, n, ec::G2::AFF_WORDS); }

int gkrb200ec_g1_multiexp_device(gkrb200ec_ctx* c, int slot, const void* d_scalars, size_t n, int form, uint64_t* out) {
    return multiexp_device<ec::G1>(c, slot, d_scalars, n, form, out);
}
int gkrb200ec_g1_multiexp(gkrb200ec_ctx* c, int slot, const uint64_t* scalars, size_t n, int form, uint64_t* out) {
    return multiexp_host<ec::G1>(c, slot, scalars, n, form, out);
}
int gkrb200ec_g1_multiexp_points(gkrb200ec_ctx* c, const uint64_t* points, const uint64_t* scalars, size_t n, int form, uint64_t* out) {
    return multiexp_points<ec::G1>(c, points, scalars, n, form, out);
}
int gkrb200ec_g2_multiexp_device(gkrb200ec_ctx* c, int slot, const void* d_scalars, size_t n, int form, uint64_t* out) {
    return multiexp_device<ec::G2>(c, slot, d_scalars, n, form, out);
}
int gkrb200ec_g2_multiexp(gkrb200ec_ctx* c, int slot, const uint64_t* scalars, size_t n, int form, uint64_t* out) {
    return multiexp_host<ec::G2>(c, slot, scalars, n, form, out);
}
int gkrb200ec_g2_multiexp_points(gkrb200ec_ctx* c, const uint64_t* points, const uint64_t* scalars, size_t n, int form, uint64_t* out) {
    return multiexp_points<ec::G2>(c, points, scalars, n, form, out);
}

int gkrb200ec_g1_add(gkrb200ec_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    if (!c || !a || !b || !out) return fail(GKRB200EC_ERR_ARG, "null pointer");
    CU_TRY(cudaSetDevice(c->device));
    uint64_t r[2 * ec::G1::AFF_WORDS];
    if (const int rc = add_points<ec::G1>(c, a, b, r)) return rc;
    memcpy(out, r, ec::G1::AFF_WORDS * 8);
    return 0;
}
int gkrb200ec_g2_add(gkrb200ec_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    if (!c || !a || !b || !out) return fail(GKRB200EC_ERR_ARG, "null pointer");
    CU_TRY(cudaSetDevice(c->device));
    uint64_t r[2 * ec::G2::AFF_WORDS];
    if (const int rc = add_points<ec::G2>(c, a, b, r)) return rc;
    memcpy(out, r, ec::G2::AFF_WORDS * 8);
    return 0;
}

int gkrb200ec_initial_randomness(gkrb200ec_ctx* c, int slot_pub, const uint64_t* scalars_pub, size_t n_pub, int slot_priv, const uint64_t* scalars_priv,
                                 size_t n_priv, int form, uint64_t* krs_gkr_priv_out, uint64_t* initial_randomness_out) {
    if (!krs_gkr_priv_out || !initial_randomness_out) return fail(GKRB200EC_ERR_ARG, "null output");
    uint64_t krs[8], priv[8], sum[16];
    if (const int rc = gkrb200ec_g1_multiexp(c, slot_pub, scalars_pub, n_pub, form, krs)) return rc;        // hints.go:182
    if (const int rc = gkrb200ec_g1_multiexp(c, slot_priv, scalars_priv, n_priv, form, priv)) return rc;     // hints.go:183
    if (const int rc = add_points<ec::G1>(c, krs, priv, sum)) return rc;                                             // hints.go:184
    memcpy(krs_gkr_priv_out, priv, 64);                                                                      // hints.go:186
    derive_from_regular(sum + 8, initial_randomness_out);                                                    // hints.go:188-189
    return 0;
}

int gkrb200ec_g1_raw_bytes(const uint64_t* g1, uint8_t out[64]) {
    if (!g1 || !out) return fail(GKRB200EC_ERR_ARG, "null pointer");
    uint64_t reg[8];
    regular_from_mont(g1, reg);
    raw_bytes_from_regular(reg, out);
    return 0;
}
int gkrb200ec_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    if ((len && !data) || !out) return fail(GKRB200EC_ERR_ARG, "null pointer");
    keccak256(data, len, out);
    return 0;
}
int gkrb200ec_derive_randomness_from_point(const uint64_t* g1, uint64_t* out) {
    if (!g1 || !out) return fail(GKRB200EC_ERR_ARG, "null pointer");
    uint64_t reg[8];
    regular_from_mont(g1, reg);
    derive_from_regular(reg, out);
    return 0;
}

// ---- fft.Domain, FFT, FFTInverse, computeH ------------------------------------------------------------------------------------
int gkrb200ec_fft_domain_init(gkrb200ec_ctx* c, uint64_t m) {
    if (!c) return fail(GKRB200EC_ERR_ARG, "null context");
    if (m == 0 || m > ((uint64_t)1 << ec::NTT_MAX_LOG)) return fail(GKRB200EC_ERR_ARG, "fft domain for %llu elements: 1 .. 2^%d", (unsigned long long)m, ec::NTT_MAX_LOG);
    uint32_t log_n = 0;
    while (((uint64_t)1 << log_n) < m) log_n++;  // ecc.NextPowerOfTwo
    CU_TRY(cudaSetDevice(c->device));
    return domain_upload(c, log_n);
}
uint64_t gkrb200ec_fft_domain_cardinality(gkrb200ec_ctx* c) { return (c && c->have_domain) ? (uint64_t)1 << c->dom.log_n : 0; }
int gkrb200ec_fft(gkrb200ec_ctx* c, uint64_t* a, size_t n, int decimation, int coset) { return run_fft(c, a, n, decimation, coset, false); }
int gkrb200ec_fft_inverse(gkrb200ec_ctx* c, uint64_t* a, size_t n, int decimation, int coset) { return run_fft(c, a, n, decimation, coset, true); }

int gkrb200ec_compute_h(gkrb200ec_ctx* c, const uint64_t* a, const uint64_t* b, const uint64_t* cc, size_t n_in, uint64_t* h_out, const void** d_h_out) {
    if (!c) return fail(GKRB200EC_ERR_ARG, "null context");
    if (!c->have_domain) return fail(GKRB200EC_ERR_ARG, "no fft domain: call gkrb200ec_fft_domain_init first");
    const size_t n = (size_t)1 << c->dom.log_n;
    if (n_in > n) return fail(GKRB200EC_ERR_ARG, "%zu constraints for a domain of cardinality %zu", n_in, n);
    if (n_in && (!a || !b || !cc)) return fail(GKRB200EC_ERR_ARG, "null pointer");
    if (!h_out && !d_h_out) return fail(GKRB200EC_ERR_ARG, "no output requested");
    CU_TRY(cudaSetDevice(c->device));
    if (const int rc = ensure_abc(c)) return rc;
    uint64_t* v[3] = {c->d_abc, c->d_abc + 4 * n, c->d_abc + 8 * n};
    const uint64_t* src[3] = {a, b, cc};
    for (int t = 0; t < 3; t++) {  // append(a, padding...) (prove.go:319-323)
        if (n_in) CU_TRY(cudaMemcpyAsync(v[t], src[t], n_in * 32, cudaMemcpyHostToDevice, c->stream));
        if (n_in < n) CU_TRY(cudaMemsetAsync(v[t] + 4 * n_in, 0, (n - n_in) * 32, c->stream));
    }
    CudaExec ex{c->stream};
    CU_TRY(cudaEventRecord(c->e0, c->stream));
    const int launches = ec::compute_h_enqueue(ex, c->dom, v[0], v[1], v[2]);
    CU_TRY(cudaEventRecord(c->e1, c->stream));
    if (ex.err != cudaSuccess) return fail(GKRB200EC_ERR_CUDA, "computeH launch: %s", cudaGetErrorString(ex.err));
    if (h_out) CU_TRY(cudaMemcpyAsync(h_out, v[0], n * 32, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, c->e0, c->e1));
    c->st.launches_total += (uint64_t)launches;
    c->st.fft_calls++;
    c->st.last_fft_device_ms = ms;
    c->st.h2d_bytes += 3 * n_in * 32;
    if (h_out) c->st.d2h_bytes += n * 32;
    if (d_h_out) *d_h_out = v[0];
    return 0;
}

// ---- ComputeGroth16Proof ------------------------------------------------------------------------------------------------------
namespace {
struct CudaGroth16Ops {
    gkrb200ec_ctx* c;
    const gkrb200ec_groth16_pk* pk;
    const uint64_t *a, *b, *cc;
    size_t n_constraints;
    const uint64_t *wa, *wb;
    size_t na, nb;
    int form;
    const void* d_h = nullptr;
    const uint64_t* alpha() const { return pk->g1_alpha; }
    const uint64_t* beta() const { return pk->g1_beta; }
    const uint64_t* delta() const { return pk->g1_delta; }
    const uint64_t* beta2() const { return pk->g2_beta; }
    const uint64_t* delta2() const { return pk->g2_delta; }
    int smul_g1(const uint64_t* pt, const uint64_t* k_regular, uint64_t* out) { return multiexp_points<ec::G1>(c, pt, k_regular, 1, GKRB200EC_SCALARS_REGULAR, out); }
    int smul_g2(const uint64_t* pt, const uint64_t* k_regular, uint64_t* out) { return multiexp_points<ec::G2>(c, pt, k_regular, 1, GKRB200EC_SCALARS_REGULAR, out); }
    int add_g1(const uint64_t* x, const uint64_t* y, uint64_t* out) {
        uint64_t r[2 * ec::G1::AFF_WORDS];
        if (const int rc = add_points<ec::G1>(c, x, y, r)) return rc;
        memcpy(out, r, ec::G1::AFF_WORDS * 8);
        return 0;
    }
    int add_g2(const uint64_t* x, const uint64_t* y, uint64_t* out) {
        uint64_t r[2 * ec::G2::AFF_WORDS];
        if (const int rc = add_points<ec::G2>(c, x, y, r)) return rc;
        memcpy(out, r, ec::G2::AFF_WORDS * 8);
        return 0;
    }
    int compute_h() { return gkrb200ec_compute_h(c, a, b, cc, n_constraints, nullptr, &d_h); }
    int msm_g1(int which, uint64_t* out) {  // 0: pk.G1.A x wireValuesA, 1: pk.G1.B x wireValuesB
        return which == 0 ? multiexp_host<ec::G1>(c, pk->slot_g1_a, wa, na, form, out) : multiexp_host<ec::G1>(c, pk->slot_g1_b, wb, nb, form, out);
    }
    int msm_g1_h(uint64_t* out) {  // pk.G1.Z x h, h read where computeH left it (regular form)
        return multiexp_device<ec::G1>(c, pk->slot_g1_z, d_h, (size_t)1 << c->dom.log_n, GKRB200EC_SCALARS_REGULAR, out);
    }
    int msm_g2(uint64_t* out) { return multiexp_host<ec::G2>(c, pk->slot_g2_b, wb, nb, form, out); }
};
}  // namespace

int gkrb200ec_groth16_prove(gkrb200ec_ctx* c, const gkrb200ec_groth16_pk* pk, const uint64_t* a, const uint64_t* b, const uint64_t* cc, size_t n_constraints,
                            const uint64_t* wa, size_t na, const uint64_t* wb, size_t nb, int form, const uint64_t* r, const uint64_t* s, uint64_t* ar_out,
                            uint64_t* bs_out, uint64_t* krs_out) {
    if (!c || !pk || !r || !s || !ar_out || !bs_out || !krs_out) return fail(GKRB200EC_ERR_ARG, "null pointer");
    if (!pk->g1_alpha || !pk->g1_beta || !pk->g1_delta || !pk->g2_beta || !pk->g2_delta) return fail(GKRB200EC_ERR_ARG, "null proving-key point");
    if (!c->have_domain) return fail(GKRB200EC_ERR_ARG, "no fft domain: call gkrb200ec_fft_domain_init first");
    if (const int rc = check_slot(c, pk->slot_g1_z, (size_t)1 << c->dom.log_n, ec::G1::AFF_WORDS)) return rc;  // len(pk.G1.Z) == domain.Cardinality
    if (!ec::f_is_canonical<ec::FrMod>(ec::big_load(r)) || !ec::f_is_canonical<ec::FrMod>(ec::big_load(s))) return fail(GKRB200EC_ERR_ARG, "r or s is not a reduced fr.Element");
    CudaGroth16Ops ops{c, pk, a, b, cc, n_constraints, wa, wb, na, nb, form};
    ec::Groth16Out out;
    if (const int rc = ec::groth16_compose(ops, r, s, out)) return rc;
    memcpy(ar_out, out.ar, sizeof out.ar);
    memcpy(bs_out, out.bs, sizeof out.bs);
    memcpy(krs_out, out.krs, sizeof out.krs);
    return 0;
}

int gkrb200ec_set_plan(gkrb200ec_ctx* c, int window_bits, int task_size) {
    if (!c) return fail(GKRB200EC_ERR_ARG, "null context");
    if (window_bits != 0 && (window_bits < 2 || window_bits > 16)) return fail(GKRB200EC_ERR_ARG, "window width %d outside 2..16", window_bits);
    if (task_size < 0) return fail(GKRB200EC_ERR_ARG, "task size %d", task_size);
    c->c_force = window_bits;
    c->T_force = task_size;
    return 0;
}
int gkrb200ec_get_stats(gkrb200ec_ctx* c, gkrb200ec_stats* out) {
    if (!c || !out) return fail(GKRB200EC_ERR_ARG, "null pointer");
    *out = c->st;
    return 0;
}

}  // extern "C"
