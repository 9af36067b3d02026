// CPU emulation of the multi-exponentiation and FFT kernels -- TEST INFRASTRUCTURE, never shipped or loaded by the product.
//
// Compiles the product headers gkr-mimc_b200/csrc/ec/{field,curve,msm,ntt}.cuh with a plain C++ compiler: every kernel body is a function of
// its thread index, and the executor below runs each "launch" as a loop.  Host-side, field.cuh's carry-chain primitives are plain
// 64-bit C++ with the semantics of the inline-PTX ones the device uses (those are exercised on the GPU by every GKR parity test).
// tests/test_msm_cpu.py and tests/test_ntt_cpu.py drive this against the oracles, so the digit decomposition, counting sort, task
// splitting, XYZZ formulas with all exceptional cases, window reduction, the in-register butterfly passes, coset scalings and the
// drivers' launch sequences are checked without a GPU; tests/test_zz2_msm_gpu.py and tests/test_zz1_ntt_gpu.py then check the real
// library on the device against the same oracles.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../gkr-mimc_b200/csrc/ec/msm.cuh"
#include "../../gkr-mimc_b200/csrc/ec/ntt.cuh"
#include "../../gkr-mimc_b200/csrc/ec/groth16.hpp"

namespace {
struct HostExec {
    template <class K, class... A>
    int launch(size_t n, A... a) {
        if (n == 0) return 0;
        for (size_t i = 0; i < n; i++) K::run(i, a...);
        return 1;
    }
    // a second order of execution: threads run from the last to the first (shakes out order dependence of the atomics)
    bool reverse = false;
    void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
};
struct HostExecReverse {
    template <class K, class... A>
    int launch(size_t n, A... a) {
        for (size_t i = n; i-- > 0;) K::run(i, a...);
        return n ? 1 : 0;
    }
    void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
};
using namespace ec;
Big8 ld(const uint64_t* p) { return big_load(p); }
}  // namespace

extern "C" {

// op: 0 mul, 1 sqr, 2 add, 3 sub, 4 inv, 5 from_mont, 6 to_mont, 7 out-of-line mul;  field: 0 = Fp, 1 = Fr
void emu_field_op(int field, int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
    for (size_t i = 0; i < n; i++) {
        const Big8 x = ld(a + 4 * i), y = b ? ld(b + 4 * i) : big_zero();
        Big8 r = big_zero();
        if (field == 0) {
            switch (op) {
                case 0: r = f_mul<FpMod>(x, y); break;
                case 1: r = f_sqr<FpMod>(x); break;
                case 2: r = f_add<FpMod>(x, y); break;
                case 3: r = f_sub<FpMod>(x, y); break;
                case 4: r = f_inv<FpMod>(x); break;
                case 5: r = f_from_mont<FpMod>(x); break;
                case 6: r = f_to_mont<FpMod>(x); break;
                case 7: r = f_mulc<FpMod>(x, y); break;
            }
        } else {
            switch (op) {
                case 0: r = f_mul<FrMod>(x, y); break;
                case 1: r = f_sqr<FrMod>(x); break;
                case 2: r = f_add<FrMod>(x, y); break;
                case 3: r = f_sub<FrMod>(x, y); break;
                case 4: r = f_inv<FrMod>(x); break;
                case 5: r = f_from_mont<FrMod>(x); break;
                case 6: r = f_to_mont<FrMod>(x); break;
                case 7: r = f_mulc<FrMod>(x, y); break;
            }
        }
        big_store(out + 4 * i, r);
    }
}

}  // extern "C"

namespace {
// XYZZ operations on affine inputs, result affine (Montgomery), on curve C.  op: 0 madd (a as XYZZ + affine b), 1 full add,
// 2 double a, 3 k * a with k = b[0] (small), 4 madd onto a non-trivial representation of a (ZZ != 1), incl. the doubling /
// cancellation cases, 5 full add of two non-trivial representations
template <class C>
void g_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    typedef typename C::MInline MI;
    typedef typename C::MCall MC;
    const typename C::Affine pa = C::aff_load(a);
    typename C::X r = C::x_inf();
    if (op == 0) r = C::template add_affine<MI>(C::from_affine(pa), C::aff_load(b));
    else if (op == 1) r = C::template add<MC>(C::from_affine(pa), C::from_affine(C::aff_load(b)));
    else if (op == 2) r = C::template dbl<MC>(C::from_affine(pa));
    else if (op == 3) r = C::template mul_small<MC>(C::from_affine(pa), (uint32_t)b[0]);
    else if (op == 4) {
        typename C::X t = C::template dbl<MI>(C::from_affine(pa));  // 2a
        t = C::template add_affine<MI>(t, pa);                      // 3a
        const typename C::Affine na = C::aff_neg(pa);
        t = C::template add_affine<MI>(t, na);                      // 2a
        t = C::template add_affine<MI>(t, na);                      // a, ZZ != 1
        r = C::template add_affine<MI>(t, C::aff_load(b));
    } else if (op == 5) {
        typename C::X t = C::template dbl<MC>(C::from_affine(pa));
        t = C::template add_affine<MC>(t, C::aff_neg(pa));  // a, ZZ != 1
        const typename C::Affine pb = C::aff_load(b);
        typename C::X u = C::template dbl<MC>(C::from_affine(pb));
        u = C::template add_affine<MC>(u, C::aff_neg(pb));  // b, ZZ != 1
        r = C::template add<MC>(t, u);
    }
    C::aff_store(out, C::to_affine(r));
}
// The whole multi-exponentiation through msm_enqueue on the host executor.  out as the device writes it (2 * AFF_WORDS words);
// returns the error flag, or -1 when the plan does not fit.  reverse != 0 runs every launch's threads in descending order.
template <class C>
int g_msm(const uint64_t* points, const uint64_t* scalars, size_t n, int scalars_mont, int c_force, int t_force, int reverse, uint64_t* out,
          uint32_t* plan_out) {
    if (n == 0) {
        memset(out, 0, 2 * C::AFF_WORDS * 8);
        return 0;
    }
    const MsmPlan pl = msm_make_plan(n, scalars_mont, c_force, t_force);
    if (plan_out) plan_out[0] = pl.c, plan_out[1] = pl.W, plan_out[2] = pl.T, plan_out[3] = pl.L, plan_out[4] = pl.nchunks, plan_out[5] = (uint32_t)pl.max_tasks;
    if ((uint64_t)pl.n * pl.W >= 0xffffffffull) return -1;
    const MsmWorkspace ws = msm_layout(pl, 8 * C::X_WORDS, 8 * C::AFF_WORDS);
    std::vector<unsigned char> buf(ws.bytes + 256, 0xA5);  // poisoned: nothing may rely on zero-initialised workspace
    unsigned char* base = (unsigned char*)(((uintptr_t)buf.data() + 255) & ~(uintptr_t)255);
    if (reverse) {
        HostExecReverse ex;
        (void)msm_enqueue<C>(ex, pl, ws, base, points, scalars);
    } else {
        HostExec ex;
        (void)msm_enqueue<C>(ex, pl, ws, base, points, scalars);
    }
    memcpy(out, base + ws.out, 2 * C::AFF_WORDS * 8);
    uint32_t flag;
    memcpy(&flag, base + ws.err, 4);
    return (int)flag;
}
}  // namespace

extern "C" {
void emu_g1_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) { g_op<G1>(op, a, b, out); }
void emu_g2_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) { g_op<G2>(op, a, b, out); }
int emu_msm(const uint64_t* points, const uint64_t* scalars, size_t n, int scalars_mont, int c_force, int t_force, int reverse, uint64_t* out16,
            uint32_t* plan_out) {
    return g_msm<G1>(points, scalars, n, scalars_mont, c_force, t_force, reverse, out16, plan_out);
}
int emu_g2_msm(const uint64_t* points, const uint64_t* scalars, size_t n, int scalars_mont, int c_force, int t_force, int reverse, uint64_t* out32,
               uint32_t* plan_out) {
    return g_msm<G2>(points, scalars, n, scalars_mont, c_force, t_force, reverse, out32, plan_out);
}
void emu_g1_add_affine(const uint64_t* a, const uint64_t* b, uint64_t* out16) { KAddAffine<G1>::run(0, a, b, out16); }
void emu_g2_add_affine(const uint64_t* a, const uint64_t* b, uint64_t* out32) { KAddAffine<G2>::run(0, a, b, out32); }
// Fp2 (fptower.E2) operations: op 0 mul (inlined base multiplier), 1 sqr, 2 add, 3 sub, 4 inv, 5 mul (out-of-line base multiplier)
void emu_fp2_op(int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
    for (size_t i = 0; i < n; i++) {
        const Fp2El x = Fp2Base::load(a + 8 * i), y = b ? Fp2Base::load(b + 8 * i) : Fp2Base::zero();
        Fp2El r = Fp2Base::zero();
        switch (op) {
            case 0: r = Fp2Mul<FpMulInline>::mul(x, y); break;
            case 1: r = Fp2Mul<FpMulInline>::sqr(x); break;
            case 2: r = Fp2Base::add(x, y); break;
            case 3: r = Fp2Base::sub(x, y); break;
            case 4: r = Fp2Base::inv(x); break;
            case 5: r = Fp2Mul<FpMulCall>::mul(x, y); break;
        }
        Fp2Base::store(out + 8 * i, r);
    }
}
}  // extern "C"

// ---- FFT -----------------------------------------------------------------------------------------------------------------------
namespace {
struct EmuDomain {
    ec::NttDomainHost h;
    std::vector<uint64_t> tw, tw_inv;
    ec::NttDomainDev d;
};
template <class Exec>
bool emu_domain(Exec& ex, uint32_t log_n, EmuDomain& e) {
    if (!ec::ntt_domain_host(log_n, e.h)) return false;
    const size_t half = ((size_t)1 << log_n) / 2;
    e.tw.assign(4 * (half ? half : 1), 0xA5A5A5A5A5A5A5A5ull);
    e.tw_inv.assign(4 * (half ? half : 1), 0xA5A5A5A5A5A5A5A5ull);
    e.d.log_n = log_n;
    e.d.tw = e.tw.data(), e.d.tw_inv = e.tw_inv.data();
    e.d.w_lo = e.h.w_lo.data(), e.d.w_hi = e.h.w_hi.data(), e.d.wi_lo = e.h.wi_lo.data(), e.d.wi_hi = e.h.wi_hi.data();
    e.d.u_lo = e.h.u_lo.data(), e.d.u_hi = e.h.u_hi.data(), e.d.u_hi_n = e.h.u_hi_n.data();
    e.d.ui_lo = e.h.ui_lo.data(), e.d.ui_hi_n = e.h.ui_hi_n.data();
    e.d.n_inv = e.h.n_inv, e.d.minus_two_inv = e.h.minus_two_inv;
    ec::ntt_domain_enqueue(ex, e.d);
    return true;
}
}  // namespace

extern "C" {
// fft.Domain.FFT / FFTInverse in place on 2^log_n elements.  decimation: 0 DIT, 1 DIF; coset 0 / 1.  Returns launches, -1 on a bad size.
int emu_fft(uint64_t* a, uint32_t log_n, int decimation, int coset, int inverse, int reverse) {
    EmuDomain e;
    if (reverse) {
        HostExecReverse ex;
        if (!emu_domain(ex, log_n, e)) return -1;
        return ec::fft_enqueue(ex, e.d, a, decimation == 1, coset, inverse != 0);
    }
    HostExec ex;
    if (!emu_domain(ex, log_n, e)) return -1;
    return ec::fft_enqueue(ex, e.d, a, decimation == 1, coset, inverse != 0);
}
// computeH: a, b, c hold n_in elements; h_out 2^log_n elements (regular form, bit-reversed coefficient order)
int emu_compute_h(const uint64_t* a, const uint64_t* b, const uint64_t* c, size_t n_in, uint32_t log_n, int reverse, uint64_t* h_out) {
    const size_t n = (size_t)1 << log_n;
    if (n_in > n) return -1;
    std::vector<uint64_t> va(4 * n, 0), vb(4 * n, 0), vc(4 * n, 0);
    memcpy(va.data(), a, 32 * n_in);
    memcpy(vb.data(), b, 32 * n_in);
    memcpy(vc.data(), c, 32 * n_in);
    EmuDomain e;
    int launches;
    if (reverse) {
        HostExecReverse ex;
        if (!emu_domain(ex, log_n, e)) return -1;
        launches = ec::compute_h_enqueue(ex, e.d, va.data(), vb.data(), vc.data());
    } else {
        HostExec ex;
        if (!emu_domain(ex, log_n, e)) return -1;
        launches = ec::compute_h_enqueue(ex, e.d, va.data(), vb.data(), vc.data());
    }
    memcpy(h_out, va.data(), 32 * n);
    return launches;
}
// Domain constants as the product derives them: out[0..4) Generator, [4..8) FinerGenerator, [8..12) CardinalityInv, [12..16) (-2)^-1
int emu_fft_domain(uint32_t log_n, uint64_t* out) {
    ec::NttDomainHost h;
    if (!ec::ntt_domain_host(log_n, h)) return -1;
    memcpy(out, h.generator, 32);
    memcpy(out + 4, h.finer_generator, 32);
    memcpy(out + 8, h.n_inv, 32);
    memcpy(out + 12, h.minus_two_inv, 32);
    return 0;
}
uint32_t emu_ntt_rev(uint32_t x, uint32_t log_n) { return ec::ntt_rev(x, log_n); }

// ---- ComputeGroth16Proof: the product's sequencing (groth16.hpp) over the emulated kernels ------------------------------------------
struct EmuGroth16In {
    const uint64_t *g1_a, *g1_b, *g1_z, *g2_b;      // base arrays: n_a, n_b, cardinality, n_b points
    const uint64_t *alpha, *beta, *delta, *beta2, *delta2;
    const uint64_t *a, *b, *c;                      // n_constraints Montgomery elements each
    const uint64_t *wa, *wb;                        // n_a / n_b scalars
    uint64_t n_constraints, n_a, n_b;
    uint32_t log_n;
    int scalars_mont;
};
}  // extern "C"
namespace {
struct EmuGroth16Ops {
    const EmuGroth16In& in;
    std::vector<uint64_t> h;
    const uint64_t* alpha() const { return in.alpha; }
    const uint64_t* beta() const { return in.beta; }
    const uint64_t* delta() const { return in.delta; }
    const uint64_t* beta2() const { return in.beta2; }
    const uint64_t* delta2() const { return in.delta2; }
    int smul_g1(const uint64_t* pt, const uint64_t* k, uint64_t* out) {
        uint64_t r[16];
        const int rc = g_msm<G1>(pt, k, 1, 0, 0, 0, 0, r, nullptr);
        memcpy(out, r, 64);
        return rc;
    }
    int smul_g2(const uint64_t* pt, const uint64_t* k, uint64_t* out) {
        uint64_t r[32];
        const int rc = g_msm<G2>(pt, k, 1, 0, 0, 0, 0, r, nullptr);
        memcpy(out, r, 128);
        return rc;
    }
    int add_g1(const uint64_t* x, const uint64_t* y, uint64_t* out) {
        uint64_t r[16];
        KAddAffine<G1>::run(0, x, y, r);
        memcpy(out, r, 64);
        return 0;
    }
    int add_g2(const uint64_t* x, const uint64_t* y, uint64_t* out) {
        uint64_t r[32];
        KAddAffine<G2>::run(0, x, y, r);
        memcpy(out, r, 128);
        return 0;
    }
    int compute_h() {
        h.assign(4 * ((size_t)1 << in.log_n), 0);
        return emu_compute_h(in.a, in.b, in.c, in.n_constraints, in.log_n, 0, h.data()) > 0 ? 0 : -1;
    }
    int msm_g1(int which, uint64_t* out) {
        uint64_t r[16];
        const int rc = which == 0 ? g_msm<G1>(in.g1_a, in.wa, in.n_a, in.scalars_mont, 0, 0, 0, r, nullptr)
                                  : g_msm<G1>(in.g1_b, in.wb, in.n_b, in.scalars_mont, 0, 0, 1, r, nullptr);
        memcpy(out, r, 64);
        return rc;
    }
    int msm_g1_h(uint64_t* out) {
        uint64_t r[16];
        const int rc = g_msm<G1>(in.g1_z, h.data(), (size_t)1 << in.log_n, 0, 0, 0, 0, r, nullptr);
        memcpy(out, r, 64);
        return rc;
    }
    int msm_g2(uint64_t* out) {
        uint64_t r[32];
        const int rc = g_msm<G2>(in.g2_b, in.wb, in.n_b, in.scalars_mont, 0, 0, 0, r, nullptr);
        memcpy(out, r, 128);
        return rc;
    }
};
}  // namespace
extern "C" {
int emu_groth16(const EmuGroth16In* in, const uint64_t* r, const uint64_t* s, uint64_t* ar, uint64_t* bs, uint64_t* krs) {
    EmuGroth16Ops ops{*in, {}};
    ec::Groth16Out out;
    const int rc = ec::groth16_compose(ops, r, s, out);
    if (rc) return rc;
    memcpy(ar, out.ar, 64);
    memcpy(bs, out.bs, 128);
    memcpy(krs, out.krs, 64);
    return 0;
}
}
