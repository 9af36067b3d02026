// Host-side check of the Karatsuba multiplier (tools/exp/fr_kara.cuh, portable fallbacks) against the host multiplier
// (csrc/fr_host.hpp) and a schoolbook 512-bit product.  Build: nvcc -O2 -std=c++17 -o /tmp/kara_host_test tools/exp/kara_host_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "fr_kara.cuh"
#include "../../gkr-mimc_b200/csrc/fr_host.hpp"
namespace H = gkr::host;
using gkr::Fr;

static Fr to_dev(const H::Fr& x) {
    Fr r;
    for (int i = 0; i < 4; i++) r.v[2 * i] = (uint32_t)x.l[i], r.v[2 * i + 1] = (uint32_t)(x.l[i] >> 32);
    return r;
}
static H::Fr to_host(const Fr& x) {
    H::Fr r;
    for (int i = 0; i < 4; i++) r.l[i] = (uint64_t)x.v[2 * i] | ((uint64_t)x.v[2 * i + 1] << 32);
    return r;
}
static void school(const Fr& a, const Fr& b, uint32_t* t) {
    uint64_t acc[17] = {0};
    for (int i = 0; i < 16; i++) t[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) {
            unsigned __int128 v = (unsigned __int128)a.v[j] * b.v[i] + t[i + j] + c;
            t[i + j] = (uint32_t)v;
            c = (uint64_t)(v >> 32);
        }
        int k = i + 8;
        while (c && k < 16) {
            uint64_t v = (uint64_t)t[k] + c;
            t[k] = (uint32_t)v;
            c = v >> 32;
            k++;
        }
    }
    (void)acc;
}
int main() {
    std::mt19937_64 rng(12345);
    std::vector<H::Fr> vals;
    auto canon = [&](H::Fr x) {
        x.l[3] %= H::Q[3];  // top limb below q's top limb => canonical
        return x;
    };
    const uint64_t pat[] = {0, 1, 2, 0xffffffffULL, 0x100000000ULL, 0xffffffffffffffffULL, 0x8000000000000000ULL, 0x7fffffffffffffffULL};
    for (uint64_t p0 : pat)
        for (uint64_t p1 : pat)
            for (uint64_t p2 : {pat[0], pat[5], pat[3]})
                for (uint64_t p3 : {(uint64_t)0, (uint64_t)1, (uint64_t)(H::Q[3] - 1)}) vals.push_back(canon(H::Fr{{p0, p1, p2, p3}}));
    vals.push_back(H::Fr{{H::Q[0] - 1, H::Q[1], H::Q[2], H::Q[3]}});  // q - 1
    vals.push_back(H::one());
    // values with equal halves / ordered halves (the sign paths of the subtractive middle term)
    vals.push_back(canon(H::Fr{{5, 7, 5, 7}}));
    vals.push_back(canon(H::Fr{{9, 9, 1, 1}}));
    vals.push_back(canon(H::Fr{{1, 1, 9, 9}}));
    for (int i = 0; i < 2000; i++) vals.push_back(canon(H::Fr{{rng(), rng(), rng(), rng()}}));
    long bad = 0, n = 0;
    auto check = [&](const H::Fr& x, const H::Fr& y) {
        const Fr a = to_dev(x), b = to_dev(y);
        uint32_t t[16], u[16];
        gkr::hd_mul_wide_k(a, b, t);
        school(a, b, u);
        if (memcmp(t, u, sizeof t)) {
            if (bad < 5) printf("wide mismatch\n");
            bad++;
        }
        const H::Fr want = H::mul(x, y), got = to_host(gkr::hd_mul_k(a, b));
        if (!H::eq(want, got)) {
            if (bad < 5) printf("mul mismatch x=%016llx %016llx %016llx %016llx y=%016llx %016llx %016llx %016llx\n", (unsigned long long)x.l[3], (unsigned long long)x.l[2], (unsigned long long)x.l[1], (unsigned long long)x.l[0], (unsigned long long)y.l[3], (unsigned long long)y.l[2], (unsigned long long)y.l[1], (unsigned long long)y.l[0]);
            bad++;
        }
        n++;
    };
    for (size_t i = 0; i < vals.size(); i += 7)
        for (size_t j = 0; j < vals.size(); j += 3) check(vals[i], vals[j]);
    for (size_t i = 0; i < vals.size(); i++) check(vals[i], vals[i]);
    for (int i = 0; i < 2000000; i++) check(canon(H::Fr{{rng(), rng(), rng(), rng()}}), canon(H::Fr{{rng(), rng(), rng(), rng()}}));
    printf("%ld products checked, %ld mismatches\n", n, bad);
    return bad != 0;
}
